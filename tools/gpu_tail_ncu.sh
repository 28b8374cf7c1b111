#!/bin/bash
# ncu --set full of the tail server on a multi-blob clip (source-level stall attribution).  usage: gpu_tail_ncu.sh <tag>
T=${1:-r02}
mkdir -p gpurun_out
which nvidia-cuda-mps-control nvidia-smi | head -3
timeout -k 10 600 ncu --set full --clock-control none --import-source on -k regex:tail_stream -s 1 -c 1 -f -o gpurun_out/${T}_tail_multiblob python tools/tail_probe.py --blobs 60 --frames 24 > gpurun_out/${T}_ncu_tail_multiblob.log 2>&1
tail -n 4 gpurun_out/${T}_ncu_tail_multiblob.log | cut -c1-300
