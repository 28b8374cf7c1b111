#!/bin/bash
# ncu --set full capture of the fused kernel (1080p and 4k), plus kbench timings
mkdir -p gpurun_out
TAG=${1:-r01}
timeout 300 python tools/kbench.py --res 1080p --alpha 0.01 --steps 200 2>&1 | tail -1
timeout 300 python tools/kbench.py --res 1080p --alpha 0.01 --steps 200 --noflush 2>&1 | tail -1
timeout 300 python tools/kbench.py --res 4k --alpha 0.01 --steps 100 2>&1 | tail -1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mog_fused -s 40 -c 2 -f -o gpurun_out/fused_1080p_$TAG python tools/kbench.py --res 1080p --alpha 0.01 --steps 20 > gpurun_out/ncu_fused_$TAG.log 2>&1
echo "ncu rc=$?"
