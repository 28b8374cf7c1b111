#!/bin/bash
mkdir -p gpurun_out
T=${1:-r02s}
timeout -k 10 1500 python -m pytest tests -q -m gpu --timeout 300 -x > gpurun_out/${T}_pytest_gpu.log 2>&1
echo "gpu suite rc=$?"; tail -n 4 gpurun_out/${T}_pytest_gpu.log | cut -c1-300
python tools/tail_probe.py 2>&1 | tee gpurun_out/${T}_tail_probe.txt | cut -c1-330
