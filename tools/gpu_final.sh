#!/bin/bash
# round-end evidence: the whole GPU suite, the sanitizer pass over the late additions, then tools/gpu_profiles.sh
T=${1:-r02y}
mkdir -p gpurun_out
timeout -k 10 1500 python -m pytest tests -q -m gpu > gpurun_out/${T}_pytest_gpu.log 2>&1
tail -n 3 gpurun_out/${T}_pytest_gpu.log
bash tools/gpu_sanitize2.sh $T
bash tools/gpu_profiles.sh $T
timeout -k 10 300 python tools/tail_probe.py > gpurun_out/${T}_tail_probe.txt 2>&1
cat gpurun_out/${T}_tail_probe.txt
