#!/bin/bash
mkdir -p gpurun_out
T=${1:-r02q}
timeout -k 10 900 python -m pytest tests/test_gpu_resident.py tests/test_gpu_detect.py tests/test_gpu_tracker.py -q -m gpu --timeout 300 -x > gpurun_out/${T}_pytest.log 2>&1
echo "tests rc=$?"; tail -n 4 gpurun_out/${T}_pytest.log | cut -c1-300
python tools/tail_probe.py 2>&1 | tee gpurun_out/${T}_tail_probe.txt | cut -c1-400
