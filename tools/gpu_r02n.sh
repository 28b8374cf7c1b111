#!/bin/bash
mkdir -p gpurun_out
T=${1:-r02n}
timeout -k 10 900 python tools/graph_bench.py --frames 1000 --out gpurun_out/${T}_graph_bench.json > gpurun_out/${T}_graph_bench.txt 2>&1
echo "graph rc=$?"; grep -vE "mogfilt\[|colorconvert\[" gpurun_out/${T}_graph_bench.txt | cut -c1-330
i=0
for k in "equals_frame_by_frame and shape0 and 8-0.02" "close_dependencies and shape1"; do
  i=$((i+1))
  OAT_STRESS_REPS=2 timeout 900 compute-sanitizer --tool racecheck --error-exitcode 3 python -X faulthandler -m pytest tests/test_gpu_resident.py -m gpu -x -q -k "$k" -p no:cacheprovider > gpurun_out/${T}_sanitize_racecheck_small$i.log 2>&1
  echo "== racecheck [$k] rc=$?"; grep -vE "^=========     " gpurun_out/${T}_sanitize_racecheck_small$i.log | tail -n 12 | cut -c1-250
done
