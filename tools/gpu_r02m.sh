#!/bin/bash
mkdir -p gpurun_out
T=${1:-r02m}
timeout -k 10 600 python -m pytest tests/test_gpu_resident.py tests/test_gpu_pipeline.py -q -m gpu --timeout 200 -x -k "stream or pipeline or buffer or track" > gpurun_out/${T}_pytest_stream.log 2>&1
echo "tests rc=$?"; tail -n 3 gpurun_out/${T}_pytest_stream.log | cut -c1-300
timeout -k 10 300 python bench.py --workload 1mp --steps 1000 --no-extras --no-cpu-baseline > gpurun_out/${T}_bench_1mp.json 2> gpurun_out/${T}_bench_1mp.err
python - <<P
import json
d = json.loads(open("gpurun_out/${T}_bench_1mp.json").read().splitlines()[-1])
print("1mp value", d["value"], "frac", d["roofline"]["frac"], "e2e", d["e2e"]["value"])
P
timeout -k 10 600 python tools/graph_bench.py --frames 1000 --out gpurun_out/${T}_graph_bench.json > gpurun_out/${T}_graph_bench.txt 2>&1
echo "graph rc=$?"; grep -E "track_|tracker\[" gpurun_out/${T}_graph_bench.txt | cut -c1-330
for k in "equals_frame_by_frame and 120-160 and 8-" "close_dependencies and 480"; do
  timeout 600 compute-sanitizer --tool racecheck --error-exitcode 3 python -X faulthandler -m pytest tests/test_gpu_resident.py -m gpu -x -q -k "$k" -p no:cacheprovider > gpurun_out/${T}_sanitize_racecheck_small.log 2>&1
  echo "== racecheck [$k] rc=$?"; grep -vE "^=========     " gpurun_out/${T}_sanitize_racecheck_small.log | tail -n 15 | cut -c1-250
done
