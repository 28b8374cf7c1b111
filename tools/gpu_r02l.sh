#!/bin/bash
mkdir -p gpurun_out
T=${1:-r02l}
eval timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 3 python -m pytest tests/test_gpu_resident.py -m gpu -x -q -k '"close_dependencies or (equals_frame_by_frame and 120-160)"' -p no:cacheprovider > gpurun_out/${T}_sanitize_racecheck.log 2>&1
echo "== racecheck small rc=$?"; grep -E "RACECHECK SUMMARY|passed|failed|rror" gpurun_out/${T}_sanitize_racecheck.log | tail -n 4 | cut -c1-200
timeout -k 10 600 python bench.py > gpurun_out/${T}_bench_1080p.json 2> gpurun_out/${T}_bench_1080p.err
timeout -k 10 300 python bench.py --workload 1mp --steps 1000 --no-extras > gpurun_out/${T}_bench_1mp.json 2> gpurun_out/${T}_bench_1mp.err
python - <<P
import json
for w in ("1080p", "1mp"):
    d = json.loads(open("gpurun_out/${T}_bench_%s.json" % w).read().splitlines()[-1])
    print(w, "value", d["value"], "frac", d["roofline"]["frac"], "e2e", d["e2e"]["value"], "multi_blob", d.get("multi_blob"), "cfg4", d.get("config4_8x4k_per_gpu"))
P
timeout -k 10 600 python tools/graph_bench.py --frames 1000 --out gpurun_out/${T}_graph_bench.json > gpurun_out/${T}_graph_bench.txt 2>&1
echo "graph rc=$?"; grep -E "track_pipeline64|tracker\[" gpurun_out/${T}_graph_bench.txt | cut -c1-330
