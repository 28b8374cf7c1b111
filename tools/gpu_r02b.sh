#!/bin/bash
# resident tests again + A/B of the release fence
mkdir -p gpurun_out
T=r02b
timeout -k 10 1200 python -m pytest tests/test_gpu_resident.py -q -m gpu --timeout 600 > gpurun_out/${T}_pytest_resident.log 2>&1
echo "resident rc=$?"; tail -15 gpurun_out/${T}_pytest_resident.log
for w in 1080p 4k; do
  timeout -k 10 300 python bench.py --workload $w --no-extras --no-cpu-baseline --steps 512 > gpurun_out/${T}_bench_${w}_fence.json 2> gpurun_out/${T}_bench_${w}_fence.err
  OAT_B200_RELAXED_PUBLISH=1 timeout -k 10 300 python bench.py --workload $w --no-extras --no-cpu-baseline --steps 512 > gpurun_out/${T}_bench_${w}_relaxed.json 2> gpurun_out/${T}_bench_${w}_relaxed.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02b_bench_*.json')):
    try:
        d=json.load(open(f)); r=d['roofline']
        print(f, 'value',round(d['value']), 'frac',round(r['frac'],3),'ms/frame',round(r['ms_per_frame']*1e3,2),'e2e',round(d['e2e']['value']))
    except Exception as e: print(f,e)
PY
