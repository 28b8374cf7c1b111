#!/bin/bash
# device-side labelling fallback, bigger bands, pipelined component, graph bench
mkdir -p gpurun_out
T=r02f
timeout -k 10 1500 python -m pytest tests/test_gpu_resident.py tests/test_gpu_tracker.py tests/test_gpu_pipeline.py -q -m gpu --timeout 900 > gpurun_out/${T}_pytest.log 2>&1
echo "tests rc=$?"; tail -6 gpurun_out/${T}_pytest.log
timeout -k 10 300 python bench.py --steps 20 --warmup 5 > gpurun_out/${T}_bench_20.json 2> gpurun_out/${T}_bench_20.err
timeout -k 10 600 python bench.py > gpurun_out/${T}_bench_1080p.json 2> gpurun_out/${T}_bench_1080p.err
timeout -k 10 200 python tools/clip_rate.py 1080p > gpurun_out/${T}_clip_rate.txt 2>&1
timeout -k 10 900 python tools/graph_bench.py --frames 1000 --out gpurun_out/${T}_graph_bench.json > gpurun_out/${T}_graph_bench.txt 2>&1
echo "graph rc=$?"
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02f_bench_*.json')):
    try:
        d=json.load(open(f)); r=d['roofline']
        print(f, 'value',round(d['value']), 'frac',round(r['frac'],3),'us/frame',round(r['ms_per_frame']*1e3,2),'e2e',round(d['e2e']['value']), 'host us/frame', round(d['host']['call_us_per_frame'],1), {k:(round(v['value']) if isinstance(v,dict) and 'value' in v else v) for k,v in d.items() if k in ('multi_stream','multi_blob','config4_8x4k_per_gpu')}, d.get('framefilt_mog_egress',{}).get('frac'), d.get('cold_frame'))
    except Exception as e: print(f,e)
PY
cat gpurun_out/${T}_clip_rate.txt; cat gpurun_out/${T}_graph_bench.txt | tail -25
