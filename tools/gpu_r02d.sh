#!/bin/bash
# conditional release fence; variant B (4 stages, 24 reserved CTA slots, 48 tail CTAs) vs A (3 stages, 80 regs, co-resident tail)
mkdir -p gpurun_out
T=r02d
timeout -k 10 1200 python -m pytest tests/test_gpu_resident.py tests/test_gpu_tracker.py -q -m gpu --timeout 600 -x > gpurun_out/${T}_pytest_B.log 2>&1
echo "B tests rc=$?"; tail -4 gpurun_out/${T}_pytest_B.log
OAT_B200_LIB=$PWD/oat_b200/liboatgpu_s3.so timeout -k 10 1200 python -m pytest tests/test_gpu_resident.py tests/test_gpu_tracker.py -q -m gpu --timeout 600 -x > gpurun_out/${T}_pytest_A.log 2>&1
echo "A tests rc=$?"; tail -4 gpurun_out/${T}_pytest_A.log
for w in 1080p 4k; do
  timeout -k 10 300 python bench.py --workload $w --no-extras --no-cpu-baseline --steps 2000 > gpurun_out/${T}_bench_${w}_B.json 2> gpurun_out/${T}_bench_${w}_B.err
  OAT_B200_LIB=$PWD/oat_b200/liboatgpu_s3.so timeout -k 10 300 python bench.py --workload $w --no-extras --no-cpu-baseline --steps 2000 > gpurun_out/${T}_bench_${w}_A.json 2> gpurun_out/${T}_bench_${w}_A.err
  OAT_B200_FENCE_ALWAYS=1 timeout -k 10 300 python bench.py --workload $w --no-extras --no-cpu-baseline --steps 512 > gpurun_out/${T}_bench_${w}_Bfence.json 2> gpurun_out/${T}_bench_${w}_Bfence.err
done
timeout -k 10 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/${T}_bench_20_B.json 2> gpurun_out/${T}_bench_20_B.err
OAT_B200_LIB=$PWD/oat_b200/liboatgpu_s3.so timeout -k 10 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/${T}_bench_20_A.json 2> gpurun_out/${T}_bench_20_A.err
timeout -k 10 200 python tools/clip_rate.py 1080p > gpurun_out/${T}_clip_rate_B.txt 2>&1
OAT_B200_LIB=$PWD/oat_b200/liboatgpu_s3.so timeout -k 10 200 python tools/clip_rate.py 1080p > gpurun_out/${T}_clip_rate_A.txt 2>&1
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02d_bench_*.json')):
    try:
        d=json.load(open(f)); r=d['roofline']
        print(f, 'value',round(d['value']), 'frac',round(r['frac'],3),'us/frame',round(r['ms_per_frame']*1e3,2),'e2e',round(d['e2e']['value']), 'host us/frame', round(d['host']['call_us_per_frame'],1), {k:(round(v['value']) if isinstance(v,dict) and 'value' in v else None) for k,v in d.items() if k in ('multi_stream','multi_blob','config4_8x4k_per_gpu')})
    except Exception as e: print(f,e)
PY
cat gpurun_out/${T}_clip_rate_B.txt gpurun_out/${T}_clip_rate_A.txt
