#!/bin/bash
# batched publication (one release fence per 2 tiles): tests + stress, then B (batch 2) vs b3 (batch 3) vs A (3 stages)
mkdir -p gpurun_out
T=r02e
timeout -k 10 1500 python -m pytest tests/test_gpu_resident.py tests/test_gpu_tracker.py -q -m gpu --timeout 900 > gpurun_out/${T}_pytest_B.log 2>&1
echo "B tests rc=$?"; tail -4 gpurun_out/${T}_pytest_B.log
OAT_B200_LIB=$PWD/oat_b200/liboatgpu_b3.so timeout -k 10 900 python -m pytest tests/test_gpu_resident.py -q -m gpu --timeout 900 -k "stress or equals or interleaved" > gpurun_out/${T}_pytest_b3.log 2>&1
echo "b3 tests rc=$?"; tail -4 gpurun_out/${T}_pytest_b3.log
OAT_B200_LIB=$PWD/oat_b200/liboatgpu_s3.so timeout -k 10 900 python -m pytest tests/test_gpu_resident.py -q -m gpu --timeout 900 -k "stress or equals or interleaved" > gpurun_out/${T}_pytest_A.log 2>&1
echo "A tests rc=$?"; tail -4 gpurun_out/${T}_pytest_A.log
OAT_B200_RELAXED_PUBLISH=1 timeout -k 10 900 python -m pytest tests/test_gpu_resident.py -q -m gpu --timeout 900 -k "stress" > gpurun_out/${T}_pytest_relaxed.log 2>&1
echo "relaxed (expected to FAIL: shows the stress test detects a missing fence) rc=$?"; tail -4 gpurun_out/${T}_pytest_relaxed.log
for w in 1080p 4k; do
  for v in B b3 A; do
    lib=$PWD/oat_b200/liboatgpu.so; [ $v = b3 ] && lib=$PWD/oat_b200/liboatgpu_b3.so; [ $v = A ] && lib=$PWD/oat_b200/liboatgpu_s3.so
    OAT_B200_LIB=$lib timeout -k 10 300 python bench.py --workload $w --no-extras --no-cpu-baseline --steps 2000 > gpurun_out/${T}_bench_${w}_${v}.json 2> gpurun_out/${T}_bench_${w}_${v}.err
  done
done
timeout -k 10 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/${T}_bench_20_B.json 2> gpurun_out/${T}_bench_20_B.err
timeout -k 10 200 python tools/clip_rate.py 1080p > gpurun_out/${T}_clip_rate_B.txt 2>&1
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02e_bench_*.json')):
    try:
        d=json.load(open(f)); r=d['roofline']
        print(f, 'value',round(d['value']), 'frac',round(r['frac'],3),'us/frame',round(r['ms_per_frame']*1e3,2),'e2e',round(d['e2e']['value']), 'host us/frame', round(d['host']['call_us_per_frame'],1), {k:(round(v['value']) if isinstance(v,dict) and 'value' in v else None) for k,v in d.items() if k in ('multi_stream','multi_blob','config4_8x4k_per_gpu')})
    except Exception as e: print(f,e)
PY
cat gpurun_out/${T}_clip_rate_B.txt
