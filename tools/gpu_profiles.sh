#!/bin/bash
# Round evidence: bench lines, clocks during the bench, ncu launch list of the bench command, ncu --set full
# captures of the two kernels of the path.  usage: gpu_profiles.sh <tag>
TAG=${1:-r01}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1; tail -5 gpurun_out/${TAG}_pytest_gpu.log
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 200 > gpurun_out/${TAG}_clocks.csv &
SMI=$!
python bench.py > gpurun_out/${TAG}_bench_1080p.json 2> gpurun_out/${TAG}_bench_1080p.err
python bench.py --workload 4k --steps 600 > gpurun_out/${TAG}_bench_4k.json 2> gpurun_out/${TAG}_bench_4k.err
python bench.py --alpha 0 --steps 1000 --no-cpu-baseline > gpurun_out/${TAG}_bench_1080p_a0.json 2> gpurun_out/${TAG}_bench_1080p_a0.err
python bench.py --workload 4k --steps 400 --streams 8 --no-cpu-baseline > gpurun_out/${TAG}_bench_4k_8streams.json 2> gpurun_out/${TAG}_bench_4k_8streams.err
python bench.py --steps 1000 --streams 8 --no-cpu-baseline > gpurun_out/${TAG}_bench_1080p_8streams.json 2> gpurun_out/${TAG}_bench_1080p_8streams.err
python bench.py --impl reference --steps 100 --warmup 10 > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err
kill $SMI
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 120 -c 400 --csv --log-file gpurun_out/${TAG}_launches_bench.csv python bench.py --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/${TAG}_ncu_launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mog_pipe -s 40 -c 1 -f -o gpurun_out/${TAG}_fused_1080p python tools/kbench.py --res 1080p --steps 20 > gpurun_out/${TAG}_ncu_fused_1080p.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mog_pipe -s 40 -c 1 -f -o gpurun_out/${TAG}_fused_4k python tools/kbench.py --res 4k --steps 20 > gpurun_out/${TAG}_ncu_fused_4k.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tail_fast -s 40 -c 1 -f -o gpurun_out/${TAG}_tail_1080p python tools/kbench.py --res 1080p --steps 20 > gpurun_out/${TAG}_ncu_tail_1080p.log 2>&1
timeout 120 python tools/hostrate.py 1080p > gpurun_out/${TAG}_hostrate.log 2>&1; timeout 120 python tools/hostrate.py 4k >> gpurun_out/${TAG}_hostrate.log 2>&1; cat gpurun_out/${TAG}_hostrate.log
for f in 1080p 4k 1080p_a0 4k_8streams 1080p_8streams reference; do echo "== $f"; cat gpurun_out/${TAG}_bench_$f.json | cut -c1-600; done
