#!/bin/bash
# Round evidence: bench lines, clocks during the bench, ncu launch list of the bench command, ncu --set full
# captures of the resident kernels, host-rate tools.  usage: gpu_profiles.sh <tag>   (run under gpurun, 1 GPU)
TAG=${1:-r02}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 200 > gpurun_out/${TAG}_clocks.csv &
SMI=$!
timeout -k 10 600 python bench.py > gpurun_out/${TAG}_bench_1080p.json 2> gpurun_out/${TAG}_bench_1080p.err
timeout -k 10 300 python bench.py --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_1080p_20steps.json 2> gpurun_out/${TAG}_bench_1080p_20steps.err
timeout -k 10 600 python bench.py --workload 4k --steps 600 > gpurun_out/${TAG}_bench_4k.json 2> gpurun_out/${TAG}_bench_4k.err
timeout -k 10 300 python bench.py --alpha 0 --steps 1000 --no-cpu-baseline --no-extras > gpurun_out/${TAG}_bench_1080p_a0.json 2> gpurun_out/${TAG}_bench_1080p_a0.err
timeout -k 10 300 python bench.py --workload 1mp --steps 1000 --no-extras > gpurun_out/${TAG}_bench_1mp.json 2> gpurun_out/${TAG}_bench_1mp.err
timeout -k 10 300 python bench.py --workload 1mp-tight --steps 1000 --no-extras --no-cpu-baseline > gpurun_out/${TAG}_bench_1mp_tight.json 2> gpurun_out/${TAG}_bench_1mp_tight.err
timeout -k 10 300 python bench.py --impl reference --steps 100 --warmup 10 > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err
kill $SMI
timeout -k 10 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches_bench.csv python bench.py --steps 64 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/${TAG}_ncu_launches.log 2>&1
timeout -k 10 600 ncu --set full --clock-control none --import-source on -k regex:mog_stream -s 2 -c 1 -f -o gpurun_out/${TAG}_fused_1080p python tools/kbench.py --res 1080p --fused-only > gpurun_out/${TAG}_ncu_fused_1080p.log 2>&1
timeout -k 10 600 ncu --set full --clock-control none --import-source on -k regex:mog_stream -s 2 -c 1 -f -o gpurun_out/${TAG}_fused_4k python tools/kbench.py --res 4k --frames 2 --fused-only > gpurun_out/${TAG}_ncu_fused_4k.log 2>&1
timeout -k 10 600 ncu --set full --clock-control none --import-source on -k regex:tail_stream -s 2 -c 1 -f -o gpurun_out/${TAG}_tail_1080p python tools/kbench.py --res 1080p > gpurun_out/${TAG}_ncu_tail_1080p.log 2>&1
timeout -k 10 200 python tools/hostrate.py 1080p > gpurun_out/${TAG}_hostrate.log 2>&1
timeout -k 10 200 python tools/clip_rate.py 1080p > gpurun_out/${TAG}_clip_rate.txt 2>&1
timeout -k 10 200 python tools/clip_rate.py 4k >> gpurun_out/${TAG}_clip_rate.txt 2>&1
for f in 1080p 1080p_20steps 4k 1080p_a0 1mp 1mp_tight reference; do echo "== $f"; cat gpurun_out/${TAG}_bench_$f.json | cut -c1-400; done
cat gpurun_out/${TAG}_hostrate.log gpurun_out/${TAG}_clip_rate.txt
tail -n 3 gpurun_out/${TAG}_ncu_fused_1080p.log; tail -n 3 gpurun_out/${TAG}_ncu_tail_1080p.log
