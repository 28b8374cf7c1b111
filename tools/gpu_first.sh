#!/bin/bash
# first GPU pass: parity tests, smoke, bench, ncu launch list
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
lscpu | head -20 > gpurun_out/lscpu.txt; nproc >> gpurun_out/lscpu.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
tail -3 gpurun_out/smoke.log
timeout 600 python bench.py --steps 1000 --warmup 50 > gpurun_out/bench_1080p.json 2> gpurun_out/bench_1080p.err; echo "bench rc=$?"
cat gpurun_out/bench_1080p.json; tail -3 gpurun_out/bench_1080p.err
timeout 300 python bench.py --steps 500 --warmup 50 --alpha 0 --no-cpu-baseline > gpurun_out/bench_1080p_a0.json 2> gpurun_out/bench_1080p_a0.err
cat gpurun_out/bench_1080p_a0.json
timeout 300 python bench.py --steps 300 --warmup 30 --workload 4k --no-cpu-baseline > gpurun_out/bench_4k.json 2> gpurun_out/bench_4k.err
cat gpurun_out/bench_4k.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_1080p.csv python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
echo "ncu launches rc=$?"
