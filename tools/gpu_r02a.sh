#!/bin/bash
# round 2, first GPU pass: the resident engine's parity tests, the whole GPU suite, first bench lines
mkdir -p gpurun_out
T=r02a
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${T}_smi.txt 2>&1
timeout -k 10 900 python -m pytest tests/test_gpu_resident.py -x -q -m gpu --timeout 600 > gpurun_out/${T}_pytest_resident.log 2>&1
echo "resident rc=$?" | tee -a gpurun_out/${T}_pytest_resident.log
tail -5 gpurun_out/${T}_pytest_resident.log
timeout -k 10 1500 python -m pytest tests -q -m gpu --timeout 600 --deselect tests/test_gpu_resident.py > gpurun_out/${T}_pytest_gpu.log 2>&1
echo "gpu suite rc=$?" | tee -a gpurun_out/${T}_pytest_gpu.log
tail -5 gpurun_out/${T}_pytest_gpu.log
timeout -k 10 600 python bench.py --steps 20 --warmup 5 > gpurun_out/${T}_bench_20.json 2> gpurun_out/${T}_bench_20.err
echo "bench20 rc=$?"; cut -c1-1500 gpurun_out/${T}_bench_20.json; tail -3 gpurun_out/${T}_bench_20.err
timeout -k 10 600 python bench.py > gpurun_out/${T}_bench_1080p.json 2> gpurun_out/${T}_bench_1080p.err
echo "bench rc=$?"; cut -c1-600 gpurun_out/${T}_bench_1080p.json; tail -3 gpurun_out/${T}_bench_1080p.err
