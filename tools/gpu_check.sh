#!/bin/bash
# what the driver runs at round end, in its order: smoke(), the GPU suite, the reference arm, the bench (20 steps and default)
T=${1:-r02z}
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 2
timeout -k 10 1500 python -m pytest tests -x -q -m gpu > gpurun_out/${T}_pytest_gpu.log 2>&1
tail -n 3 gpurun_out/${T}_pytest_gpu.log
timeout -k 10 300 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/${T}_bench_reference_20steps.json 2> gpurun_out/${T}_bench_reference_20steps.err
timeout -k 10 300 python bench.py --steps 20 --warmup 5 > gpurun_out/${T}_bench_1080p_20steps.json 2> gpurun_out/${T}_bench_1080p_20steps.err
timeout -k 10 600 python bench.py > gpurun_out/${T}_bench_1080p.json 2> gpurun_out/${T}_bench_1080p.err
for f in reference_20steps 1080p_20steps 1080p; do echo "== $f"; cut -c1-330 gpurun_out/${T}_bench_$f.json; tail -n 2 gpurun_out/${T}_bench_$f.err; done
