#!/bin/bash
mkdir -p gpurun_out
T=r02h
timeout -k 10 400 python -m pytest tests/test_gpu_pipeline.py -q -m gpu --timeout 150 -x > gpurun_out/${T}_pytest_pipeline.log 2>&1
echo "pipeline tests rc=$?"; tail -5 gpurun_out/${T}_pytest_pipeline.log | cut -c1-300
timeout -k 10 400 python tools/graph_bench.py --frames 1000 --out gpurun_out/${T}_graph_bench.json > gpurun_out/${T}_graph_bench.txt 2>&1
echo "graph rc=$?"; cat gpurun_out/${T}_graph_bench.txt | tail -60 | cut -c1-400
timeout -k 10 900 python -m pytest tests -q -m gpu --timeout 300 --deselect tests/test_gpu_pipeline.py > gpurun_out/${T}_pytest_gpu.log 2>&1
echo "gpu suite rc=$?"; tail -5 gpurun_out/${T}_pytest_gpu.log | cut -c1-300
timeout -k 10 200 python tools/hostrate.py 1080p > gpurun_out/${T}_hostrate.log 2>&1; cat gpurun_out/${T}_hostrate.log
timeout -k 10 100 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
