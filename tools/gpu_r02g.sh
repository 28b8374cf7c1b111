#!/bin/bash
# pipeline tests (short timeouts) + graph bench
mkdir -p gpurun_out
T=r02g
timeout -k 10 600 python -m pytest tests/test_gpu_pipeline.py -q -m gpu --timeout 150 -x > gpurun_out/${T}_pytest_pipeline.log 2>&1
echo "pipeline tests rc=$?"; tail -30 gpurun_out/${T}_pytest_pipeline.log | cut -c1-300
timeout -k 10 700 python tools/graph_bench.py --frames 1000 --out gpurun_out/${T}_graph_bench.json > gpurun_out/${T}_graph_bench.txt 2>&1
echo "graph rc=$?"; cat gpurun_out/${T}_graph_bench.txt | tail -30 | cut -c1-600
