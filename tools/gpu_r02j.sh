#!/bin/bash
mkdir -p gpurun_out
T=${1:-r02j}
timeout -k 10 600 python -m pytest tests/test_gpu_resident.py -q -m gpu --timeout 200 -x -k stream > gpurun_out/${T}_pytest_stream.log 2>&1
echo "stream tests rc=$?"; tail -n 5 gpurun_out/${T}_pytest_stream.log | cut -c1-300
timeout -k 10 400 python -m pytest tests/test_gpu_pipeline.py -q -m gpu --timeout 150 -x > gpurun_out/${T}_pytest_pipeline.log 2>&1
echo "pipeline tests rc=$?"; tail -n 5 gpurun_out/${T}_pytest_pipeline.log | cut -c1-300
timeout -k 10 600 python tools/graph_bench.py --frames 1000 --out gpurun_out/${T}_graph_bench.json > gpurun_out/${T}_graph_bench.txt 2>&1
echo "graph rc=$?"; cat gpurun_out/${T}_graph_bench.txt | tail -n 70 | cut -c1-330
