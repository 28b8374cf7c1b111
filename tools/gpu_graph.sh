#!/bin/bash
# the reference's perf protocol through the drop-in components (tools/graph_bench.py).  usage: gpu_graph.sh <tag>
mkdir -p gpurun_out
T=${1:-r02}
timeout -k 10 900 python tools/graph_bench.py --frames 1000 --out gpurun_out/${T}_graph_bench.json > gpurun_out/${T}_graph_bench.txt 2>&1
echo "graph rc=$?"; grep -vE "mogfilt\[|colorconvert\[|tracker\[" gpurun_out/${T}_graph_bench.txt | cut -c1-400
