#!/bin/bash
# usage: gpu_ncu_k2.sh <kernel regex> <tag> [kbench args...]
mkdir -p gpurun_out
K=$1; TAG=$2; shift; shift
timeout 600 ncu --set full --clock-control none --import-source on -k regex:$K -s 40 -c 1 -f -o gpurun_out/${TAG} python tools/kbench.py --steps 20 "$@" > gpurun_out/ncu_${TAG}.log 2>&1
echo "ncu rc=$?"
