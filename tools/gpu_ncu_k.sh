#!/bin/bash
# usage: gpu_ncu_k.sh <kernel regex> <tag> [res] [alpha]
mkdir -p gpurun_out
K=$1; TAG=$2; RES=${3:-1080p}; A=${4:-0.01}
timeout 600 ncu --set full --clock-control none --import-source on -k regex:$K -s 40 -c 1 -f -o gpurun_out/${TAG} python tools/kbench.py --res $RES --alpha $A --steps 20 > gpurun_out/ncu_${TAG}.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/ncu_${TAG}.log
