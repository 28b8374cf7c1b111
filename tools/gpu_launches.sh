#!/bin/bash
# per-kernel durations (ncu, serialized) of a few steady-state frames
mkdir -p gpurun_out
RES=${1:-1080p}; A=${2:-0.01}; TAG=${3:-x}
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 150 -c 120 --csv --log-file gpurun_out/launches_${TAG}.csv python tools/kbench.py --res $RES --alpha $A --steps 60 > gpurun_out/launches_${TAG}.log 2>&1
echo rc=$?
