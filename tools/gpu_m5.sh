#!/bin/bash
export OAT_B200_LIB=$PWD/oat_b200/liboatgpu_exp5.so
run() { timeout 120 python bench.py --no-cpu-baseline --steps 1000 2>&1 | tail -1 | cut -c1-140; }
echo "== chain on";  for i in 1 2 3; do run; done
echo "== NO_CHAIN";  for i in 1 2 3; do OAT_B200_NO_CHAIN=1 run; done
echo "== NO_MIRROR"; for i in 1 2; do OAT_B200_NO_MIRROR=1 run; done
echo "== memcheck (chain on, 300 steps)"
timeout 200 compute-sanitizer --tool memcheck --error-exitcode 3 python bench.py --no-cpu-baseline --steps 300 --warmup 10 2>&1 | grep -E "=========|Error|error" | head -20
