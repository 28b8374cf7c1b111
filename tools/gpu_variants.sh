#!/bin/bash
for lib in oat_b200/liboatgpu.so oat_b200/liboatgpu_t128_m4.so oat_b200/liboatgpu_t128_m5.so; do
  echo "== $lib"
  OAT_B200_LIB=$PWD/$lib python tools/kbench.py --res 1080p --steps 200 | tail -1
  OAT_B200_LIB=$PWD/$lib python tools/kbench.py --res 1080p --steps 200 --static | tail -1
  OAT_B200_LIB=$PWD/$lib python tools/kbench.py --res 4k --steps 100 | tail -1
  OAT_B200_LIB=$PWD/$lib python tools/hostcost.py
done
