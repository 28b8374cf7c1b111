#!/bin/bash
mkdir -p gpurun_out
T=${1:-r02r}
for v in "" _s48 _s74; do
  export OAT_B200_LIB=$PWD/oat_b200/liboatgpu$v.so
  echo "=== variant [$v]"
  python tools/tail_probe.py --blobs 1 8 60 2>&1 | cut -c1-200
  timeout -k 10 300 python bench.py --steps 1000 --no-cpu-baseline --no-extras > gpurun_out/${T}_bench$v.json 2> gpurun_out/${T}_bench$v.err
  python - <<P
import json
d = json.loads(open("gpurun_out/${T}_bench$v.json").read().splitlines()[-1])
print("bench value", d["value"], "roofline", d["roofline"]["frac"], "kernel ms/frame", d["roofline"]["ms_per_frame"])
P
done
