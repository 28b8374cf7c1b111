#!/bin/bash
# Try an experimental library build (oat_b200/liboatgpu_exp.so): GPU tests + bench A/B against the current one;
# adopt it (on this box) when it is green and not slower, then capture the round's evidence.
TAG=${1:-r01}
EXP=$PWD/oat_b200/liboatgpu_exp.so
val() { python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['roofline']['frac'], d['cold_frame']['latency_ms'])"; }
ADOPT=0
if [ -f "$EXP" ]; then
  OAT_B200_LIB=$EXP timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/${TAG}_exp_pytest.log
  if grep -q " passed" gpurun_out/${TAG}_exp_pytest.log && ! grep -q "failed\|error" gpurun_out/${TAG}_exp_pytest.log; then
    for rep in 1 2; do
      M=$(timeout 300 python bench.py --no-cpu-baseline 2>/dev/null | val); echo "main 1080p: $M"
      E=$(OAT_B200_LIB=$EXP timeout 300 python bench.py --no-cpu-baseline 2>/dev/null | val); echo "exp  1080p: $E"
    done
    M4=$(timeout 300 python bench.py --no-cpu-baseline --workload 4k --steps 600 2>/dev/null | val); echo "main 4k: $M4"
    E4=$(OAT_B200_LIB=$EXP timeout 300 python bench.py --no-cpu-baseline --workload 4k --steps 600 2>/dev/null | val); echo "exp  4k: $E4"
    ADOPT=$(python -c "
m=float('$M'.split()[0]); e=float('$E'.split()[0]); m4=float('$M4'.split()[0]); e4=float('$E4'.split()[0])
print(1 if (e >= 1.005*m and e4 >= 0.995*m4) else 0)")
  fi
fi
echo "ADOPT_EXP=$ADOPT"
if [ "$ADOPT" = "1" ]; then cp -f $EXP oat_b200/liboatgpu.so; fi
tools/gpu_profiles.sh $TAG 2>&1 | tail -40
