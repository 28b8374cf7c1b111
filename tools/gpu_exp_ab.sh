#!/bin/bash
# GPU tests on the experimental build, then bench A/B (main vs exp) for the live and the frozen model
EXP=${EXP:-$PWD/oat_b200/liboatgpu_exp.so}
val() { python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value']), round(d['roofline']['frac'],3), round(d['cold_frame']['latency_ms']*1e3,1), d['multi_stream'] and round(d['multi_stream']['value']))"; }
OAT_B200_LIB=$EXP timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for args in "" "--workload 4k --steps 600" "--streams 8 --steps 1000"; do
  echo "main [$args]: $(timeout 300 python bench.py --no-cpu-baseline $args 2>/dev/null | val)"
  echo "exp  [$args]: $(OAT_B200_LIB=$EXP timeout 300 python bench.py --no-cpu-baseline $args 2>/dev/null | val)"
done
