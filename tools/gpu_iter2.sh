#!/bin/bash
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 300 python tools/kbench.py --res 1080p --alpha 0.01 --steps 200 2>&1 | tail -2
timeout 300 python bench.py --steps 1000 --warmup 50 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, 'pipelined',d['pipelined']['value'],'e2e',d['e2e']['value'],'roof',d['roofline']['achieved'],d['roofline']['kernel_ms'])"
OAT_B200_NO_OVERLAP=1 timeout 300 python bench.py --steps 1000 --warmup 50 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('NO_OVERLAP',{k:d[k] for k in ('value','ms_per_step','gpu_launches')}, 'pipelined',d['pipelined']['value'],'e2e',d['e2e']['value'])"
