#!/bin/bash
# A/B of an environment switch on the judged bench line: gpu_ab.sh VAR [bench args...]
VAR=$1; shift
summ() { python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$1', 'value',round(d['value']),'ms/step',round(d['ms_per_step']*1e3,2),'us e2e',round(d['e2e']['value']),'roof',round(d['roofline']['achieved']),round(d['roofline']['frac'],3),'kernel_us',round(d['roofline']['kernel_ms']*1e3,2),'single',round(d['roofline']['single_launch_event_bracket_ms']*1e3,2),'cold',round(d['cold_frame']['latency_ms']*1e3,1))"; }
for rep in 1 2; do
timeout 300 python bench.py --no-cpu-baseline "$@" 2>/dev/null | summ "on "
env $VAR=1 timeout 300 python bench.py --no-cpu-baseline "$@" 2>/dev/null | summ "$VAR"
done
