#!/bin/bash
summ() { python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$1', 'value',round(d['value']),'us/step',round(d['ms_per_step']*1e3,2),'e2e',round(d['e2e']['value']),'roof',round(d['roofline']['frac'],3),'cold',round(d['cold_frame']['latency_ms']*1e3,1), 'multi', d['multi_stream'] and round(d['multi_stream']['value']))"; }
for dep in 4 8 16; do
timeout 300 python bench.py --no-cpu-baseline --depth $dep 2>/dev/null | summ "1080p depth $dep"
done
timeout 300 python bench.py --no-cpu-baseline --depth 16 --workload 4k --steps 600 2>/dev/null | summ "4k depth 16"
timeout 300 python bench.py --no-cpu-baseline --streams 8 --steps 1000 2>/dev/null | summ "1080p 8 streams"
OAT_B200_NO_MIRROR=1 timeout 300 python bench.py --no-cpu-baseline 2>/dev/null | summ "1080p no-mirror"
