#!/bin/bash
# quick iteration: GPU parity tests + kernel timings
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
for a in 0.01 0; do
timeout 300 python tools/kbench.py --res 1080p --alpha $a --steps 300 2>&1 | tail -1
done
timeout 300 python tools/kbench.py --res 1080p --alpha 0.01 --steps 300 --noflush 2>&1 | tail -1
timeout 300 python tools/kbench.py --res 4k --alpha 0.01 --steps 100 2>&1 | tail -1
