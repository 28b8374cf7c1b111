#!/bin/bash
# compute-sanitizer over what the last part of round 2 added: band pre-labelling + the labelling CTA's pre-labelled mode
# (forced on small frames), vectorised band morphology, two storer lanes in the fused kernel.
# usage: gpu_sanitize2.sh <tag>   (under gpurun, 1 GPU)
T=${1:-r02y}
mkdir -p gpurun_out
SEL='tests/test_gpu_resident.py -m gpu -x -q -k "(prelabelled and (shape0 or shape1)) or (equals_frame_by_frame and 120-160 and not 64) or (many_blobs and 8)"'
for tool in memcheck synccheck racecheck; do
  eval timeout 1200 compute-sanitizer --tool $tool --error-exitcode 3 python -m pytest $SEL -p no:cacheprovider > gpurun_out/${T}_sanitize_$tool.log 2>&1
  echo "== $tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" gpurun_out/${T}_sanitize_$tool.log | tail -n 3
done
grep -hE "hazard|Race reported|and (Read|Write) access" gpurun_out/${T}_sanitize_racecheck.log | sed 's/0x[0-9a-f]*/ADDR/g' | sort | uniq -c | sort -rn | head -n 16
