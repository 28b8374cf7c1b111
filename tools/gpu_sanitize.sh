#!/bin/bash
# compute-sanitizer over the round-2 paths: the resident fused kernel + tail server (clip engine and streaming use),
# frame-to-frame tile hand-off inside one launch and across chained launches, staged frames, the repitch kernel.
# usage: gpu_sanitize.sh <tag>   (under gpurun, 1 GPU)
T=${1:-r02}
mkdir -p gpurun_out
SMALL='tests/test_gpu_resident.py -m gpu -x -q -k "(equals_frame_by_frame and 120-160 and not 64) or ragged or staged_frames or mixed_parameters or interleaved or many_blobs"'
for tool in memcheck synccheck; do
  eval timeout 900 compute-sanitizer --tool $tool --error-exitcode 3 python -m pytest $SMALL -p no:cacheprovider > gpurun_out/${T}_sanitize_$tool.log 2>&1
  echo "== $tool rc=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/${T}_sanitize_$tool.log | tail -n 3
done
# racecheck (shared-memory hazards) on small frames AND on a 1080p chained clip (OAT_SOAK_FRAMES keeps it short)
eval timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 3 python -m pytest tests/test_gpu_resident.py -m gpu -x -q -k '"close_dependencies or (equals_frame_by_frame and 120-160)"' -p no:cacheprovider > gpurun_out/${T}_sanitize_racecheck.log 2>&1
echo "== racecheck small rc=$?"; grep -E "RACECHECK SUMMARY|passed|failed" gpurun_out/${T}_sanitize_racecheck.log | tail -n 3
OAT_SOAK_FRAMES=96 timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 3 python -m pytest tests/test_gpu_resident.py -m gpu -x -q -k soak_1080p -p no:cacheprovider > gpurun_out/${T}_sanitize_racecheck_1080p.log 2>&1
echo "== racecheck 1080p rc=$?"; grep -E "RACECHECK SUMMARY|passed|failed" gpurun_out/${T}_sanitize_racecheck_1080p.log | tail -n 3
grep -hE "hazard" gpurun_out/${T}_sanitize_racecheck*.log | sed 's/0x[0-9a-f]*/ADDR/g' | sort | uniq -c | sort -rn | head -n 12
