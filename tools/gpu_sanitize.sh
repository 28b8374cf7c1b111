#!/bin/bash
# compute-sanitizer over the paths added in round 1e (position epilogue, chained launches, kernel-written result mirror)
mkdir -p gpurun_out
SEL='tests/test_posfilt.py tests/test_gpu_tracker.py -m gpu -x -q -k "gpu_kalman or gpu_mean or two_colour or tracker_epilogue or run_clip or chained_launches_across or device_resident_and_async"'
for tool in memcheck synccheck; do
  eval timeout 900 compute-sanitizer --tool $tool --error-exitcode 3 python -m pytest $SEL > gpurun_out/sanitize_$tool.log 2>&1
  echo "== $tool rc=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/sanitize_$tool.log | tail -3
done
eval timeout 900 compute-sanitizer --tool racecheck --error-exitcode 3 python -m pytest tests/test_posfilt.py tests/test_gpu_tracker.py -m gpu -x -q -k '"tracker_epilogue or run_clip or device_resident_and_async"' > gpurun_out/sanitize_racecheck.log 2>&1
echo "== racecheck rc=$?"; grep -E "RACECHECK SUMMARY|passed|failed" gpurun_out/sanitize_racecheck.log | tail -3; grep -E "hazard" gpurun_out/sanitize_racecheck.log | sed 's/0x[0-9a-f]*/ADDR/g' | sort | uniq -c | sort -rn | head -8
