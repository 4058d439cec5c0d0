#!/bin/bash
# compute-sanitizer over tools/san_driver.py only (every kernel family through the C ABI)
set +e
mkdir -p gpurun_out
: > gpurun_out/sanitizer_summary2.txt
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 python tools/san_driver.py > gpurun_out/san2_${tool}.log 2>&1
  rc=$?
  echo "$tool san_driver (scan incl. blocked / fallback / RTS variants, wide + wide fallback, sort incl. key maps + fallback look-back, occupancy variant, L2 policies, host pipeline, multi-GPU host entry) rc=$rc $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/san2_${tool}.log | tail -1) done=$(grep -c 'san_driver done' gpurun_out/san2_${tool}.log)" | tee -a gpurun_out/sanitizer_summary2.txt
  grep -E "Invalid|Race reported|Barrier error|Error" gpurun_out/san2_${tool}.log | head -5
done
