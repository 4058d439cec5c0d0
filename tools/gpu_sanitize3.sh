#!/bin/bash
# compute-sanitizer over tools/san_driver.py after the ticket-free dispatch (every kernel family through the C ABI)
set +e
mkdir -p gpurun_out
: > gpurun_out/r02_sanitizer_final.txt
for tool in memcheck racecheck synccheck; do
  timeout 400 compute-sanitizer --tool $tool --error-exitcode 9 python tools/san_driver.py > gpurun_out/san4_${tool}.log 2>&1
  rc=$?
  echo "$tool san_driver (all of round 2 + ticket-free launches: one-wave blockIdx order, Decoupled Fallback with tile = blockIdx.x incl. emulated stalls, in place, float) rc=$rc $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/san4_${tool}.log | tail -1) done=$(grep -c 'san_driver done' gpurun_out/san4_${tool}.log)" | tee -a gpurun_out/r02_sanitizer_final.txt
  grep -E "Invalid|Race reported|Barrier error|Error" gpurun_out/san4_${tool}.log | head -5
done
