#!/bin/bash
# final: compute-sanitizer over the driver (now incl. the warp-level rescue), then what the driver runs at round end
set +e
mkdir -p gpurun_out
: > gpurun_out/r02_sanitizer_rescue.txt
for tool in memcheck racecheck synccheck; do
  timeout 400 compute-sanitizer --tool $tool --error-exitcode 9 python tools/san_driver.py > gpurun_out/san5_${tool}.log 2>&1
  rc=$?
  echo "$tool san_driver (everything above + the warp-level rescue of the one-wave kernels: late CTA on the tiny / small / mid register tile, uint32 and float32) rc=$rc $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/san5_${tool}.log | tail -1) done=$(grep -c 'san_driver done' gpurun_out/san5_${tool}.log)" | tee -a gpurun_out/r02_sanitizer_rescue.txt
  grep -E "Invalid|Race reported|Barrier error|Error" gpurun_out/san5_${tool}.log | head -5
done
bash tools/gpu_r2_o.sh
