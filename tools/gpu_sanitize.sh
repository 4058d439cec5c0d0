#!/bin/bash
# compute-sanitizer: (1) the tuning harness at small sizes (scan kernels incl. fallback with
# emulated stalls), (2) tools/san_driver.py = every kernel family through the C ABI (scan, wide,
# sort); memcheck, racecheck, synccheck
set +e
mkdir -p gpurun_out
T=gpuprefixsums_b200/bin/gps_tune
: > gpurun_out/sanitizer_summary.txt
for tool in memcheck racecheck synccheck; do
  : > gpurun_out/sanitizer_${tool}.log
  for filt in "csdl_reg<128,12>" "csdl_smem<128,8>/g4/r40" "csdl_tma<512" "rts<128,12>" "csdl_fallback<128,12>/mask7"; do
    timeout 600 compute-sanitizer --tool $tool --error-exitcode 9 $T --n 200003 --iters 1 --filter "$filt" > gpurun_out/san_${tool}.tmp 2>&1
    rc=$?
    echo "$tool gps_tune $filt rc=$rc $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/san_${tool}.tmp | tail -1)" | tee -a gpurun_out/sanitizer_summary.txt
    cat gpurun_out/san_${tool}.tmp >> gpurun_out/sanitizer_${tool}.log
  done
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 python tools/san_driver.py > gpurun_out/san_${tool}.tmp 2>&1
  rc=$?
  echo "$tool san_driver (scan, wide, sort via C ABI) rc=$rc $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/san_${tool}.tmp | tail -1) $(grep -c 'san_driver done' gpurun_out/san_${tool}.tmp)" | tee -a gpurun_out/sanitizer_summary.txt
  cat gpurun_out/san_${tool}.tmp >> gpurun_out/sanitizer_${tool}.log
done
grep -E "Race reported|Invalid|Barrier error" gpurun_out/sanitizer_*.log | sort | uniq -c | head -20
