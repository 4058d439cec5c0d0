#!/bin/bash
# compute-sanitizer on the tuning harness at small sizes: memcheck, racecheck, synccheck
set +e
mkdir -p gpurun_out
T=gpuprefixsums_b200/bin/gps_tune
for tool in memcheck racecheck synccheck; do
  for filt in "csdl_reg<128,12>" "csdl_tma<512" "rts<256,8>" "csdl_fallback<128,12>/mask7"; do
    echo "== $tool $filt"
    timeout 600 compute-sanitizer --tool $tool --error-exitcode 9 $T --n 200003 --iters 1 --filter "$filt" > gpurun_out/san_${tool}.tmp 2>&1
    rc=$?
    echo "rc=$rc $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/san_${tool}.tmp | tail -1)"
    cat gpurun_out/san_${tool}.tmp >> gpurun_out/sanitizer_${tool}.log
  done
done
