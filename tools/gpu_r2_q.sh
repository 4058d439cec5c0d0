#!/bin/bash
# after the 64-bit sort: sort tests with the final tile shapes, compute-sanitizer over the driver (now incl. gps_sort_keys64),
# full ncu capture of one 64-bit digit pass
set +e
mkdir -p gpurun_out
echo "== sort tests"; timeout 600 python -m pytest tests/test_sort_gpu.py -q -x 2>&1 | tail -3
: > gpurun_out/r02_sanitizer_sort64.txt
for tool in memcheck racecheck synccheck; do
  timeout 300 compute-sanitizer --tool $tool --error-exitcode 9 python tools/san_driver.py > gpurun_out/san3_${tool}.log 2>&1
  rc=$?
  echo "$tool san_driver (all of round 2 + gps_sort_keys64 keys / pairs / key maps / 45-bit / Decoupled Fallback with stalls) rc=$rc $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/san3_${tool}.log | tail -1) done=$(grep -c 'san_driver done' gpurun_out/san3_${tool}.log)" | tee -a gpurun_out/r02_sanitizer_sort64.txt
  grep -E "Invalid|Race reported|Barrier error|Error" gpurun_out/san3_${tool}.log | head -5
done
echo "== ncu sort64 pass"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:sort_pass_kernel -s 10 -c 1 -f -o gpurun_out/r02_prof_sort64_pass python tools/sort64_bench.py 27 2 > gpurun_out/r02_ncu_sort64.log 2>&1; echo "rc=$?"
python tools/ncu_summary.py gpurun_out/r02_ncu_sort64.json gpurun_out/r02_prof_sort64_pass.ncu-rep | head -50
echo "== sort64 bench"; timeout 300 python tools/sort64_bench.py 27 10 2>&1 | tee gpurun_out/r2q_sort64_bench.json | tail -1
