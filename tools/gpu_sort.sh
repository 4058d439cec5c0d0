#!/bin/bash
# radix-sort consumer: parity tests, throughput of the default build and of every alternative
# build under gpuprefixsums_b200/lib_alt/, per-kernel device times of one default run
set +e
mkdir -p gpurun_out
echo "== pytest sort"; timeout 600 python -m pytest tests/test_sort_gpu.py -x -q > gpurun_out/pytest_sort.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/pytest_sort.log
echo "== bench sort"
: > gpurun_out/sort_bench.jsonl
timeout 300 python tools/sort_bench.py 28 10 >> gpurun_out/sort_bench.jsonl 2>gpurun_out/sort_bench.err
for l in gpuprefixsums_b200/lib_alt/libgps_s_*.so; do GPS_B200_LIB=$PWD/$l timeout 300 python tools/sort_bench.py 28 10 >> gpurun_out/sort_bench.jsonl 2>>gpurun_out/sort_bench.err; done
cat gpurun_out/sort_bench.jsonl; tail -5 gpurun_out/sort_bench.err
echo "== launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/sort_launches.csv python tools/sort_bench.py 28 1 > gpurun_out/sort_under_ncu.log 2>&1
grep -v "^==" gpurun_out/sort_launches.csv | awk -F'","' 'NR>1{print $5, $NF}' | sed 's/"//g' | cut -c1-150 | head -40
echo "== ncu full (default build, keys-only pass kernel)"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:sort_pass_kernel -s 5 -c 1 -o gpurun_out/prof_sort_pass -f python tools/sort_bench.py 28 1 > gpurun_out/ncu_sort.log 2>&1; tail -3 gpurun_out/ncu_sort.log
