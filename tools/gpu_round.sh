#!/bin/bash
# Round check on one B200: every GPU test, smoke, the size sweep (BASELINE configs[3]), both
# bench arms, and ncu captures of the wide-element kernels.
set +e
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"; tail -6 gpurun_out/pytest_gpu.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/smoke.log
echo "== gps_main --ext"; timeout 600 gpuprefixsums_b200/bin/gps_main --ext > gpurun_out/gps_main_ext.log 2>&1; echo "rc=$?"; grep -E "passed|failed|speed" gpurun_out/gps_main_ext.log
echo "== size sweep"; timeout 900 gpuprefixsums_b200/bin/gps_tune --log2n 30 --sweep > gpurun_out/sweep_sizes2.jsonl 2> gpurun_out/sweep_sizes2.err; echo "rc=$?"; tail -2 gpurun_out/sweep_sizes2.err
echo "== bench"; timeout 900 python bench.py > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "rc=$?"; cat gpurun_out/bench.log | cut -c1-1500; tail -3 gpurun_out/bench.err
echo "== bench reference arm"; timeout 900 python bench.py --impl reference > gpurun_out/bench_ref.log 2> gpurun_out/bench_ref.err; echo "rc=$?"; cat gpurun_out/bench_ref.log | cut -c1-1500; tail -3 gpurun_out/bench_ref.err
echo "== ncu wide"; timeout 600 ncu --set full --import-source on --clock-control none -k regex:wide_scan_kernel -s 6 -c 1 -f -o gpurun_out/prof_wide_u64 python tools/wide_bench.py 2 > gpurun_out/ncu_wide.log 2>&1; tail -2 gpurun_out/ncu_wide.log
