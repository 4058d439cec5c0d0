#!/bin/bash
# Full GPU pass: golden fixtures, parity tests, reference + CUB baselines, C++ driver, bench.
set +e
mkdir -p gpurun_out
echo "== golden"; timeout 300 python oracle/gen_golden.py gpurun_out/golden > gpurun_out/golden.log 2>&1; echo "rc=$?"
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"
tail -5 gpurun_out/pytest_gpu.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/smoke.log
echo "== reference kernels"; timeout 300 python tools/run_ref_kernel.py 28 50 3 > gpurun_out/ref_kernels.json 2> gpurun_out/ref_kernels.err; echo "rc=$?"; cat gpurun_out/ref_kernels.json
echo "== gps_main (full reference protocol)"; timeout 900 gpuprefixsums_b200/bin/gps_main > gpurun_out/gps_main.log 2>&1; echo "rc=$?"
grep -E "passed|failed|Failed|speed|stats" gpurun_out/gps_main.log
echo "== bench"; timeout 900 python bench.py > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "rc=$?"
cat gpurun_out/bench.log; tail -3 gpurun_out/bench.err
echo "== bench reference arm"; timeout 900 python bench.py --impl reference > gpurun_out/bench_ref.log 2> gpurun_out/bench_ref.err; echo "rc=$?"
cat gpurun_out/bench_ref.log; tail -3 gpurun_out/bench_ref.err
