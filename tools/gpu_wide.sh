#!/bin/bash
# wide-element scans: parity tests, then throughput of the default build and of every
# alternative build under gpuprefixsums_b200/lib_alt/
set +e
mkdir -p gpurun_out
echo "== pytest wide"; timeout 900 python -m pytest tests/test_wide_gpu.py -x -q > gpurun_out/pytest_wide.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/pytest_wide.log
echo "== bench wide"
: > gpurun_out/wide_bench.jsonl
timeout 300 python tools/wide_bench.py >> gpurun_out/wide_bench.jsonl 2>gpurun_out/wide_bench.err
for l in gpuprefixsums_b200/lib_alt/*.so; do GPS_B200_LIB=$PWD/$l timeout 300 python tools/wide_bench.py >> gpurun_out/wide_bench.jsonl 2>>gpurun_out/wide_bench.err; done
cat gpurun_out/wide_bench.jsonl; tail -5 gpurun_out/wide_bench.err
