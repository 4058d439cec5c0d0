#!/bin/bash
# round 2, call H (1 GPU): wide + sort Decoupled Fallback parity and throughput, full GPU test suite
set +e
mkdir -p gpurun_out
echo "== wide tests"; timeout 900 python -m pytest tests/test_wide_gpu.py -q -x 2>&1 | tail -15
echo "== sort tests"; timeout 900 python -m pytest tests/test_sort_gpu.py -q -x 2>&1 | tail -15
echo "== wide bench"; timeout 600 python tools/wide_bench.py 30 2>&1 | tee gpurun_out/r2h_wide_bench.json | tail -3
echo "== all gpu tests"; timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2h_pytest.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/r2h_pytest.log
