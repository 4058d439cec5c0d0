#!/bin/bash
set +e
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/pytest_gpu.log
echo "== size sweep"; timeout 900 gpuprefixsums_b200/bin/gps_tune --log2n 30 --sweep > gpurun_out/sweep_sizes3.jsonl 2> gpurun_out/sweep_sizes3.err; echo "rc=$?"; tail -2 gpurun_out/sweep_sizes3.err
echo "== bench"; timeout 900 python bench.py > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "rc=$?"; cat gpurun_out/bench.log | cut -c1-1200; tail -3 gpurun_out/bench.err
echo "== bench 4000 steps"; timeout 900 python bench.py --steps 4000 --no-cpu-baseline --e2e-steps 2 > gpurun_out/bench_4000.log 2> gpurun_out/bench_4000.err; cat gpurun_out/bench_4000.log | cut -c1-400
