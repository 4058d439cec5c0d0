#!/bin/bash
# round 2, call A (1 GPU): all GPU tests, both bench arms, the PCIe probe on one GPU
set +e
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2a_pytest.log 2>&1; echo "rc=$?"; tail -30 gpurun_out/r2a_pytest.log
echo "== bench ours"; timeout 600 python bench.py --steps 100 > gpurun_out/r2a_bench.log 2> gpurun_out/r2a_bench.err; echo "rc=$?"; cut -c1-3000 gpurun_out/r2a_bench.log; tail -5 gpurun_out/r2a_bench.err
echo "== bench reference"; timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2a_bench_ref.log 2> gpurun_out/r2a_bench_ref.err; echo "rc=$?"; cut -c1-2000 gpurun_out/r2a_bench_ref.log; tail -3 gpurun_out/r2a_bench_ref.err
echo "== pcie probe"; timeout 300 python tools/pcie_aggregate.py --out gpurun_out/r2a_pcie_1gpu.json 2>&1 | tail -20
