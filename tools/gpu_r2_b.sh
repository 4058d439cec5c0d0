#!/bin/bash
# round 2, call B (8 GPUs): box-wide PCIe probe, sharded tests on real NVLink peers, bench at N = 8
set +e
mkdir -p gpurun_out
nvidia-smi -L | wc -l
echo "== pcie probe"; timeout 400 python tools/pcie_aggregate.py --out gpurun_out/r02_pcie_aggregate.json 2>&1 | tail -40
echo "== sharded tests on 8 GPUs"; timeout 400 python -m pytest tests/test_sharded_gpu.py -q -x > gpurun_out/r2b_pytest_sharded.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/r2b_pytest_sharded.log
echo "== bench N=8"; timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 100 --warmup 5 > gpurun_out/r2b_bench_8gpu.log 2> gpurun_out/r2b_bench_8gpu.err; echo "rc=$?"; cut -c1-4000 gpurun_out/r2b_bench_8gpu.log; tail -5 gpurun_out/r2b_bench_8gpu.err
