#!/bin/bash
# round 2, call M (8 GPUs): sharded tests on real NVLink peers, bench at N = 8 (weak, 2^33 strong, both e2e legs)
set +e
mkdir -p gpurun_out
echo "== sharded tests on 8 GPUs"; timeout 400 python -m pytest tests/test_sharded_gpu.py -q -x > gpurun_out/r2m_pytest_sharded.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/r2m_pytest_sharded.log
echo "== bench N=8"; timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 100 --warmup 5 > gpurun_out/r02_bench_8gpu.json 2> gpurun_out/r02_bench_8gpu.err; echo "rc=$?"; cut -c1-6000 gpurun_out/r02_bench_8gpu.json; grep -v "^\*\|OMP_NUM\|^$" gpurun_out/r02_bench_8gpu.err | tail -5
echo "== reference arm N=8"; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29534 bench.py --impl reference --gpus 8 --steps 20 --warmup 5 2>/dev/null | cut -c1-1500
