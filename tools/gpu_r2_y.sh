#!/bin/bash
# after the ticket-free dispatch: sharded tests on 2 real GPUs + the N = 2 bench (weak, 2^33 strong, both end-to-end legs)
set +e
mkdir -p gpurun_out
echo "== sharded tests on 2 GPUs"; timeout 400 python -m pytest tests/test_sharded_gpu.py -q -x > gpurun_out/r2y_pytest_sharded.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/r2y_pytest_sharded.log
echo "== bench N=2"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 100 --warmup 5 > gpurun_out/r2y_bench_2gpu.json 2> gpurun_out/r2y_bench_2gpu.err; echo "rc=$?"; cut -c1-4000 gpurun_out/r2y_bench_2gpu.json; grep -v "^\*\|OMP_NUM\|^$" gpurun_out/r2y_bench_2gpu.err | tail -5
