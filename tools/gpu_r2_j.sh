#!/bin/bash
# round 2, call J (N GPUs given as $1): bench at N (weak + 2^33 strong + both end-to-end legs)
set +e
N=${1:-2}
mkdir -p gpurun_out
echo "== bench N=$N"; timeout 800 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 100 --warmup 5 > gpurun_out/r2j_bench_${N}gpu.log 2> gpurun_out/r2j_bench_${N}gpu.err; echo "rc=$?"; cut -c1-5000 gpurun_out/r2j_bench_${N}gpu.log; grep -v "^\*\|OMP_NUM\|^$" gpurun_out/r2j_bench_${N}gpu.err | tail -8
