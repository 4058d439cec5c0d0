#!/bin/bash
set +e
run() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 2 --steps 300 --warmup 10 --device-only "$@" 2>/dev/null | grep device_only; }
run --option sharded_l2=0
run --option sharded_l2=1 --option sharded_l2_keep_mb=96
run --option sharded_l2=3 --option sharded_l2_keep_mb=96
run --option sharded_l2=7 --option sharded_l2_keep_mb=96
run --option sharded_l2=2
run --option sharded_l2=6
run --option sharded_l2=7 --option sharded_l2_keep_mb=64
run --option sharded_l2=7 --option sharded_l2_keep_mb=120
run --option sharded_l2=0
