#!/bin/bash
set +e
mkdir -p gpurun_out
timeout 600 python tools/sharded_host_ab.py 2>&1 | tee gpurun_out/r2z_sharded_host_ab.jsonl
