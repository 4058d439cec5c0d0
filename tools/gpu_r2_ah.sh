#!/bin/bash
set +e
mkdir -p gpurun_out
timeout 900 python tools/size_sweep_lib.py 30 2>&1 | tee gpurun_out/r02_size_sweep_lib.jsonl | cut -c1-260
