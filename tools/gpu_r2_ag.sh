#!/bin/bash
set +e
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_parity_gpu.py tests/test_fuzz_gpu.py -x -q 2>&1 | tail -4
