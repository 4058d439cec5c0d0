#!/bin/bash
set +e
mkdir -p gpurun_out
echo "== plain run of the sanitizer driver"; timeout 300 python tools/san_driver.py 2>&1 | tail -3
echo "== occupancy test"; timeout 600 python -m pytest tests/test_parity_gpu.py -q -x -k "occupancy" 2>&1 | tail -3
bash tools/gpu_sanitize2.sh
