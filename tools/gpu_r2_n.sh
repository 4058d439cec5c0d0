#!/bin/bash
set +e
mkdir -p gpurun_out
bash tools/gpu_sanitize2.sh
cp gpurun_out/sanitizer_summary2.txt gpurun_out/r02_sanitizer_summary.txt
echo "== wide tests"; timeout 600 python -m pytest tests/test_wide_gpu.py -q -x 2>&1 | tail -3
