#!/bin/bash
set +e
mkdir -p gpurun_out
echo "== wide tests"; timeout 900 python -m pytest tests/test_wide_gpu.py -q -x 2>&1 | tail -4
echo "== wide bench"; timeout 600 python tools/wide_bench.py 30 2>&1 | tee gpurun_out/r2i_wide_bench.json | tail -3
echo "== wide bench fb48"; GPS_B200_LIB=$PWD/gpuprefixsums_b200/lib_alt/libgps_fb48.so timeout 600 python tools/wide_bench.py 30 2>&1 | tail -3
