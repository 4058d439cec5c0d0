#!/bin/bash
# 64-bit key sort: GPU tests of the sort (32- and 64-bit), throughput next to cub::DeviceRadixSort, tile-shape variants
set +e
mkdir -p gpurun_out
echo "== sort tests"; timeout 900 python -m pytest tests/test_sort_gpu.py -q -x 2>&1 | tail -6
echo "== sort64 bench default"; timeout 300 python tools/sort64_bench.py 27 10 2>&1 | tee gpurun_out/r2p_sort64_bench.jsonl | tail -2
for l in gpuprefixsums_b200/lib_alt/libgps_s64_*.so; do
  echo "== $l"; GPS_B200_LIB=$PWD/$l timeout 300 python tools/sort64_bench.py 27 10 2>&1 | tee -a gpurun_out/r2p_sort64_bench.jsonl | tail -1
done
echo "== sort32 bench (regression check of the refactored dispatch)"; timeout 300 python tools/sort_bench.py 28 10 2>&1 | tee gpurun_out/r2p_sort_bench.json | tail -1
