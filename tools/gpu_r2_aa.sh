#!/bin/bash
set +e
mkdir -p gpurun_out
echo "== gps_main --ext"; timeout 600 gpuprefixsums_b200/bin/gps_main --ext > gpurun_out/gps_main_ext_r2.log 2>&1; echo "rc=$?"; grep -v "^$" gpurun_out/gps_main_ext_r2.log | tail -30
echo "== cpp driver tests"; timeout 900 python -m pytest tests/test_cpp_driver_gpu.py tests/test_abi.py -q -x 2>&1 | tail -3
