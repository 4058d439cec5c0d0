#!/bin/bash
# single-wave blockIdx order for the small register tiles + ticket-free fallback at mid sizes: parity tests, A/B through the library
set +e
mkdir -p gpurun_out
echo "== parity + fuzz + sharded + cpp driver"; timeout 1500 python -m pytest tests/test_parity_gpu.py tests/test_fuzz_gpu.py tests/test_sharded_gpu.py tests/test_cpp_driver_gpu.py -x -q > gpurun_out/r2v_pytest.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/r2v_pytest.log
echo "== A/B through the library"; timeout 600 python tools/ticket_free_ab.py 2>&1 | tee gpurun_out/r2v_ticket_free_ab.jsonl | cut -c1-200
