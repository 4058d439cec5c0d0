#!/bin/bash
# ticket-free Decoupled Fallback as GPS_ALGO_CSDL's mid-size kernel: full GPU suite, A/B through the library
set +e
mkdir -p gpurun_out
echo "== pytest -m gpu -x"; timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2u_pytest.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/r2u_pytest.log
echo "== A/B through the library"; timeout 600 python tools/ticket_free_ab.py 2>&1 | tee gpurun_out/r2u_ticket_free_ab.jsonl | cut -c1-220
