#!/bin/bash
# last call of the round: every GPU test (-x) and smoke()
set +e
mkdir -p gpurun_out
echo "== pytest -m gpu -x"; timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2an_pytest.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/r2an_pytest.log
echo "== the slot-timeout test, 10 more times"; for i in 1 2 3 4 5 6 7 8 9 10; do timeout 120 python -m pytest tests/test_sharded_gpu.py -q -k carry_slot_timeout 2>&1 | tail -1; done
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
