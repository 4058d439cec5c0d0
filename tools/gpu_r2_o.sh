#!/bin/bash
# what the driver runs at round end on one GPU: all GPU tests (-x), smoke(), both bench arms
set +e
mkdir -p gpurun_out
echo "== pytest -m gpu -x"; timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2o_pytest.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/r2o_pytest.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
echo "== bench reference"; timeout 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 2>&1 | cut -c1-400
echo "== bench ours"; timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 2>&1 | cut -c1-2500
