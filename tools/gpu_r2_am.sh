#!/bin/bash
set +e
timeout 900 python -m pytest tests/test_sharded_gpu.py -x -q 2>&1 | tail -4
