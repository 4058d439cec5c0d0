#!/bin/bash
set +e
timeout 600 python -m pytest tests/test_parity_gpu.py -x -q -k "ticket_free" 2>&1 | tail -4
