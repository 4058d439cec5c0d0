#!/bin/bash
# sustained bench.py runs (2000 launches) under different look-back back-off / tile-order options
for o in "poll_sleep_ns=0" "poll_sleep_ns=50" "poll_sleep_ns=200" "poll_sleep_ns=1000" "tile_order=1"; do
  timeout 200 python bench.py --steps 2000 --no-cpu-baseline --e2e-steps 1 --option $o 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$o', round(d['value'], 1), d['clocks'])"
done
