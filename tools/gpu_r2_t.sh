#!/bin/bash
# ticket-free Decoupled Fallback on the register tiles (small sizes) and the smem tile without stats counters
set +e
mkdir -p gpurun_out
T=gpuprefixsums_b200/bin/gps_tune
: > gpurun_out/r2t_static_fb_small.jsonl
for lg in 10 12 14 16 17 18 19 20 21 22 23 24 25; do
  echo "## 2^$lg"
  $T --log2n $lg --iters 400 --filter exp --pdl 2>/dev/null | grep "w1/flags\|reg_fb\|nostats" | grep -v "reg<64" | tee -a gpurun_out/r2t_static_fb_small.jsonl | python -c "
import sys, json
for l in sys.stdin:
    r = json.loads(l)
    if 'ms' in r: print('  %-44s %8.2f us  errs %s' % (r['variant'], r['ms'] * 1e3, r.get('errors')))
"
done
