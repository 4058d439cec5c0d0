#!/bin/bash
# Decoupled Fallback without tickets (tile = blockIdx.x), spin budgets, no emulated stalls, against plain CSDL with / without tickets
set +e
mkdir -p gpurun_out
T=gpuprefixsums_b200/bin/gps_tune
: > gpurun_out/r2s_static_fb.jsonl
for lg in 20 21 22 23 24 25 26 28; do
  echo "## 2^$lg"
  $T --log2n $lg --iters 300 --filter exp --pdl 2>/dev/null | grep "smem_fb<128,8>\|smem<128,8>/flags0\|smem<128,8>/flags16\|eager8192/sleep0" | tee -a gpurun_out/r2s_static_fb.jsonl | python -c "
import sys, json
for l in sys.stdin:
    r = json.loads(l)
    if 'ms' in r: print('  %-44s %8.2f us  errs %s' % (r['variant'], r['ms'] * 1e3, r.get('errors')))
    else: print('      stats', {k: v for k, v in r.items() if k != 'variant'})
"
done
