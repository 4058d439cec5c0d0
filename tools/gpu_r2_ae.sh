#!/bin/bash
# persistent one-wave smem kernel (static strided tiles) against the ticketed / blockIdx / fallback variants, 2^22..2^28
set +e
mkdir -p gpurun_out
T=gpuprefixsums_b200/bin/gps_tune
: > gpurun_out/r2ae_persistent.jsonl
for lg in 22 23 24 25 26 27 28; do
  echo "## 2^$lg"
  $T --log2n $lg --iters 300 --filter exp --pdl 2>/dev/null | grep "smem_persistent\|smem<128,8>/flags0\|smem<128,8>/flags16\|eager8192/sleep0\|nostats/flags16" | tee -a gpurun_out/r2ae_persistent.jsonl | python -c "
import sys, json
for l in sys.stdin:
    r = json.loads(l)
    if 'ms' in r: print('  %-44s %8.2f us  errs %s' % (r['variant'], r['ms'] * 1e3, r.get('errors')))
"
done
