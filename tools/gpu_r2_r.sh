#!/bin/bash
# eager look-back on the smem tile (warp 0 walks the descriptors while the TMA load is in flight): mid sizes, A/B in one process
set +e
mkdir -p gpurun_out
T=gpuprefixsums_b200/bin/gps_tune
: > gpurun_out/r2r_eager.jsonl
for lg in 21 22 23 24 25 26 27 28; do
  echo "## 2^$lg"
  $T --log2n $lg --iters 300 --filter exp --pdl 2>/dev/null | grep "smem<128,8>" | tee -a gpurun_out/r2r_eager.jsonl | python -c "
import sys, json
for l in sys.stdin:
    r = json.loads(l)
    print('  %-44s %8.2f us  errs %s' % (r['variant'], r['ms'] * 1e3, r.get('errors')))
"
done
