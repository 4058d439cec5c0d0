#!/bin/bash
# register tiles whose one wave covers 2^22 keys: <128,8> (1332 tiles of 4096 keys), <128,6> (1480 tiles of 3072 keys)
set +e
mkdir -p gpurun_out
T=gpuprefixsums_b200/bin/gps_tune
: > gpurun_out/r2ab_reg128.jsonl
for n in 1048576 2097152 3000000 3800000 4194304 4500000 5000000; do
  echo "## n=$n"
  $T --n $n --iters 400 --filter exp --pdl 2>/dev/null | grep "reg<256,4>/w1\|reg<128,8>/w1\|reg<128,6>/w1\|nostats/flags16" | tee -a gpurun_out/r2ab_reg128.jsonl | python -c "
import sys, json
for l in sys.stdin:
    r = json.loads(l)
    if 'ms' in r: print('  %-44s %8.2f us  errs %s' % (r['variant'], r['ms'] * 1e3, r.get('errors')))
"
done
