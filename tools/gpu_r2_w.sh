#!/bin/bash
# register tile <256,4> at 32 registers (8 CTAs per SM: 2^22 keys = 1024 tiles in one wave of 1184): ticket vs blockIdx order
set +e
mkdir -p gpurun_out
T=gpuprefixsums_b200/bin/gps_tune
: > gpurun_out/r2w_reg32.jsonl
for lg in 16 18 20 21 22; do
  echo "## 2^$lg"
  $T --log2n $lg --iters 400 --filter exp --pdl 2>/dev/null | grep "reg<256,4>/w1\|reg<256,4>r3\|nostats" | tee -a gpurun_out/r2w_reg32.jsonl | python -c "
import sys, json
for l in sys.stdin:
    r = json.loads(l)
    if 'ms' in r: print('  %-44s %8.2f us  errs %s' % (r['variant'], r['ms'] * 1e3, r.get('errors')))
"
done
for n in 3800000 4000000 4500000 4800000 5000000 6000000; do
  echo "## n=$n"
  $T --n $n --iters 400 --filter exp --pdl 2>/dev/null | grep "reg<256,4>/w1\|reg<256,4>r3\|nostats\|smem<128,8>/flags0\|smem<128,8>/flags16" | tee -a gpurun_out/r2w_reg32.jsonl | python -c "
import sys, json
for l in sys.stdin:
    r = json.loads(l)
    if 'ms' in r: print('  %-44s %8.2f us  errs %s' % (r['variant'], r['ms'] * 1e3, r.get('errors')))
"
done
