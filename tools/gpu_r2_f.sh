#!/bin/bash
set +e
T=gpuprefixsums_b200/bin/gps_tune
for lg in 22 24 26; do
  echo "## 2^$lg"
  $T --log2n $lg --iters 200 --filter exp --pdl --prof 2>/dev/null | grep "smem<128,8>" | cut -c1-600
  $T --log2n $lg --iters 200 --filter copy --pdl 2>/dev/null | cut -c1-300
done
