#!/bin/bash
# (1) blocked tile: ticket vs blockIdx order, plain and Decoupled Fallback, 2^26..2^29; (2) library A/B with the ticket-free smem range extended to 2^27
set +e
mkdir -p gpurun_out
T=gpuprefixsums_b200/bin/gps_tune
: > gpurun_out/r2af_blocked.jsonl
for lg in 26 27 28 29; do
  echo "## 2^$lg"
  $T --log2n $lg --iters 200 --filter exp --pdl 2>/dev/null | grep "blocked\|smem<128,8>/flags0\|nostats/flags16" | tee -a gpurun_out/r2af_blocked.jsonl | python -c "
import sys, json
for l in sys.stdin:
    r = json.loads(l)
    if 'ms' in r: print('  %-44s %8.2f us  errs %s' % (r['variant'], r['ms'] * 1e3, r.get('errors')))
"
done
echo "== library A/B, smem ticket-free range up to 2^27"
GPS_B200_LIB=$PWD/gpuprefixsums_b200/lib_alt/libgps_tf27.so timeout 600 python tools/ticket_free_ab.py big 2>&1 | tee gpurun_out/r2af_tf27_ab.jsonl | python -c "
import sys, json, collections
agg = collections.OrderedDict()
for l in sys.stdin:
    if not l.startswith('{'): print(l, end=''); continue
    r = json.loads(l); agg.setdefault((r['n'], r['ticket_free']), []).append((r['us_graph_back_to_back'], r['us_cold_l2']))
for n in sorted({k[0] for k in agg}):
    a = agg[(n, 0)]; b = agg[(n, 1)]
    ma = sum(x[0] for x in a) / len(a); mb = sum(x[0] for x in b) / len(b)
    ca = sum(x[1] for x in a) / len(a); cb = sum(x[1] for x in b) / len(b)
    print('%10d  ticket %8.2f  ticket-free %8.2f  gain %5.1f %%   cold %8.2f -> %8.2f' % (n, ma, mb, 100 * (ma - mb) / ma, ca, cb))
"
