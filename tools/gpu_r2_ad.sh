#!/bin/bash
set +e
mkdir -p gpurun_out
timeout 600 python tools/ticket_free_ab.py --in-place 2>&1 | tee gpurun_out/r2ad_ticket_free_inplace.jsonl | python -c "
import sys, json, collections
agg = collections.OrderedDict()
for l in sys.stdin:
    if not l.startswith('{'): print(l, end=''); continue
    r = json.loads(l); agg.setdefault((r['n'], r['ticket_free']), []).append(r['us_graph_back_to_back'])
ns = sorted({k[0] for k in agg})
for n in ns:
    a = sum(agg[(n, 0)]) / len(agg[(n, 0)]); b = sum(agg[(n, 1)]) / len(agg[(n, 1)])
    print('%10d  ticket %8.2f  ticket-free %8.2f  gain %5.1f %%' % (n, a, b, 100 * (a - b) / a))
"
