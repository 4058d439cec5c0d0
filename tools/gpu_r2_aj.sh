#!/bin/bash
# warp-level rescue in the one-wave ticket-free kernels: the late-CTA test, the parity + fuzz tests, and the library A/B again (cost of the extra code)
set +e
mkdir -p gpurun_out
echo "== rescue + ticket-free tests"; timeout 600 python -m pytest tests/test_parity_gpu.py -x -q -k "rescued or ticket_free" 2>&1 | tail -5
echo "== parity + fuzz"; timeout 1200 python -m pytest tests/test_parity_gpu.py tests/test_fuzz_gpu.py -x -q 2>&1 | tail -3
echo "== A/B"; timeout 600 python tools/ticket_free_ab.py three 2>&1 | tee gpurun_out/r2aj_ticket_free_ab.jsonl | python -c "
import sys, json, collections
agg = collections.OrderedDict()
for l in sys.stdin:
    if not l.startswith('{'): print(l, end=''); continue
    r = json.loads(l); agg.setdefault((r['n'], r['ticket_free']), []).append(r['us_graph_back_to_back'])
for n in sorted({k[0] for k in agg}):
    m = lambda t: sum(agg[(n, t)]) / len(agg[(n, t)])
    print('%10d  tickets (0) %8.2f   rescue-carrying (1) %8.2f   lean (2) %8.2f' % (n, m(0), m(1), m(2)))
"
