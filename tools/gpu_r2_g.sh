#!/bin/bash
set +e
mkdir -p gpurun_out
for mw in 1 2 4 8; do
  T=gpuprefixsums_b200/bin/gps_tune_mw$mw; [ $mw = 4 ] && T=gpuprefixsums_b200/bin/gps_tune
  $T --log2n 28 --sweep --pdl --min-log2n 21 2>/dev/null | grep "smem<128,8>\|blocked<256,9>\|blocked<128,9>" > gpurun_out/r2g_mw$mw.jsonl
done
python - <<'P'
import json
res={}
for mw in (1,2,4,8):
    for l in open(f"gpurun_out/r2g_mw{mw}.jsonl"):
        r=json.loads(l); res[(r["variant"].replace("sweep/",""), r["n"], mw)]=(r["ms"]*1e3, r["errors"])
vs=sorted({k[0] for k in res}); ns=sorted({k[1] for k in res})
for v in vs:
    print(v)
    for n in ns:
        print("  2^%d"%(n.bit_length()-1), "  ".join("mw%d %7.2f us%s"%(mw, res[(v,n,mw)][0], "" if res[(v,n,mw)][1]==0 else " ERR") for mw in (1,2,4,8)))
P
