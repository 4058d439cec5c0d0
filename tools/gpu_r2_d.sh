#!/bin/bash
# round 2, call D (1 GPU): mid-size tile-shape ladder (back to back with PDL, and cold L2), sharded tests
set +e
mkdir -p gpurun_out
echo "== sharded tests"; timeout 600 python -m pytest tests/test_sharded_gpu.py -q -x > gpurun_out/r2d_pytest_sharded.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/r2d_pytest_sharded.log
T=gpuprefixsums_b200/bin/gps_tune
echo "== sweep pdl"; timeout 600 $T --log2n 27 --sweep --pdl --min-log2n 20 > gpurun_out/r2d_sweep_pdl.jsonl 2> gpurun_out/r2d_sweep_pdl.err; echo "rc=$?"
echo "== sweep pdl flush"; timeout 600 $T --log2n 27 --sweep --pdl --flush --min-log2n 20 > gpurun_out/r2d_sweep_flush.jsonl 2> gpurun_out/r2d_sweep_flush.err; echo "rc=$?"
python - <<'P'
import json
for f in ("r2d_sweep_pdl","r2d_sweep_flush"):
    print("##", f)
    rows=[json.loads(l) for l in open(f"gpurun_out/{f}.jsonl") if l.startswith("{")]
    by={}
    for r in rows:
        if "error" in r: print(r); continue
        by.setdefault(r.get("n"),[]).append(r)
    for n,rs in sorted(by.items()):
        rs.sort(key=lambda r:r.get("ms_mean",r.get("ms",9e9)))
        print(n, [(r["variant"].replace("sweep/",""), round(1e3*r.get("ms_mean",r.get("ms",0)),2), r.get("errs")) for r in rs[:8]])
P
