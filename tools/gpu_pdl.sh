#!/bin/bash
# programmatic dependent launch: parity test, size sweep with and without it, bench.py with it
set +e
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_parity_gpu.py -x -q -k "dependent_launch" 2>&1 | tail -2
T=gpuprefixsums_b200/bin/gps_tune
timeout 600 $T --log2n 28 --sweep > gpurun_out/sweep_nopdl.jsonl 2>/dev/null
timeout 600 $T --log2n 28 --sweep --pdl > gpurun_out/sweep_pdl.jsonl 2>/dev/null
python - <<'PY'
import json, math
def load(f):
    by = {}
    for l in open(f):
        if not l.startswith("{"): continue
        r = json.loads(l)
        if "variant" in r: by[(int(round(math.log2(r["n"]))), r["variant"].replace("sweep/", ""))] = r
    return by
a, b = load("gpurun_out/sweep_nopdl.jsonl"), load("gpurun_out/sweep_pdl.jsonl")
for v in ["csdl_reg<128,4>", "csdl_reg<256,4>", "csdl_smem<128,8>", "csdl_blocked<256,9>"]:
    print(v, " ".join("2^%d:%.2f->%.2f(e%d)" % (lg, a[(lg, v)]["ms"] * 1e3, b[(lg, v)]["ms"] * 1e3, b[(lg, v)]["errors"]) for lg in range(10, 29, 2) if (lg, v) in a and (lg, v) in b))
PY
for o in 0 1; do timeout 300 python bench.py --no-cpu-baseline --e2e-steps 5 --option pdl=$o 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print('pdl=$o', round(d['value'], 1), round(d['e2e']['value'], 2))"; done
