#!/bin/bash
# round 2, call L (1 GPU): CUB-pinned sort fixtures; full size sweep 2^10..2^30 (back to back + cold L2)
set +e
mkdir -p gpurun_out
echo "== sort fixtures"; python oracle/gen_golden_sort.py gpurun_out/golden_sort 2>&1 | tail -10
T=gpuprefixsums_b200/bin/gps_tune
echo "== sweep"; timeout 900 $T --log2n 30 --sweep --pdl > gpurun_out/r02_sweep_pdl.jsonl 2> gpurun_out/r02_sweep_pdl.err; echo "rc=$?"
echo "== sweep nopdl"; timeout 900 $T --log2n 30 --sweep > gpurun_out/r02_sweep_nopdl.jsonl 2> gpurun_out/r02_sweep_nopdl.err; echo "rc=$?"
echo "== sweep cold"; timeout 900 $T --log2n 26 --sweep --pdl --flush > gpurun_out/r02_sweep_flush.jsonl 2> gpurun_out/r02_sweep_flush.err; echo "rc=$?"
wc -l gpurun_out/r02_sweep_*.jsonl
echo "== library dispatch through the C ABI, 2^20..2^30"; python - <<'P'
import json, sys, torch
sys.path.insert(0, '.')
import gpuprefixsums_b200 as gps
ps = gps.PrefixSum(1 << 30)
rows = []
for lg in range(20, 31):
    n = 1 << lg
    x = torch.empty(n, dtype=torch.int32, device='cuda'); gps.init_random(x, seed=1); out = torch.empty_like(x)
    iters = 200 if lg <= 24 else 50 if lg <= 28 else 20
    for _ in range(5): ps.scan(x, out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): ps.scan(x, out)
    e1.record(); torch.cuda.synchronize(); ps.check()
    us = e0.elapsed_time(e1) * 1e3 / iters
    rows.append({"log2n": lg, "us": round(us, 2), "gkeys": round(n / us / 1e3, 1), "frac_of_6456": round(8 * n / us / 1e3 / 6456.2, 3)})
    print(rows[-1], flush=True)
    del x, out
json.dump(rows, open('gpurun_out/r02_library_sweep.json', 'w'))
P
