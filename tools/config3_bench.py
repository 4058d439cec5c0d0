#!/usr/bin/env python
"""BASELINE configs[2] as a measurement: float32 inclusive scan of 2^28 values in [0, 1) -- throughput of every algorithm
next to cub::DeviceScan, and the largest error against the float64 serial oracle (bound: rtol * sum |x_j|, rtol = 2e-5);
plus the emulated-deadlock Decoupled Fallback path (stall mask 7, spin budget 4) on the same input.  One JSON object."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gpuprefixsums_b200 as gps  # noqa: E402
from tests import oracle_lib  # noqa: E402

oracle = oracle_lib.load()
n = 1 << 28
ps = gps.PrefixSum(n)
x = torch.empty(n, dtype=torch.float32, device="cuda")
gps.init_random(x, seed=11)
hx = x.cpu().numpy()
out = torch.empty_like(x)
res = {"workload": "float32 inclusive scan, 2^28 values uniform in [0,1), one B200 (BASELINE configs[2])", "rtol": 2e-5,
       "hbm_peak_measured_gbs": 6456.2, "rows": {}}


def timed(fn, iters=50):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def row(name, fn, iters=50):
    ms = timed(fn, iters)
    ps.check()
    errs, max_rel = oracle.check_f32(hx, out.cpu().numpy(), rtol=2e-5, atol=1e-6)
    res["rows"][name] = {"ms": round(ms, 4), "Gvalues_s": round(n / ms / 1e6, 1), "frac_of_measured_hbm": round(8 * n / ms / 1e6 / 6456.2, 3),
                         "outside_tolerance": int(errs), "max_rel_error": float(f"{max_rel:.3e}")}


for algo in ("csdl", "csdl_smem", "csdl_reg", "rts", "csdl_fallback"):
    row(algo, lambda a=algo: ps.scan(x, out, algo=a))
row("emulated_deadlock_mask7_spin4", lambda: ps.scan_emulated_deadlock(x, out, stall_mask=7, max_spin=4), iters=20)
row("cub_device_scan", lambda: gps.cub_scan(x, out))
row("csdl_exclusive", lambda: ps.scan(x, out, algo="csdl", exclusive=True))
# the exclusive row is checked as exclusive
errs, max_rel = oracle.check_f32(hx, out.cpu().numpy(), exclusive=True, rtol=2e-5, atol=1e-6)
res["rows"]["csdl_exclusive"]["outside_tolerance"] = int(errs)
res["rows"]["csdl_exclusive"]["max_rel_error"] = float(f"{max_rel:.3e}")
ps.close()
print(json.dumps(res))
