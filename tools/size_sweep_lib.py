#!/usr/bin/env python
"""BASELINE configs[3] through the LIBRARY (gps_scan, GPS_ALGO_CSDL, default options) next to cub::DeviceScan: uint32 inclusive,
2^10..2^30 keys, out of place.  Back to back = 50 scans captured in one CUDA graph and replayed (so that the Python loop does not
bound the small sizes; CUB is captured the same way); cold = 512 MB written before every launch, the launch timed alone with its
own pair of events (adds ~5 us).  Every result is compared with CUB's.  One JSON line per size."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gpuprefixsums_b200 as gps  # noqa: E402

max_lg = int(sys.argv[1]) if len(sys.argv) > 1 else 30
ps = gps.PrefixSum(1 << max_lg)
flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
side = torch.cuda.Stream()
PEAK = 6456.2e9


def graph_us(fn, n):
    per_graph = 50 if n <= (1 << 26) else 10
    g = torch.cuda.CUDAGraph()
    with torch.cuda.stream(side):
        fn(side)
        side.synchronize()
        with torch.cuda.graph(g, stream=side):
            for _ in range(per_graph):
                fn(side)
    for _ in range(2):
        g.replay()
    torch.cuda.synchronize()
    reps = 10 if n <= (1 << 24) else 4
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (per_graph * reps)


def cold_us(fn):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    tot = 0.0
    for i in range(10):
        flush.fill_(i)
        e0.record()
        fn(None)
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1) * 1e3
    return tot / 10


for lg in range(10, max_lg + 1):
    n = 1 << lg
    x = torch.empty(n, dtype=torch.int32, device="cuda")
    gps.init_random(x, seed=3)
    out, ref = torch.empty_like(x), torch.empty_like(x)
    gps.cub_scan(x, ref)
    torch.cuda.synchronize()
    ours = lambda st: ps.scan(x, out, stream=st)

    def cub(st):
        if st is None:
            gps.cub_scan(x, ref)
        else:
            with torch.cuda.stream(st):
                gps.cub_scan(x, ref)

    us = graph_us(ours, n)
    ps.check()
    equal = gps.compare_u32(out, ref)[0] == 0
    try:
        us_cub = graph_us(cub, n)
        cub_how = "graph"
    except Exception:  # not capturable: plain loop
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(50):
            gps.cub_scan(x, ref)
        e1.record()
        torch.cuda.synchronize()
        us_cub, cub_how = e0.elapsed_time(e1) * 1e3 / 50, "loop"
    row = {"log2n": lg, "us": round(us, 2), "Gkeys_s": round(n / us / 1e3, 1), "frac_of_measured_hbm": round(8 * n / (us * 1e-6) / PEAK, 3),
           "us_cold_l2": round(cold_us(ours), 2) if lg <= 27 else None, "cub_us": round(us_cub, 2), "cub_timing": cub_how,
           "speedup_vs_cub": round(us_cub / us, 2), "equal_cub": equal, "l2_resident": 8 * n <= 126e6}
    print(json.dumps(row), flush=True)
    del x, out, ref
ps.close()
