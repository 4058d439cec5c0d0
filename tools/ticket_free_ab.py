#!/usr/bin/env python
"""GPS_ALGO_CSDL through the library, option "ticket_free" 1 (default) vs 0, sizes 2^19..2^26 (+ ragged ones):
back-to-back launches inside a CUDA graph of 50 scans (so the Python loop does not bound the small sizes) and cold-L2
launches; every result compared with cub::DeviceScan.  One JSON line per (size, setting)."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gpuprefixsums_b200 as gps  # noqa: E402

IN_PLACE = "--in-place" in sys.argv  # time gps_scan(x, x): results are then not comparable (parity in place: tests/)
ps = gps.PrefixSum(1 << 26)
flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
sizes = [1 << lg for lg in range(10, 27)] + [(1 << 20) + (1 << 19), 3 << 20, 3_800_000, 4_500_000, 5_000_000, 6_000_000, 3 << 22, 3 << 23]
if "big" in sys.argv:  # only the upper mid sizes (a build with a wider ticket-free range)
    sizes = [1 << 24, 3 << 23, 1 << 25, 3 << 24, 1 << 26, 100_000_000, (1 << 27) - 4096]
    ps.close()
    ps = gps.PrefixSum(1 << 27)
side = torch.cuda.Stream()
for n in sorted(sizes):
    x = torch.empty(n, dtype=torch.int32, device="cuda")
    gps.init_random(x, seed=3)
    out, ref = torch.empty_like(x), torch.empty_like(x)
    gps.cub_scan(x, ref)
    if IN_PLACE:
        out = x
    for tf in ((0, 1, 2, 0, 1, 2) if "three" in sys.argv else (0, 1, 0, 1)):
        ps.set_option("ticket_free", tf)
        ps.scan(x, out)
        ps.check()
        ok = IN_PLACE or gps.compare_u32(out, ref)[0] == 0
        g = torch.cuda.CUDAGraph()
        with torch.cuda.stream(side):
            ps.scan(x, out, stream=side)
            side.synchronize()
            with torch.cuda.graph(g, stream=side):
                for _ in range(50):
                    ps.scan(x, out, stream=side)
        for _ in range(2):
            g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 10 if n <= (1 << 24) else 4
        e0.record()
        for _ in range(reps):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / (50 * reps)
        ps.check()
        ok = ok and (IN_PLACE or gps.compare_u32(out, ref)[0] == 0)
        cold = 0.0
        for i in range(20):
            flush.fill_(i)
            e0.record()
            ps.scan(x, out)
            e1.record()
            torch.cuda.synchronize()
            cold += e0.elapsed_time(e1) * 1e3
        cold /= 20
        print(json.dumps({"in_place": IN_PLACE, "n": n, "log2n": round(torch.log2(torch.tensor(float(n))).item(), 3), "ticket_free": tf,
                          "us_graph_back_to_back": round(us, 2), "us_cold_l2": round(cold, 2), "equal_cub": ok,
                          "frac_of_measured_hbm": round(8 * n / (us * 1e-6) / 6456.2e9, 3)}), flush=True)
ps.close()
