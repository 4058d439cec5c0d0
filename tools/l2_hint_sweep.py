#!/usr/bin/env python
"""Mid-size regime: L2 cache policies on the tile loads / stores (option "l2_hints") and the
dispatch by size, 2^20..2^27 keys, back-to-back launches (as the reference's BatchTiming does) and
cold-L2 launches (512 MB written between launches, each launch timed alone)."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gpuprefixsums_b200 as gps  # noqa: E402

out_path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "l2_hint_sweep.jsonl")
ps = gps.PrefixSum(1 << 27)
flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
rows = []
for lg in range(20, 28):
    n = 1 << lg
    x = torch.empty(n, dtype=torch.int32, device="cuda")
    gps.init_random(x, seed=3)
    out = torch.empty_like(x)
    ref = torch.empty_like(x)
    gps.cub_scan(x, ref)
    for algo in ("csdl", "csdl_smem", "csdl_blocked", "csdl_reg"):
        for hints in (0, 1, 2, 3):
            ps.set_option("l2_hints", hints)
            ps.scan(x, out, algo=algo)
            ps.check()
            ok = gps.compare_u32(out, ref)[0] == 0
            iters = 300 if lg <= 24 else 100
            for _ in range(5):
                ps.scan(x, out, algo=algo)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(iters):
                ps.scan(x, out, algo=algo)
            e1.record()
            torch.cuda.synchronize()
            us = e0.elapsed_time(e1) * 1e3 / iters
            cold = 0.0
            for i in range(20):
                flush.fill_(i)
                e0.record()
                ps.scan(x, out, algo=algo)
                e1.record()
                torch.cuda.synchronize()
                cold += e0.elapsed_time(e1) * 1e3
            cold /= 20
            # scan followed by a consumer that reads the output once (sum): the pair, back to back
            for _ in range(3):
                ps.scan(x, out, algo=algo)
                out.sum()
            torch.cuda.synchronize()
            e0.record()
            for _ in range(50):
                ps.scan(x, out, algo=algo)
                out.sum()
            e1.record()
            torch.cuda.synchronize()
            pair = e0.elapsed_time(e1) * 1e3 / 50
            row = {"log2n": lg, "algo": algo, "l2_hints": hints, "ok": ok, "us_back_to_back": round(us, 2),
                   "gkeys": round(n / us / 1e3, 1), "frac_of_measured_6456": round(8 * n / us / 1e3 / 6456.2, 3),
                   "us_cold_l2": round(cold, 2), "us_scan_plus_consumer": round(pair, 2)}
            rows.append(row)
            print(json.dumps(row), flush=True)
with open(out_path, "w") as f:
    for r in rows:
        f.write(json.dumps(r) + "\n")
