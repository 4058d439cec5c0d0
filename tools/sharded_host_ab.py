#!/usr/bin/env python
"""Host-buffer scans (gps_scan_host on one GPU, gps_sharded_scan_host on all visible GPUs): option "ticket_free" 1 vs 0.
The chunks of the host pipeline are 2^24 keys, i.e. exactly the size class whose default kernel changed; their scans chain
a running total through carry slots.  Prints one JSON line per (path, setting, repetition)."""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gpuprefixsums_b200 as gps  # noqa: E402

ngpu = torch.cuda.device_count()
steps = 10
n1 = 1 << 28
h_in = torch.empty(n1 * ngpu, dtype=torch.int32).pin_memory()
h_out = torch.empty_like(h_in).pin_memory()
tmp = torch.empty(n1, dtype=torch.int32, device="cuda:0")
for r in range(ngpu):
    gps.init_random(tmp, seed=1, first_index=r * n1)
    h_in[r * n1:(r + 1) * n1].copy_(tmp)
del tmp
want_last = int(h_in.to(torch.int64).sum().item() & 0xFFFFFFFF)

ps = gps.PrefixSum(1 << 28, 0)
for tf in (1, 0, 1, 0):
    ps.set_option("ticket_free", tf)
    for _ in range(2):
        ps.scan_host(h_in[:n1], h_out[:n1], wait=False)
    ps.host_sync()
    t0 = time.perf_counter()
    for _ in range(steps):
        ps.scan_host(h_in[:n1], h_out[:n1], wait=False)
    ps.host_sync()
    s = (time.perf_counter() - t0) / steps
    print(json.dumps({"path": "gps_scan_host, 1 GPU, 2^28 keys", "ticket_free": tf, "Gkeys_s": round(n1 / s / 1e9, 3)}), flush=True)
ps.close()

if ngpu > 1:
    for tf in (1, 0, 1, 0):
        hs = gps.ShardedHostScan(list(range(ngpu)))
        hs.set_option("ticket_free", tf)
        io, rate = hs.calibrate(h_in, h_out)
        for _ in range(2):
            hs.scan_host(h_in, h_out, wait=False)
        hs.sync()
        t0 = time.perf_counter()
        for _ in range(steps):
            hs.scan_host(h_in, h_out, wait=False)
        hs.sync()
        s = (time.perf_counter() - t0) / steps
        ok = (int(h_out[-1].item()) & 0xFFFFFFFF) == want_last
        print(json.dumps({"path": f"gps_sharded_scan_host, {ngpu} GPUs, {ngpu} x 2^28 keys", "ticket_free": tf,
                          "Gkeys_s": round(n1 * ngpu / s / 1e9, 3), "io_gpus": io, "calibrated_gbs_each_way": round(rate, 1),
                          "frac_of_calibrated": round(n1 * ngpu * 4 / s / 1e9 / rate, 3), "last_key_ok": ok}), flush=True)
        hs.close()
