"""Burst vs sustained throughput at 2^28 uint32 keys: the lane-strided smem kernel (default
GPS_ALGO_CSDL) against the blocked-tile kernel (gps_scan_wide at 32 bits), 30 / 500 / 4000
launches back to back; clocks and power sampled by nvidia-smi while the long run is in flight."""
import json
import os
import subprocess
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gpuprefixsums_b200 as gps  # noqa: E402

n = 1 << 28
ps = gps.PrefixSum(n)
x = torch.randint(0, 1 << 30, (n,), dtype=torch.int32, device="cuda")
out = torch.empty_like(x)
res = {}


def timed(fn, iters):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return round(n / (e0.elapsed_time(e1) / iters) / 1e6, 1)


def smi():
    o = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,power.draw,clocks_throttle_reasons.sw_power_cap",
                        "--format=csv,noheader"], capture_output=True, text=True).stdout.strip().split("\n")[0]
    return o


kernels = {"csdl_smem": lambda: ps.scan(x, out, algo="csdl_smem"),
           "blocked": lambda: ps.scan_wide(x, out, algo="csdl"),
           "csdl_reg": lambda: ps.scan(x, out, algo="csdl_reg")}
for rep in range(2):
    for name, fn in kernels.items():
        for iters in (30, 500, 4000):
            res[f"{name}_{iters}_r{rep}"] = timed(fn, iters)
            time.sleep(0.5)
        # sample clocks / power in the middle of a long run
        for _ in range(3000):
            fn()
        res[f"{name}_smi_r{rep}"] = smi()
        torch.cuda.synchronize()
        time.sleep(2.0)
ps.check()
print(json.dumps(res, indent=0))
