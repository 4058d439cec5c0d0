"""Decoupled Fallback kernels at 2^28 uint32 keys: register tile vs shared-memory-resident tile,
without stalls and with the reference's emulation parameters (mask 7, spin 4)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gpuprefixsums_b200 as gps  # noqa: E402

n = 1 << 28
ps = gps.PrefixSum(n)
x = torch.randint(0, 1 << 30, (n,), dtype=torch.int32, device="cuda")
out = torch.empty_like(x)
res = {}


def timed(fn, iters=30):
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


for variant, name in [(1, "reg"), (2, "smem"), (3, "blocked")]:
    ps.set_option("fallback_variant", variant)
    res[f"{name}_no_stall"] = timed(lambda: ps.scan(x, out, algo="csdl_fallback"))
    for mask, spin in [(7, 4), (63, 4), (7, 1)]:
        stats = torch.zeros(4, dtype=torch.int32, device="cuda")
        res[f"{name}_mask{mask}_spin{spin}"] = timed(lambda: ps.scan_emulated_deadlock(x, out, stall_mask=mask, max_spin=spin, stats=stats), 10)
        res[f"{name}_mask{mask}_spin{spin}_stats_per_launch"] = [int(v) // 13 for v in stats.cpu()]
ps.check()
res["csdl_default"] = timed(lambda: ps.scan(x, out, algo="csdl"))
for variant, name in [(1, "reg"), (2, "blocked")]:
    ps.set_option("rts_variant", variant)
    res[f"rts_{name}"] = timed(lambda: ps.scan(x, out, algo="rts"))
    res[f"rts_{name}_500"] = timed(lambda: ps.scan(x, out, algo="rts"), 500)
ps.set_option("fallback_variant", 0)
res["fallback_auto_500"] = timed(lambda: ps.scan(x, out, algo="csdl_fallback"), 500)
print(json.dumps(res))
