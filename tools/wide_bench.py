"""Throughput of the wide-element scans (uint64 / float64 / vec4<u32>) at 1 GiB in + 1 GiB out,
both algorithms.  Prints one JSON line; GPS_B200_LIB selects an alternative build."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gpuprefixsums_b200 as gps  # noqa: E402

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 30
ps = gps.PrefixSum(1 << 28)
res = {"lib": os.environ.get("GPS_B200_LIB", "default")}
cases = [("u32", torch.randint(0, 1 << 30, (1 << 28,), dtype=torch.int32, device="cuda"), {}),
         ("u64", torch.ones(1 << 27, dtype=torch.int64, device="cuda"), {}),
         ("f64", torch.rand(1 << 27, dtype=torch.float64, device="cuda"), {}),
         ("u32x4", torch.ones((1 << 26, 4), dtype=torch.int32, device="cuda"), {"struct4": True})]
for name, t, kw in cases:
    out = torch.empty_like(t)
    for algo in ("csdl", "csdl_fallback", "rts"):
        for _ in range(3):
            ps.scan_wide(t, out, algo=algo, **kw)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            ps.scan_wide(t, out, algo=algo, **kw)
        e1.record()
        torch.cuda.synchronize()
        ps.check()
        ms = e0.elapsed_time(e1) / iters
        nbytes = t.numel() * t.element_size()
        res[f"{name}_{algo}"] = {"ms": round(ms, 4), "GB_s_useful": round(2 * nbytes / ms / 1e6, 1),
                                 "Gelem_s": round(t.shape[0] / ms / 1e6, 2)}
    # Decoupled Fallback under emulated stalls (mask 7, spin 4: the reference's harness settings)
    if name != "u32":
        stats = torch.zeros(4, dtype=torch.int32, device="cuda")
        for _ in range(2):
            ps.scan_wide_emulated_deadlock(t, out, stall_mask=7, max_spin=4, **kw)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            ps.scan_wide_emulated_deadlock(t, out, stall_mask=7, max_spin=4, stats=stats, **kw)
        e1.record()
        torch.cuda.synchronize()
        ps.check()
        ms = e0.elapsed_time(e1) / iters
        nbytes = t.numel() * t.element_size()
        res[f"{name}_emulated_deadlock_mask7_spin4"] = {
            "ms": round(ms, 4), "GB_s_useful": round(2 * nbytes / ms / 1e6, 1),
            "fallbacks_per_launch": int(stats[1].item()) // iters, "insertions_per_launch": int(stats[2].item()) // iters}
# the 32-bit production kernels in the same process, for comparison
t = cases[0][1]
out = torch.empty_like(t)
for algo in ("csdl_smem", "csdl_reg"):
    for _ in range(3):
        ps.scan(t, out, algo=algo)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        ps.scan(t, out, algo=algo)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    res[f"u32_{algo}"] = {"ms": round(ms, 4), "GB_s_useful": round(2 * t.numel() * 4 / ms / 1e6, 1)}
ps.close()
print(json.dumps(res))
