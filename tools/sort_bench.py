"""Throughput of gps_sort_u32 (keys / pairs) next to cub::DeviceRadixSort, random uint32 keys.
Prints one JSON line; GPS_B200_LIB selects an alternative build."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gpuprefixsums_b200 as gps  # noqa: E402

log2n = int(sys.argv[1]) if len(sys.argv) > 1 else 28
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 10
n = 1 << log2n
ps = gps.PrefixSum(1 << 20)
g = torch.Generator(device="cuda").manual_seed(1)
keys = torch.randint(-(1 << 31), (1 << 31) - 1, (n,), dtype=torch.int32, device="cuda", generator=g)
vals = torch.arange(n, dtype=torch.int32, device="cuda")
ko, vo = torch.empty_like(keys), torch.empty_like(vals)
res = {"lib": os.environ.get("GPS_B200_LIB", "default"), "log2n": log2n}


def timed(fn):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    return {"ms": round(ms, 4), "Gkeys_s": round(n / ms / 1e6, 2)}


res["gps_keys"] = timed(lambda: ps.sort(keys, keys_out=ko))
ps.check()
ref = torch.empty_like(keys)
gps.cub_sort(keys, ref)
res["keys_equal_cub"] = bool(torch.equal(ko, ref))
res["gps_pairs"] = timed(lambda: ps.sort(keys, vals, keys_out=ko, values_out=vo))
ps.check()
refv = torch.empty_like(vals)
gps.cub_sort(keys, ref, vals, refv)
res["pairs_equal_cub"] = bool(torch.equal(ko, ref) and torch.equal(vo, refv))
res["cub_keys"] = timed(lambda: gps.cub_sort(keys, ref))
res["cub_pairs"] = timed(lambda: gps.cub_sort(keys, ref, vals, refv))
kf = keys.view(torch.float32)
res["gps_keys_f32"] = timed(lambda: ps.sort(kf, key_type="f32"))
res["gps_keys_16bit"] = timed(lambda: ps.sort(keys, keys_out=ko, end_bit=16))
# Decoupled Fallback in the digit look-back: production variant (no stalls) and emulated stalls
res["gps_keys_fallback"] = timed(lambda: ps.sort(keys, keys_out=ko, fallback=True, max_spin=64))
ps.check()
gps.cub_sort(keys, ref)
res["keys_fallback_equal_cub"] = bool(torch.equal(ko, ref))
res["gps_pairs_fallback"] = timed(lambda: ps.sort(keys, vals, keys_out=ko, values_out=vo, fallback=True, max_spin=64))
stats = torch.zeros(4, dtype=torch.int32, device="cuda")
res["gps_keys_emulated_deadlock_mask7_spin4"] = timed(
    lambda: ps.sort(keys, keys_out=ko, fallback=True, stall_mask=7, max_spin=4, stats=stats))
ps.check()
res["keys_emulated_equal_cub"] = bool(torch.equal(ko, ref))
res["emulated_stats_per_sort"] = [int(v) // (iters + 2) for v in stats.tolist()]
ps.close()
print(json.dumps(res))
