"""Throughput of gps_sort_keys64 (uint64 keys, keys / pairs with a 32-bit payload) next to
cub::DeviceRadixSort on the same keys.  Prints one JSON line; GPS_B200_LIB selects an alternative build."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gpuprefixsums_b200 as gps  # noqa: E402

log2n = int(sys.argv[1]) if len(sys.argv) > 1 else 27
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 10
n = 1 << log2n
ps = gps.PrefixSum(1 << 20)
g = torch.Generator(device="cuda").manual_seed(1)
keys = torch.randint(-(1 << 62), 1 << 62, (n,), dtype=torch.int64, device="cuda", generator=g) * 3 + 1  # wraps: all 64 bits vary
vals = torch.arange(n, dtype=torch.int32, device="cuda")
ko, vo = torch.empty_like(keys), torch.empty_like(vals)
res = {"lib": os.path.basename(os.environ.get("GPS_B200_LIB", "default")), "log2n": log2n, "key_bits": 64}


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


res["gps_keys"] = timed(lambda: ps.sort(keys, keys_out=ko, key_type="u64"))
ps.check()
ref = torch.empty_like(keys)
gps.cub_sort(keys, ref)
res["keys_equal_cub"] = bool(torch.equal(ko, ref))
res["gps_pairs"] = timed(lambda: ps.sort(keys, vals, keys_out=ko, values_out=vo, key_type="u64"))
ps.check()
refv = torch.empty_like(vals)
gps.cub_sort(keys, ref, vals, refv)
res["pairs_equal_cub"] = bool(torch.equal(ko, ref) and torch.equal(vo, refv))
res["cub_keys"] = timed(lambda: gps.cub_sort(keys, ref))
res["cub_pairs"] = timed(lambda: gps.cub_sort(keys, ref, vals, refv))
res["gps_keys_f64"] = timed(lambda: ps.sort(keys.view(torch.float64), key_type="f64"))
res["gps_keys_48bit"] = timed(lambda: ps.sort(keys, keys_out=ko, key_type="u64", end_bit=48))
res["gps_keys_fallback"] = timed(lambda: ps.sort(keys, keys_out=ko, key_type="u64", fallback=True, max_spin=64))
ps.check()
gps.cub_sort(keys, ref)
res["keys_fallback_equal_cub"] = bool(torch.equal(ko, ref))
# floor of the 8 digit passes + histogram sweeps: keys 8 x 16 B + 2 x 8 B = 144 B/key, pairs 8 x 24 + 16 = 208 B/key
res["floor_bytes_per_key"] = {"keys": 144, "pairs": 208}
ps.close()
print(json.dumps(res))
