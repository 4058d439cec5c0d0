#!/usr/bin/env python
"""Developer probe: host-buffer (end-to-end) scan strategies at 2^28 keys.
(a) gps_scan_host with several pipeline chunk sizes, (b) zero-copy: the scan kernel reads
and writes pinned host memory directly over PCIe, (c) raw H2D / D2H copy bandwidth."""
import ctypes as C
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gpuprefixsums_b200 as gps  # noqa: E402
from gpuprefixsums_b200 import _capi as capi  # noqa: E402

n = 1 << 28
res = {}
h_in = torch.randint(0, 1 << 30, (n,), dtype=torch.int32).pin_memory()
h_out = torch.empty(n, dtype=torch.int32).pin_memory()
d = torch.empty(n, dtype=torch.int32, device="cuda")


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps


t = timed(lambda: d.copy_(h_in, non_blocking=True))
res["h2d_GBs"] = 4 * n / t / 1e9
t = timed(lambda: h_out.copy_(d, non_blocking=True))
res["d2h_GBs"] = 4 * n / t / 1e9
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
d2 = torch.empty_like(d)


def both():
    with torch.cuda.stream(s1):
        d.copy_(h_in, non_blocking=True)
    with torch.cuda.stream(s2):
        h_out.copy_(d2, non_blocking=True)


t = timed(both)
res["h2d_and_d2h_concurrent_GBs_each"] = 4 * n / t / 1e9

for lg in (21, 22, 23, 24, 25, 26):
    ps = gps.PrefixSum(n)
    ps.set_option("host_chunk_keys", 1 << lg)
    t = timed(lambda: ps.scan_host(h_in, h_out))
    res[f"scan_host_chunk_2^{lg}_gkeys"] = n / t / 1e9
    ps.close()

ref = gps.PrefixSum(n)
want = ref.scan(h_in.cuda()).cpu()
assert torch.equal(want, h_out), "scan_host mismatch"

# zero-copy: pinned host pointers are device-accessible under UVA
zc = gps.PrefixSum(n)


def zero_copy(algo):
    capi.check(capi.lib.gps_scan(zc._ctx, algo, capi.INCLUSIVE, capi.U32, h_in.data_ptr(), h_out.data_ptr(), n, 0,
                                 None, torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()


for name, algo in (("csdl_reg", capi.ALGO_CSDL_REG), ("csdl_tma", capi.ALGO_CSDL_TMA)):
    try:
        h_out.zero_()
        t = timed(lambda: zero_copy(algo), reps=3)
        zc.check()
        res[f"zero_copy_{name}_gkeys"] = n / t / 1e9
        res[f"zero_copy_{name}_ok"] = bool(torch.equal(want, h_out))
    except Exception as e:  # noqa: BLE001
        res[f"zero_copy_{name}_error"] = str(e)
# hybrid: zero-copy read, device output + D2H? and H2D + zero-copy write
print(json.dumps(res, indent=1))
