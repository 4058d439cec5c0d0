#!/usr/bin/env python
"""Developer tool: run / time the reference's unmodified CUDA kernels (oracle/_ref shim)
on seeded random input, without torch (so ncu stays light).
  python tools/run_ref_kernel.py [log2n] [iters] [mask]
Prints Gkeys/s per reference algorithm with the reference's own timing protocol."""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gpuprefixsums_b200 import _capi as capi  # noqa: E402
from tests import oracle_lib  # noqa: E402

log2n = int(sys.argv[1]) if len(sys.argv) > 1 else 28
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 20
mask = int(sys.argv[3], 0) if len(sys.argv) > 3 else 3
n = 1 << log2n
shim = oracle_lib.load_ref_shim()
src, buf = C.c_void_p(), C.c_void_p()
capi.check(capi.lib.gps_malloc(C.byref(src), n * 4))
capi.check(capi.lib.gps_malloc(C.byref(buf), n * 4))
capi.check(capi.lib.gps_init_random(src, n, capi.U32, 1234, mask, 0, None))
capi.check(capi.lib.gps_device_sync())
res = {}
for name, algo in [("ref_csdl", 0), ("ref_rts", 1), ("ref_emulated_deadlock", 2), ("ref_cub", 3)]:
    ms = shim.time_u32(algo, False, buf.value, src.value, n, iters)
    res[name] = {"ms": ms, "gkeys_s": n / (ms * 1e-3) / 1e9 if ms > 0 else None}
    ms1 = shim.time_u32(algo, False, buf.value, None, n, iters)  # all-ones (reference protocol)
    res[name + "_ones"] = {"ms": ms1, "gkeys_s": n / (ms1 * 1e-3) / 1e9 if ms1 > 0 else None}
print(json.dumps({"n": n, "mask": mask, "results": res}))
