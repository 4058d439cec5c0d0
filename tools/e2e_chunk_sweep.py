"""gps_scan_host (pinned host in -> pinned host out) at 2^28 uint32 keys for several pipeline
chunk sizes; the PCIe floor is 2 x 1 GiB at ~50 GB/s per direction (tools/e2e_probe.py)."""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gpuprefixsums_b200 as gps  # noqa: E402

n = 1 << 28
h_in = torch.randint(0, 1 << 30, (n,), dtype=torch.int32).pin_memory()
h_out = torch.empty(n, dtype=torch.int32).pin_memory()
res = {}
for lg in (20, 21, 22, 23, 24, 25):
    ps = gps.PrefixSum(n)
    ps.set_option("host_chunk_keys", 1 << lg)
    for _ in range(2):
        ps.scan_host(h_in, h_out)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(10):
        ps.scan_host(h_in, h_out)
    dt = (time.perf_counter() - t0) / 10
    res[f"chunk_2^{lg}"] = {"ms": round(dt * 1e3, 2), "Gkeys_s": round(n / dt / 1e9, 2)}
    ps.close()
print(json.dumps(res))
