"""Reduce-then-Scan at 2^28 uint32 keys: propagate kernel on the register tile vs the blocked
shared-memory tile, with and without programmatic dependent launch (two rounds)."""
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
for rnd in range(2):
    for variant, name in ((1, "reg"), (2, "blocked")):
        for pdl in (0, 1):
            ps.set_option("rts_variant", variant)
            ps.set_option("pdl", pdl)
            for _ in range(5):
                ps.scan(x, out, algo="rts")
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(200):
                ps.scan(x, out, algo="rts")
            e1.record()
            torch.cuda.synchronize()
            res[f"{name}_pdl{pdl}_r{rnd}"] = round(n / (e0.elapsed_time(e1) / 200) / 1e6, 1)
ps.check()
print(json.dumps(res))
