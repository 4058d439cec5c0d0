"""Back-to-back small scans: plain launches vs one CUDA graph of 100 captured scans (replayed),
chained scan (shadow scratch cleared inside the graph) vs Reduce-then-Scan (no launch state)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gpuprefixsums_b200 as gps  # noqa: E402

res = {}
ps = gps.PrefixSum(1 << 22)
for log2n in (12, 16, 20):
    n = 1 << log2n
    x = torch.randint(0, 1 << 20, (n,), dtype=torch.int32, device="cuda")
    out = torch.empty_like(x)
    for algo in ("csdl", "rts"):
        for _ in range(3):
            ps.scan(x, out, algo=algo)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(1000):
            ps.scan(x, out, algo=algo)
        e1.record()
        torch.cuda.synchronize()
        plain = e0.elapsed_time(e1)  # us per scan (ms per 1000)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(100):
                ps.scan(x, out, algo=algo)
        g.replay()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(10):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        res[f"2^{log2n}_{algo}"] = {"plain_us": round(plain, 2), "graph_us": round(e0.elapsed_time(e1), 2)}
ps.check()
print(json.dumps(res))
