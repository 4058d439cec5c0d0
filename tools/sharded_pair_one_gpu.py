#!/usr/bin/env python
"""The kernel sequence of ONE rank's sharded step (reduce + fold/publish + scan with carry slots) on a
single GPU, for ncu: the slot array has one rank, everything else is what bench.py --gpus N runs."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gpuprefixsums_b200 as gps  # noqa: E402

log2n = int(sys.argv[1]) if len(sys.argv) > 1 else 28
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 8
n = 1 << log2n
x = torch.empty(n, dtype=torch.int32, device="cuda")
gps.init_random(x, seed=1)
out = torch.empty_like(x)
ps = gps.PrefixSum(n)
slots = torch.zeros(32, dtype=torch.int64, device="cuda")
for i in range(steps):
    tag, par = i + 1, (i + 1) & 1
    base = slots.data_ptr() + par * 16 * 8
    ps.reduce_publish(x, [base], tag)
    ps.scan_carry_slots(x, out, base, 1, 0, tag)
ps.check()
ref = torch.empty_like(x)
gps.cub_scan(x, ref)
assert gps.compare_u32(out, ref)[0] == 0
print("sharded pair ok")
