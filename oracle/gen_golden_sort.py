#!/usr/bin/env python
"""oracle/gen_golden_sort.py -- TEST INFRASTRUCTURE.  Generates tests/golden/cub_sort_*.npz by running
cub::DeviceRadixSort (CCCL 2.8.2 of the CUDA 12.9 toolkit, through gpuprefixsums_b200/lib/libgps_cub.so --
library code, not ours) on a B200:

    gpurun -- python oracle/gen_golden_sort.py gpurun_out/golden_sort      # then copy into tests/golden/

The radix sort the reference's README names as the consumer of its scans lives in a sibling repository
that is not under /root/reference, so there are no reference vectors for it; these fixtures pin the
sort oracle (oracle_sort_u32 / oracle_sort_keys32) on the toolkit's own radix sort instead of on numpy
alone.  Inputs are the seeded arrays of oracle_init_random_u32 (seed and mask recorded); small cases
store the full sorted keys and payload, large ones a checksum of checksums:
[sum of keys[i] * (i + 1) mod 2^64, xor of keys[i] rotated by (i mod 32), sum of vals[i] * (i + 1) mod 2^64].
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gpuprefixsums_b200 as gps  # noqa: E402  (only for the CUB wrapper library)
from tests import oracle_lib  # noqa: E402


def checksums(keys: np.ndarray, vals: np.ndarray) -> np.ndarray:
    idx = np.arange(1, keys.size + 1, dtype=np.uint64)
    k = keys.astype(np.uint64)
    rot = (np.arange(keys.size, dtype=np.uint64) % np.uint64(32))
    xr = np.bitwise_xor.reduce(((k << rot) | (k >> (np.uint64(32) - rot))) & np.uint64(0xFFFFFFFF)) if keys.size else np.uint64(0)
    return np.array([(k * idx).sum(dtype=np.uint64), xr, (vals.astype(np.uint64) * idx).sum(dtype=np.uint64)], dtype=np.uint64)


def main(outdir: str) -> None:
    os.makedirs(outdir, exist_ok=True)
    oracle = oracle_lib.load()
    assert torch.cuda.is_available()
    cases = [  # (name, n, seed, mask, end_bit, store full output?)
        ("full_4099", 4099, 11, 0xFFFFFFFF, 32, True),
        ("full_65537", 65537, 12, 0xFFFFFFFF, 32, True),
        ("few_20000", 20000, 13, 0x3, 32, True),              # four distinct keys: stability decides the payload
        ("bits13_50001", 50001, 14, 0xFFFFFFFF, 13, True),    # end_bit not a multiple of 8, full-range keys
        ("bits24_30011", 30011, 15, 0xFFFFFFFF, 24, True),
        ("full_2p22", 1 << 22, 16, 0xFFFFFFFF, 32, False),
        ("bits17_2p22", 1 << 22, 17, 0xFFFFFFFF, 17, False),
        ("full_2p24", 1 << 24, 18, 0xFFFFFFFF, 32, False),
    ]
    for name, n, seed, mask, end_bit, full in cases:
        x = oracle.random_u32(n, seed, mask)
        v = np.arange(n, dtype=np.uint32)
        dk = torch.from_numpy(x.view(np.int32)).cuda()
        dv = torch.from_numpy(v.view(np.int32)).cuda()
        ko, vo = torch.empty_like(dk), torch.empty_like(dv)
        gps.cub_sort(dk, ko, dv, vo, end_bit=end_bit)
        torch.cuda.synchronize()
        keys = ko.cpu().numpy().view(np.uint32)
        vals = vo.cpu().numpy().view(np.uint32)
        rec = {"n": n, "seed": seed, "mask": mask, "end_bit": end_bit, "sums": checksums(keys, vals)}
        if full:
            rec["keys"], rec["vals"] = keys, vals
        np.savez_compressed(os.path.join(outdir, f"cub_sort_{name}.npz"), **rec)
        print(name, n, end_bit, rec["sums"])


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "golden_sort"))
