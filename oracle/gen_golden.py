#!/usr/bin/env python
"""oracle/gen_golden.py -- TEST INFRASTRUCTURE.  Generates tests/golden/ref_cuda_*.npz by
running the reference's own, unmodified CUDA kernels (oracle/_ref/libref_shim.so: reference
CSDL, RTS, EmulatedDeadlock and the CubDispatcher's cub::DeviceScan) on a B200:

    gpurun -- python oracle/gen_golden.py gpurun_out/golden      # then copy into tests/golden/

Inputs are the seeded arrays of oracle_init_random_u32 (seed, mask recorded in the file;
seed -1 = the reference's all-ones input).  Small cases store the full output
(out_<algo>_<inc|exc>); large cases store a checksum of checksums (sum_<...> = [sum mod
2^64, xor, last element]).  Reference CSDL / EmulatedDeadlock are only recorded where the
grand total stays below 2^30 (their 30-bit aggregate).
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests import oracle_lib  # noqa: E402

ALGOS = {"csdl": 0, "rts": 1, "emulated": 2, "cub": 3}


def main(outdir: str) -> None:
    os.makedirs(outdir, exist_ok=True)
    oracle = oracle_lib.load()
    shim = oracle_lib.load_ref_shim()
    assert shim is not None and torch.cuda.is_available()
    cases = [  # (name, n, seed, mask, store full output?)
        ("ones_4097", 4097, -1, 0, True),
        ("ones_8192", 8192, -1, 0, True),
        ("rand1023_6143", 6143, 6143, 1023, True),        # the D3D12 protocol's value range
        ("rand1023_65539", 65539, 3, 1023, True),
        ("randfull_524288", 524288, 4, 0xFFFFFFFF, False),  # largest size the reference RTS gets right
        ("randfull_40000", 40000, 5, 0xFFFFFFFF, True),   # totals wrap: RTS / CUB only
        ("rand3_2p24", 1 << 24, 7, 3, False),             # total < 2^30: all four kernels
        ("randfull_2p24", 1 << 24, 9, 0xFFFFFFFF, False),
        ("ones_2p26", 1 << 26, -1, 0, False),
    ]
    for name, n, seed, mask, full in cases:
        x = np.ones(n, dtype=np.uint32) if seed < 0 else oracle.random_u32(n, seed, mask)
        total = int(x.astype(np.uint64).sum())
        rec = {"n": n, "seed": seed, "mask": mask}
        for aname, algo in ALGOS.items():
            if algo in (0, 2) and total >= (1 << 30):
                continue
            if algo == 1 and n > 129 * 4096:
                # the reference's RTS SpineScan is wrong beyond 129 tiles (it adds the total
                # of all warps instead of the preceding warps', ReduceThenScan.cuh:92-93;
                # see tests/test_parity_gpu.py::test_reference_rts_spine_bug)
                continue
            for exclusive in (False, True):
                if algo == 2 and exclusive:
                    continue
                buf = torch.zeros((n + 3) // 4 * 4, dtype=torch.int32, device="cuda")
                buf[:n] = torch.from_numpy(x.view(np.int32)).cuda()
                torch.cuda.synchronize()
                rc = shim.scan_u32(algo, exclusive, buf.data_ptr(), n)
                torch.cuda.synchronize()
                if rc != 0:
                    print(f"  {name}: {aname} refused (rc={rc})")
                    continue
                got = buf[:n].cpu().numpy().view(np.uint32)
                key = f"{aname}_{'exc' if exclusive else 'inc'}"
                if full:
                    rec["out_" + key] = got
                else:
                    rec["sum_" + key] = np.array([int(got.astype(np.uint64).sum() & 0xFFFFFFFFFFFFFFFF),
                                                  int(np.bitwise_xor.reduce(got)), int(got[-1])], dtype=np.uint64)
        np.savez_compressed(os.path.join(outdir, f"ref_cuda_{name}.npz"), **rec)
        print(f"wrote ref_cuda_{name}.npz: {sorted(k for k in rec if k[:4] in ('out_', 'sum_'))}")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "tests", "golden"))
