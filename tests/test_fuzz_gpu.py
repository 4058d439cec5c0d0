"""Seeded randomised parity sweep: sizes drawn around the tile boundaries of every kernel shape
(2048 / 4096 / 6144 / 8192 / 9216 keys and the wide / sort tiles) and log-uniformly up to 2^23,
every algorithm, both kinds, in place or not, with a carry -- bit-exact against the oracle."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

ALGOS = ["csdl", "csdl_blocked", "csdl_smem", "csdl_reg", "csdl_tma", "csdl_fallback", "rts"]
TILES = [2048, 4096, 6144, 8192, 9216, 10240, 2304, 4608]


def draw_size(rng) -> int:
    r = rng.random()
    if r < 0.5:
        t = int(rng.choice(TILES)) * int(rng.integers(1, 40))
        return max(0, t + int(rng.integers(-3, 4)))
    return int(2 ** rng.uniform(0, 23))


def test_fuzz_u32(oracle):
    import gpuprefixsums_b200 as gps

    rng = np.random.default_rng(2026)
    ps = gps.PrefixSum(1 << 23)
    for case in range(160):
        n = draw_size(rng)
        algo = ALGOS[case % len(ALGOS)]
        exclusive = bool(rng.integers(0, 2))
        in_place = bool(rng.integers(0, 2))
        carry = int(rng.integers(0, 2**32))
        x = oracle.random_u32(n, seed=1000 + case)
        want, total = oracle.scan_u32(x, exclusive=exclusive, carry=carry)
        d = torch.from_numpy(x.view(np.int32)).cuda()
        tot = torch.zeros(1, dtype=torch.int32, device="cuda")
        out = ps.scan(d, d if in_place else None, algo=algo, exclusive=exclusive, carry=carry, total_out=tot)
        got = out.cpu().numpy().view(np.uint32)
        assert np.array_equal(got, want), (case, n, algo, exclusive, in_place)
        assert int(tot.cpu().numpy().view(np.uint32)[0]) == total, (case, n, algo)
    ps.check()
    ps.close()


def test_fuzz_wide_and_sort(oracle):
    import gpuprefixsums_b200 as gps

    rng = np.random.default_rng(7)
    ps = gps.PrefixSum(1 << 23)
    for case in range(60):
        n = draw_size(rng) // 4
        exclusive = bool(rng.integers(0, 2))
        algo = "csdl" if case % 3 else "rts"
        x = rng.integers(0, 2**63, size=n, dtype=np.uint64) * np.uint64(2) + np.uint64(case & 1)
        want, _ = oracle.scan_u64(x, exclusive=exclusive, carry=case)
        got = ps.scan_wide(torch.from_numpy(x.view(np.int64)).cuda(), algo=algo, exclusive=exclusive, carry=case)
        assert np.array_equal(got.cpu().numpy().view(np.uint64), want), ("u64", case, n, algo)
        y = oracle.random_u32(4 * n, seed=case)
        want4, _ = oracle.scan_u32x4(y, exclusive=exclusive)
        d4 = torch.from_numpy(y.view(np.int32)).cuda().view(-1, 4) if n else torch.empty((0, 4), dtype=torch.int32, device="cuda")
        got4 = ps.scan_wide(d4, algo=algo, exclusive=exclusive, struct4=True)
        assert np.array_equal(got4.cpu().numpy().view(np.uint32).reshape(-1), want4), ("u32x4", case, n, algo)
        m = max(1, draw_size(rng) // 2)
        k = oracle.random_u32(m, seed=50 + case) >> np.uint32(int(rng.integers(0, 28)))  # few distinct high digits
        v = np.arange(m, dtype=np.uint32)
        ko, vo = ps.sort(torch.from_numpy(k.view(np.int32)).cuda(), torch.from_numpy(v.view(np.int32)).cuda(),
                         descending=bool(case & 1))
        wk, wv = oracle.sort_keys32(k, v, "u32", descending=bool(case & 1))
        assert np.array_equal(ko.cpu().numpy().view(np.uint32), wk) and np.array_equal(vo.cpu().numpy().view(np.uint32), wv), ("sort", case, m)
    ps.check()
    ps.close()
