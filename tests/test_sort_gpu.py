"""GPU parity for the radix-sort consumer: keys bit-exact against the oracle's
stable LSD sort (itself pinned on numpy), pairs in the unique stable order; ragged sizes around
the 10240-key (keys) and 8192-key (pairs) tiles, few distinct keys, already-sorted / reversed / constant inputs, partial bit
ranges, in == out, back-to-back sorts on one ctx (epoch-tagged descriptors, no clearing)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

SIZES = [1, 2, 3, 31, 255, 256, 257, 4095, 4097, 8191, 8192, 8193, 10_239, 10_240, 10_241, 16_385, 20_479, 20_481,
         100_003, (1 << 20) + 7,
         (1 << 22) + 4099]


@pytest.fixture(scope="module")
def gps():
    import gpuprefixsums_b200 as g

    return g


@pytest.fixture(scope="module")
def ps():
    import gpuprefixsums_b200 as gps

    p = gps.PrefixSum(1 << 24)
    yield p
    p.close()


def dev(x):
    return torch.from_numpy(x.view(np.int32)).cuda()


def host(t):
    return t.cpu().numpy().view(np.uint32)


def test_sort_keys_bit_exact(ps, oracle):
    for n in SIZES:
        x = oracle.random_u32(n, seed=300 + n)
        out = ps.sort(dev(x))
        ps.check()
        assert np.array_equal(host(out), oracle.sort_u32(x)), n


def test_sort_pairs_stable(ps, oracle):
    rng = np.random.default_rng(8)
    for n in SIZES:
        # few distinct values in every byte: many ties, stability decides the payload order
        x = (rng.integers(0, 7, size=n, dtype=np.uint64) * 0x01000193 % (1 << 32)).astype(np.uint32)
        v = np.arange(n, dtype=np.uint32)
        ko, vo = ps.sort(dev(x), dev(v))
        ps.check()
        wk, wv = oracle.sort_u32(x, v)
        assert np.array_equal(host(ko), wk) and np.array_equal(host(vo), wv), n
        y = oracle.random_u32(n, seed=500 + n)
        ko, vo = ps.sort(dev(y), dev(v))
        wk, wv = oracle.sort_u32(y, v)
        assert np.array_equal(host(ko), wk) and np.array_equal(host(vo), wv), n


@pytest.mark.parametrize("end_bit", [1, 5, 8, 9, 13, 16, 17, 24, 29, 32])
def test_sort_bit_ranges(ps, gps, oracle, end_bit):
    """A sort on the low end_bit bits is the stable order of key & (2^end_bit - 1) on FULL-RANGE
    keys -- checked against an independent definition (numpy stable argsort of the masked keys),
    the oracle, and cub::DeviceRadixSort with the same end_bit."""
    n = 300_007
    x = oracle.random_u32(n, seed=7)
    v = np.arange(n, dtype=np.uint32)
    ko, vo = ps.sort(dev(x), dev(v), end_bit=end_bit)
    ps.check()
    order = np.argsort(x & np.uint32((1 << end_bit) - 1 if end_bit < 32 else 0xFFFFFFFF), kind="stable")
    assert np.array_equal(host(ko), x[order]) and np.array_equal(host(vo), v[order])
    wk, wv = oracle.sort_u32(x, v, end_bit=end_bit)
    assert np.array_equal(host(ko), wk) and np.array_equal(host(vo), wv)
    ck, cv = torch.empty_like(ko), torch.empty_like(vo)
    gps.cub_sort(dev(x), ck, dev(v), cv, end_bit=end_bit)
    torch.cuda.synchronize()
    assert torch.equal(ck, ko) and torch.equal(cv, vo)
    keys_only = ps.sort(dev(x), end_bit=end_bit)
    ps.check()
    assert np.array_equal(host(keys_only), x[order])


def test_sort_special_inputs(ps, oracle):
    n = 200_001
    cases = {
        "sorted": np.arange(n, dtype=np.uint32) * np.uint32(17),
        "reversed": (np.arange(n, dtype=np.uint32) * np.uint32(17))[::-1].copy(),
        "constant": np.full(n, 0xDEADBEEF, dtype=np.uint32),
        "all_ff": np.full(n, 0xFFFFFFFF, dtype=np.uint32),  # equals the padding key of the last tile
        "zeros": np.zeros(n, dtype=np.uint32),
        "one_digit": (np.arange(n, dtype=np.uint32) % 256) << np.uint32(16),
    }
    v = np.arange(n, dtype=np.uint32)
    for name, x in cases.items():
        ko, vo = ps.sort(dev(x), dev(v))
        ps.check()
        wk, wv = oracle.sort_u32(x, v)
        assert np.array_equal(host(ko), wk) and np.array_equal(host(vo), wv), name


def test_sort_in_place_and_input_preserved(ps, oracle):
    import gpuprefixsums_b200 as gps

    n = 150_000
    x = oracle.random_u32(n, seed=2)
    d = dev(x)
    out = ps.sort(d)
    assert np.array_equal(host(d), x)  # input untouched
    buf = d.clone()
    ps.sort(buf, keys_out=buf)  # 4 passes: in == out allowed
    ps.check()
    assert np.array_equal(host(buf), host(out))
    with pytest.raises(gps.GpsError):
        ps.sort(buf, keys_out=buf, end_bit=8)  # odd number of passes


def test_sort_repeated_on_one_ctx(ps, oracle):
    rng = np.random.default_rng(4)
    for it in range(30):
        n = int(rng.integers(1, 400_000))
        x = oracle.random_u32(n, seed=it)
        assert np.array_equal(host(ps.sort(dev(x))), np.sort(x)), (it, n)
    # interleaved with scans on the same ctx (shared ticket counter)
    x = oracle.random_u32(1 << 20, seed=99)
    want, _ = oracle.scan_u32(x)
    assert np.array_equal(host(ps.scan(dev(x))), want)
    assert np.array_equal(host(ps.sort(dev(x))), np.sort(x))
    ps.check()


def test_sort_2p28_sortedness_and_multiset():
    """BASELINE-size property test: 2^28 random keys come out non-decreasing, with the same
    multiset (sum of keys, sum of squares mod 2^64) -- checked on the device."""
    import gpuprefixsums_b200 as gps

    n = 1 << 28
    big = gps.PrefixSum(1 << 20)
    g = torch.Generator(device="cuda").manual_seed(5)
    x = torch.randint(-(1 << 31), (1 << 31) - 1, (n,), dtype=torch.int32, device="cuda", generator=g)
    out = big.sort(x)
    big.check()
    u = lambda t: t.to(torch.int64) & 0xFFFFFFFF
    ok = True
    for lo in range(0, n, 1 << 26):  # chunked to bound temporary memory
        hi = min(n, lo + (1 << 26) + 1)
        c = u(out[lo:hi])
        ok = ok and bool((c[1:] >= c[:-1]).all())
    assert ok, "not sorted"
    sx, so = u(x), u(out)
    assert int(sx.sum()) == int(so.sum())
    assert int((sx * sx).sum()) == int((so * so).sum())  # wraps mod 2^64 identically
    del sx, so
    big.close()


def test_sort_int32_and_float32_keys(ps, oracle):
    """Signed and floating-point keys (order-preserving maps applied in the first / last pass):
    bit-exact against the oracle incl. -0.0, infinities and NaN bit patterns; pairs stable."""
    rng = np.random.default_rng(12)
    for n in [1, 10_239, 10_241, 300_007]:
        vals = np.arange(n, dtype=np.uint32)
        ki = rng.integers(-2**31, 2**31, size=n, dtype=np.int64).astype(np.int32)
        ko, vo = ps.sort(torch.from_numpy(ki).cuda(), dev(vals), key_type="i32")
        ps.check()
        wk, wv = oracle.sort_keys32(ki, vals, "i32")
        assert np.array_equal(ko.cpu().numpy(), wk) and np.array_equal(host(vo), wv), n
        assert np.array_equal(wk, np.sort(ki))
        bits = rng.integers(0, 2**32, size=n, dtype=np.uint64).astype(np.uint32)  # every float pattern incl. NaN
        kf = bits.view(np.float32).copy()
        if n > 8:
            kf[:6] = [0.0, -0.0, np.inf, -np.inf, 1.5, -1.5]
        ko, vo = ps.sort(torch.from_numpy(kf).cuda(), dev(vals), key_type="f32")
        ps.check()
        wk, wv = oracle.sort_keys32(kf, vals, "f32")
        assert np.array_equal(ko.cpu().numpy().view(np.uint32), wk.view(np.uint32)) and np.array_equal(host(vo), wv), n
        finite = kf[np.isfinite(kf)]
        out_finite = ko.cpu().numpy()
        out_finite = out_finite[np.isfinite(out_finite)]
        assert np.array_equal(out_finite, np.sort(finite))  # numeric order (-0.0 == 0.0 for numpy)
        for kt, k in (("u32", ki.view(np.uint32)), ("i32", ki), ("f32", kf)):  # descending, stable
            kdev = torch.from_numpy(k.view(np.float32 if kt == "f32" else np.int32)).cuda()
            ko, vo = ps.sort(kdev, dev(vals), key_type=kt, descending=True)
            ps.check()
            wk, wv = oracle.sort_keys32(k, vals, kt, descending=True)
            assert np.array_equal(ko.cpu().numpy().view(np.uint32), wk.view(np.uint32)) and np.array_equal(host(vo), wv), (n, kt)
    import gpuprefixsums_b200 as gps
    with pytest.raises(gps.GpsError):
        ps.sort(torch.zeros(8, dtype=torch.float32, device="cuda"), key_type="f32", end_bit=16)


@pytest.mark.parametrize("mask,spin", [(0xFFFFFFFF, 4), (7, 4), (1, 1), (63, 2)])
def test_sort_decoupled_fallback(ps, oracle, mask, spin):
    """The digit look-back with Decoupled Fallback: production variant (no stalls) and emulated stalls
    (stalled tiles are re-counted from the pass's input by a successor and inserted with CAS); keys and
    pairs stay bit-exact / in the unique stable order, int32 and partial bit ranges included."""
    for n in [1, 8191, 8193, 10_240, 10_241, 9 * 10_240 + 5, 300_007, (1 << 21) + 77]:
        x = oracle.random_u32(n, seed=900 + n % 13)
        v = np.arange(n, dtype=np.uint32)
        stats = torch.zeros(4, dtype=torch.int32, device="cuda")
        ko, vo = ps.sort(dev(x), dev(v), fallback=True, stall_mask=mask, max_spin=spin, stats=stats)
        ps.check()
        wk, wv = oracle.sort_u32(x, v)
        assert np.array_equal(host(ko), wk) and np.array_equal(host(vo), wv), (n, mask)
        keys = ps.sort(dev(x), fallback=True, stall_mask=mask, max_spin=spin, end_bit=13)
        ps.check()
        assert np.array_equal(host(keys), oracle.sort_u32(x, end_bit=13)), (n, mask)
        if mask != 0xFFFFFFFF and n > 8 * 10_240:
            st = stats.cpu().numpy()
            assert st[1] > 0 and st[2] > 0, f"no fallbacks recorded: {st}"
    xi = (oracle.random_u32(100_003, seed=5).view(np.int32))
    got = ps.sort(torch.from_numpy(xi).cuda(), key_type="i32", fallback=True, stall_mask=mask, max_spin=spin)
    ps.check()
    assert np.array_equal(got.cpu().numpy(), np.sort(xi, kind="stable"))


# ---- 64-bit keys (gps_sort_keys64): the same kernels instantiated on uint64 keys, 8 digit passes ----
SIZES64 = [1, 2, 31, 257, 4095, 4097, 5119, 5120, 5121, 6143, 6144, 6145, 12_289, 100_003, (1 << 20) + 7, (1 << 22) + 4099]


def dev64(x):
    return torch.from_numpy(np.ascontiguousarray(x).view(np.int64)).cuda()


def host64(t):
    return t.cpu().numpy().view(np.uint64)


def test_sort64_keys_and_pairs_bit_exact(ps, gps, oracle):
    """uint64 keys and (key, 32-bit payload) pairs: bit-exact against the oracle and cub::DeviceRadixSort
    over ragged sizes around the 6144-key (keys) and 5120-key (pairs) tiles; ties decide the payload order."""
    rng = np.random.default_rng(64)
    for n in SIZES64:
        x = rng.integers(0, 2**64, size=n, dtype=np.uint64)
        out = ps.sort(dev64(x), key_type="u64")
        ps.check()
        assert np.array_equal(host64(out), oracle.sort_keys64(x)), n
        v = np.arange(n, dtype=np.uint32)
        ties = (rng.integers(0, 5, size=n, dtype=np.uint64) * np.uint64(0x0100000001000193))  # few values per byte
        for keys in (x, ties):
            ko, vo = ps.sort(dev64(keys), dev(v), key_type="u64")
            ps.check()
            wk, wv = oracle.sort_keys64(keys, v)
            assert np.array_equal(host64(ko), wk) and np.array_equal(host(vo), wv), n
            ck, cv = torch.empty_like(ko), torch.empty_like(vo)
            gps.cub_sort(dev64(keys), ck, dev(v), cv)
            torch.cuda.synchronize()
            assert torch.equal(ck, ko) and torch.equal(cv, vo), n


@pytest.mark.parametrize("end_bit", [1, 8, 13, 32, 33, 40, 47, 56, 61, 64])
def test_sort64_bit_ranges(ps, gps, oracle, end_bit):
    n = 200_003
    x = np.random.default_rng(end_bit).integers(0, 2**64, size=n, dtype=np.uint64)
    v = np.arange(n, dtype=np.uint32)
    ko, vo = ps.sort(dev64(x), dev(v), key_type="u64", end_bit=end_bit)
    ps.check()
    order = np.argsort(x & np.uint64((1 << end_bit) - 1), kind="stable")
    assert np.array_equal(host64(ko), x[order]) and np.array_equal(host(vo), v[order])
    ck, cv = torch.empty_like(ko), torch.empty_like(vo)
    gps.cub_sort(dev64(x), ck, dev(v), cv, end_bit=end_bit)
    torch.cuda.synchronize()
    assert torch.equal(ck, ko) and torch.equal(cv, vo)
    assert np.array_equal(host64(ps.sort(dev64(x), key_type="u64", end_bit=end_bit)), x[order])


def test_sort64_int64_float64_descending(ps, oracle):
    rng = np.random.default_rng(65)
    for n in [1, 5119, 5121, 6145, 300_007]:
        v = np.arange(n, dtype=np.uint32)
        bits = rng.integers(0, 2**64, size=n, dtype=np.uint64)  # every pattern: NaN payloads, -0.0, infinities
        ki, kf = bits.view(np.int64), bits.view(np.float64).copy()
        if n > 8:
            kf[:6] = [0.0, -0.0, np.inf, -np.inf, 1.5, -1.5]
        for kt, k, t in (("u64", bits, torch.int64), ("i64", ki, torch.int64), ("f64", kf, torch.float64)):
            for desc in (False, True):
                kd = torch.from_numpy(np.ascontiguousarray(k)).cuda().view(t) if kt != "u64" else dev64(k)
                ko, vo = ps.sort(kd, dev(v), key_type=kt, descending=desc)
                ps.check()
                wk, wv = oracle.sort_keys64(k, v, kt, descending=desc)
                assert np.array_equal(host64(ko.view(torch.int64)), wk.view(np.uint64)) and np.array_equal(host(vo), wv), (n, kt, desc)
        assert np.array_equal(ps.sort(torch.from_numpy(ki).cuda(), key_type="i64").cpu().numpy(), np.sort(ki))


@pytest.mark.parametrize("mask,spin", [(0xFFFFFFFF, 4), (7, 4), (1, 1)])
def test_sort64_decoupled_fallback(ps, oracle, mask, spin):
    for n in [1, 5121, 6145, 9 * 6144 + 5, 300_007]:
        x = np.random.default_rng(n).integers(0, 2**64, size=n, dtype=np.uint64)
        v = np.arange(n, dtype=np.uint32)
        stats = torch.zeros(4, dtype=torch.int32, device="cuda")
        ko, vo = ps.sort(dev64(x), dev(v), key_type="u64", fallback=True, stall_mask=mask, max_spin=spin, stats=stats)
        ps.check()
        wk, wv = oracle.sort_keys64(x, v)
        assert np.array_equal(host64(ko), wk) and np.array_equal(host(vo), wv), (n, mask)
        keys = ps.sort(dev64(x), key_type="u64", fallback=True, stall_mask=mask, max_spin=spin, end_bit=45)
        ps.check()
        assert np.array_equal(host64(keys), oracle.sort_keys64(x, end_bit=45)), (n, mask)


def test_sort64_2p27_sortedness_and_multiset():
    """1 GiB of uint64 keys: non-decreasing output with the same multiset (sum, sum of squares mod 2^64)."""
    import gpuprefixsums_b200 as gps

    n = 1 << 27
    big = gps.PrefixSum(1 << 20)
    g = torch.Generator(device="cuda").manual_seed(6)
    x = torch.randint(-(1 << 62), (1 << 62), (n,), dtype=torch.int64, device="cuda", generator=g) * 3 + 1  # wraps: all 64 bits vary
    out = big.sort(x, key_type="i64")
    big.check()
    assert bool((out[1:] >= out[:-1]).all()), "not sorted"
    assert int(x.sum()) == int(out.sum()) and int((x * x).sum()) == int((out * out).sum())
    big.close()
