"""GPU parity for the wide element types: uint64 and the four-lane struct
scan bit-exact against the oracle, float64 within 1e-13 * sum|x| of the long-double serial
scan; both algorithms (chained scan and Reduce-then-Scan); ragged sizes around the
2304/4608-element tiles, in place, carry and total."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

SIZES = [0, 1, 2, 3, 5, 1023, 1025, 2303, 2304, 2305, 4607, 4608, 4609, 9217, 100_003, (1 << 22) + 7]
ALGOS = ["csdl", "rts", "csdl_fallback"]
F64_RTOL = 1e-13


@pytest.fixture(scope="module")
def ps():
    import gpuprefixsums_b200 as gps

    p = gps.PrefixSum(1 << 26)
    yield p
    p.close()


@pytest.mark.parametrize("algo", ALGOS)
@pytest.mark.parametrize("exclusive", [False, True])
def test_u64_bit_exact(ps, oracle, exclusive, algo):
    rng = np.random.default_rng(5)
    for n in SIZES:
        x = rng.integers(0, 2**63, size=n, dtype=np.uint64) * np.uint64(2) + rng.integers(0, 2, size=n, dtype=np.uint64)
        carry = 0xFFFFFFFFFFFFFF00
        want, want_total = oracle.scan_u64(x, exclusive=exclusive, carry=carry)
        d_in = torch.from_numpy(x.view(np.int64)).cuda()
        total = torch.zeros(1, dtype=torch.int64, device="cuda")
        out = ps.scan_wide(d_in, algo=algo, exclusive=exclusive, carry=carry, total_out=total)
        ps.check()
        assert np.array_equal(out.cpu().numpy().view(np.uint64), want), n
        assert int(total.cpu().numpy().view(np.uint64)[0]) == want_total
        buf = d_in.clone()
        ps.scan_wide(buf, buf, algo=algo, exclusive=exclusive, carry=carry)  # in place
        ps.check()
        assert np.array_equal(buf.cpu().numpy().view(np.uint64), want), f"in place {n}"


@pytest.mark.parametrize("algo", ALGOS)
@pytest.mark.parametrize("exclusive", [False, True])
def test_struct_scan_bit_exact(ps, oracle, exclusive, algo):
    for n in SIZES:
        x = oracle.random_u32(4 * n, seed=40 + n)
        want, want_total = oracle.scan_u32x4(x, exclusive=exclusive, carry=(9, 8, 7, 0xFFFFFFFF))
        d_in = torch.from_numpy(x.view(np.int32)).cuda().view(-1, 4) if n else torch.empty((0, 4), dtype=torch.int32, device="cuda")
        total = torch.zeros(4, dtype=torch.int32, device="cuda")
        out = ps.scan_wide(d_in, algo=algo, exclusive=exclusive, struct4=True, carry=(9, 8, 7, 0xFFFFFFFF), total_out=total)
        ps.check()
        assert np.array_equal(out.cpu().numpy().view(np.uint32).reshape(-1), want), n
        assert np.array_equal(total.cpu().numpy().view(np.uint32), want_total)


@pytest.mark.parametrize("algo", ALGOS)
def test_struct_scan_reference_closed_form(ps, algo):
    """validate.wgsl:50-60: all-ones input, out[i] == (i >> 2) + 1 over the flat u32 view."""
    n = (1 << 23) + 3
    ones = torch.ones((n, 4), dtype=torch.int32, device="cuda")
    out = ps.scan_wide(ones, algo=algo, struct4=True).view(-1)
    ps.check()
    want = (torch.arange(4 * n, device="cuda", dtype=torch.int64) >> 2) + 1
    assert torch.equal(out.to(torch.int64), want)


@pytest.mark.parametrize("algo", ALGOS)
@pytest.mark.parametrize("exclusive", [False, True])
def test_f64_tolerance_and_determinism(ps, oracle, exclusive, algo):
    rng = np.random.default_rng(17)
    for n in [1, 1025, 100_003, (1 << 22) + 7]:
        x = rng.random(n) - 0.25
        d_in = torch.from_numpy(x).cuda()
        out = ps.scan_wide(d_in, algo=algo, exclusive=exclusive, carry=1.5)
        ps.check()
        errs, max_rel = oracle.check_f64(x, out.cpu().numpy(), exclusive=exclusive, carry=1.5, rtol=F64_RTOL)
        assert errs == 0, f"n={n}: max rel {max_rel:.3g}"
        if algo == "rts":  # fixed summation order; the chained scan's depends on look-back depth
            assert torch.equal(ps.scan_wide(d_in, algo=algo, exclusive=exclusive, carry=1.5), out)


def test_wide_many_launches_reuse_descriptors(ps, oracle):
    """Epoch tags: back-to-back launches of different sizes and types on one ctx never see
    each other's descriptors (no memset between them)."""
    rng = np.random.default_rng(23)
    for it in range(40):
        n = int(rng.integers(1, 300_000))
        x = rng.integers(0, 2**63, size=n, dtype=np.uint64)
        want, _ = oracle.scan_u64(x, exclusive=bool(it & 1))
        out = ps.scan_wide(torch.from_numpy(x.view(np.int64)).cuda(), exclusive=bool(it & 1))
        assert np.array_equal(out.cpu().numpy().view(np.uint64), want), (it, n)
        y = oracle.random_u32(4 * n, seed=it)
        want4, _ = oracle.scan_u32x4(y)
        out4 = ps.scan_wide(torch.from_numpy(y.view(np.int32)).cuda().view(-1, 4), struct4=True)
        assert np.array_equal(out4.cpu().numpy().view(np.uint32).reshape(-1), want4), (it, n)
    ps.check()


def test_wide_large_1gib(oracle):
    """1 GiB per type (the byte volume of BASELINE configs[1]): all-ones closed forms on the
    device for uint64 / struct scan, random float64 against a float64 torch.cumsum peer."""
    import gpuprefixsums_b200 as gps

    big = gps.PrefixSum(1 << 28)
    n = 1 << 27
    for algo in ALGOS:
        ones = torch.ones(n, dtype=torch.int64, device="cuda")
        out = big.scan_wide(ones, algo=algo, carry=(1 << 40))
        big.check()
        assert torch.equal(out, torch.arange(1, n + 1, device="cuda", dtype=torch.int64) + (1 << 40)), algo
        del ones, out
        ones4 = torch.ones((n // 2, 4), dtype=torch.int32, device="cuda")
        out4 = big.scan_wide(ones4, algo=algo, struct4=True, exclusive=True).view(-1)
        big.check()
        assert torch.equal(out4.to(torch.int64), torch.arange(2 * n, device="cuda", dtype=torch.int64) >> 2), algo
        del ones4, out4
        g = torch.Generator(device="cuda").manual_seed(3)
        x = torch.rand(n, dtype=torch.float64, device="cuda", generator=g) - 0.25
        out = big.scan_wide(x, algo=algo)
        big.check()
        peer = torch.cumsum(x, 0)
        bound = F64_RTOL * torch.cumsum(x.abs(), 0)
        assert bool(((out - peer).abs() <= 2 * bound).all()), algo
        del x, out, peer, bound
    big.close()


@pytest.mark.parametrize("algo", ALGOS)
@pytest.mark.parametrize("exclusive", [False, True])
def test_blocked_kernel_at_32_bits(ps, oracle, exclusive, algo):
    """The blocked-tile kernel instantiated for uint32 / float32 (comparison builds): same
    oracle, same tolerance as the production 32-bit kernels."""
    for n in [0, 1, 9215, 9216, 9217, 100_003, (1 << 22) + 7]:
        x = oracle.random_u32(n, seed=90 + n)
        out = ps.scan_wide(torch.from_numpy(x.view(np.int32)).cuda(), algo=algo, exclusive=exclusive, carry=77)
        ps.check()
        errs, first = oracle.check_u32(x, out.cpu().numpy().view(np.uint32), exclusive=exclusive, carry=77)
        assert errs == 0, (n, first)
    xf = oracle.random_f32((1 << 22) + 7, seed=4)
    outf = ps.scan_wide(torch.from_numpy(xf).cuda(), algo=algo, exclusive=exclusive)
    ps.check()
    errs, max_rel = oracle.check_f32(xf, outf.cpu().numpy(), exclusive=exclusive, rtol=2e-5, atol=1e-6)
    assert errs == 0, max_rel


# ---------------------------------------------------------------------------------------
# Decoupled Fallback for the wide scans (the reference's struct scan IS a fallback kernel:
# GPUPrefixSumsWebGPUapis/SharedShaders/csdldf_struct.wgsl:61-316)
# ---------------------------------------------------------------------------------------
@pytest.mark.parametrize("mask,spin", [(7, 4), (1, 1), (63, 2), (0xFFFFFFFF, 1)])
def test_wide_emulated_deadlock_bit_exact(ps, oracle, mask, spin):
    """Stalled tiles are rescued by their successors (whole-CTA re-reduction from the input, per-word
    CAS insertion); results stay bit-exact for uint64 and the struct scan, both kinds, in place too,
    and the counters show that fallbacks really happened."""
    rng = np.random.default_rng(11)
    for n in [1, 2303, 2304, 2305, 4608, 9 * 2304 + 1, 100_003, (1 << 21) + 7]:
        x = rng.integers(0, 2**63, size=n, dtype=np.uint64)
        for exclusive in (False, True):
            want, _ = oracle.scan_u64(x, exclusive=exclusive, carry=5)
            stats = torch.zeros(4, dtype=torch.int32, device="cuda")
            buf = torch.from_numpy(x.view(np.int64)).cuda()
            ps.scan_wide_emulated_deadlock(buf, buf, exclusive=exclusive, carry=5, stall_mask=mask, max_spin=spin, stats=stats)
            ps.check()
            assert np.array_equal(buf.cpu().numpy().view(np.uint64), want), (n, exclusive, mask)
            if mask != 0xFFFFFFFF and n > 4 * 4608:
                st = stats.cpu().numpy()
                assert st[1] > 0 and st[2] > 0, f"no fallbacks recorded: {st}"
        xs = oracle.random_u32(4 * (n // 2 + 1), seed=n)
        want, _ = oracle.scan_u32x4(xs, carry=(1, 2, 3, 4))
        out = ps.scan_wide_emulated_deadlock(torch.from_numpy(xs.view(np.int32)).cuda().view(-1, 4), struct4=True,
                                             carry=(1, 2, 3, 4), stall_mask=mask, max_spin=spin)
        ps.check()
        assert np.array_equal(out.cpu().numpy().view(np.uint32).reshape(-1), want), (n, mask)


def test_wide_fallback_float64_and_large_closed_form(ps, oracle):
    """float64 under emulated stalls: a rescuer reproduces the owner's association order (its aggregate
    is bit-identical to the owner's, so descriptor words written by either can mix); results are within
    the float64 tolerance of the serial oracle.  Plus the all-ones closed form over 7282 tiles."""
    rng = np.random.default_rng(3)
    n = (1 << 21) + 123
    x = rng.random(n) * 2.0 - 0.5
    d_in = torch.from_numpy(x).cuda()
    for mask, spin in [(7, 4), (1, 1)]:
        out = ps.scan_wide_emulated_deadlock(d_in, stall_mask=mask, max_spin=spin)
        ps.check()
        errs, max_rel = oracle.check_f64(x, out.cpu().numpy(), rtol=F64_RTOL)
        assert errs == 0, (mask, max_rel)
    # all-ones closed form through the fallback kernel at 1 GiB-scale tile counts
    ones = torch.ones(1 << 24, dtype=torch.int64, device="cuda")
    out = ps.scan_wide_emulated_deadlock(ones, stall_mask=7, max_spin=4)
    ps.check()
    assert torch.equal(out, torch.arange(1, (1 << 24) + 1, dtype=torch.int64, device="cuda"))
