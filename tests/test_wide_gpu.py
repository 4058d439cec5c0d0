"""GPU parity for the wide element types: uint64 and the four-lane struct
scan bit-exact against the oracle, float64 within 1e-13 * sum|x| of the long-double serial
scan; ragged sizes around the 1024/2048-element tiles, in place, carry and total."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

SIZES = [0, 1, 2, 3, 5, 1023, 1024, 1025, 2047, 2048, 2049, 4097, 100_003, (1 << 22) + 7]
F64_RTOL = 1e-13


@pytest.fixture(scope="module")
def ps():
    import gpuprefixsums_b200 as gps

    p = gps.PrefixSum(1 << 26)
    yield p
    p.close()


@pytest.mark.parametrize("exclusive", [False, True])
def test_u64_bit_exact(ps, oracle, exclusive):
    rng = np.random.default_rng(5)
    for n in SIZES:
        x = rng.integers(0, 2**63, size=n, dtype=np.uint64) * np.uint64(2) + rng.integers(0, 2, size=n, dtype=np.uint64)
        carry = 0xFFFFFFFFFFFFFF00
        want, want_total = oracle.scan_u64(x, exclusive=exclusive, carry=carry)
        d_in = torch.from_numpy(x.view(np.int64)).cuda()
        total = torch.zeros(1, dtype=torch.int64, device="cuda")
        out = ps.scan_wide(d_in, exclusive=exclusive, carry=carry, total_out=total)
        ps.check()
        assert np.array_equal(out.cpu().numpy().view(np.uint64), want), n
        assert int(total.cpu().numpy().view(np.uint64)[0]) == want_total
        buf = d_in.clone()
        ps.scan_wide(buf, buf, exclusive=exclusive, carry=carry)  # in place
        ps.check()
        assert np.array_equal(buf.cpu().numpy().view(np.uint64), want), f"in place {n}"


@pytest.mark.parametrize("exclusive", [False, True])
def test_struct_scan_bit_exact(ps, oracle, exclusive):
    for n in SIZES:
        x = oracle.random_u32(4 * n, seed=40 + n)
        want, want_total = oracle.scan_u32x4(x, exclusive=exclusive, carry=(9, 8, 7, 0xFFFFFFFF))
        d_in = torch.from_numpy(x.view(np.int32)).cuda().view(-1, 4) if n else torch.empty((0, 4), dtype=torch.int32, device="cuda")
        total = torch.zeros(4, dtype=torch.int32, device="cuda")
        out = ps.scan_wide(d_in, exclusive=exclusive, struct4=True, carry=(9, 8, 7, 0xFFFFFFFF), total_out=total)
        ps.check()
        assert np.array_equal(out.cpu().numpy().view(np.uint32).reshape(-1), want), n
        assert np.array_equal(total.cpu().numpy().view(np.uint32), want_total)


def test_struct_scan_reference_closed_form(ps):
    """validate.wgsl:50-60: all-ones input, out[i] == (i >> 2) + 1 over the flat u32 view."""
    n = (1 << 23) + 3
    ones = torch.ones((n, 4), dtype=torch.int32, device="cuda")
    out = ps.scan_wide(ones, struct4=True).view(-1)
    ps.check()
    want = (torch.arange(4 * n, device="cuda", dtype=torch.int64) >> 2) + 1
    assert torch.equal(out.to(torch.int64), want)


@pytest.mark.parametrize("exclusive", [False, True])
def test_f64_tolerance_and_determinism(ps, oracle, exclusive):
    rng = np.random.default_rng(17)
    for n in [1, 1025, 100_003, (1 << 22) + 7]:
        x = rng.random(n) - 0.25
        d_in = torch.from_numpy(x).cuda()
        out = ps.scan_wide(d_in, exclusive=exclusive, carry=1.5)
        ps.check()
        errs, max_rel = oracle.check_f64(x, out.cpu().numpy(), exclusive=exclusive, carry=1.5, rtol=F64_RTOL)
        assert errs == 0, f"n={n}: max rel {max_rel:.3g}"
        assert torch.equal(ps.scan_wide(d_in, exclusive=exclusive, carry=1.5), out)  # RTS: deterministic


def test_wide_throughput_report(ps, oracle):
    """Not a gate: prints Gelem/s and GB/s of the wide scans at 2^26 bytes-equivalent sizes."""
    import gpuprefixsums_b200 as gps

    big = gps.PrefixSum(1 << 28)
    res = {}
    for name, t in [("u64", torch.ones(1 << 27, dtype=torch.int64, device="cuda")),
                    ("f64", torch.ones(1 << 27, dtype=torch.float64, device="cuda")),
                    ("u32x4", torch.ones((1 << 26, 4), dtype=torch.int32, device="cuda"))]:
        out = torch.empty_like(t)
        kw = {"struct4": True} if name == "u32x4" else {}
        for _ in range(3):
            big.scan_wide(t, out, **kw)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            big.scan_wide(t, out, **kw)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        nbytes = t.numel() * t.element_size()
        res[name] = {"ms": ms, "GB_s_useful": 2 * nbytes / ms / 1e6, "Gelem_s": (t.shape[0]) / ms / 1e6}
    big.close()
    print(res)
