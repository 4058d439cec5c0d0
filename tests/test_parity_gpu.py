"""GPU parity tests: every CUDA scan path, called through the C ABI, against the CPU
oracle (oracle/scan_oracle.c) on the same seeded inputs -- bit-exact for uint32, within
the stated tolerance for float32 -- plus the reference's own all-ones protocol, the
reference's CUDA kernels (oracle/_ref shim) and size-independent properties at 2^28.
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

ALGOS = ["csdl_blocked", "csdl_smem", "csdl_reg", "csdl_tma", "rts", "csdl_fallback"]
# tile sizes in play: 4096 (smem path, small-input register path), 6144 (register path / RTS /
# fallback), 8192 (TMA path), 9216 (blocked path)
EDGE_SIZES = [0, 1, 2, 3, 4, 5, 7, 8, 31, 32, 33, 127, 128, 129, 1023, 1024, 1025, 4095, 4096, 4097, 8191,
              6143, 6144, 6145, 8192, 8193, 9215, 9216, 9217, 12287, 12288, 12289, 16384, 16385, 18431, 18432, 18433,
              65535, 65536, 65537, 100_003, 1 << 20]
# float32 tolerance: |gpu_i - ref_i| <= RTOL * sum_{j<=i}|x_j| + ATOL vs the float64
# serial scan.  Observed max is ~1e-6 (reported by test_float_2p28_tolerance).
F32_RTOL = 2e-5
F32_ATOL = 1e-6


@pytest.fixture(scope="module")
def gps():
    import gpuprefixsums_b200 as g

    return g


@pytest.fixture(scope="module")
def ps(gps):
    p = gps.PrefixSum(1 << 28)
    yield p
    p.close()


def to_dev_u32(x: np.ndarray) -> torch.Tensor:
    return torch.from_numpy(x.view(np.int32)).cuda()


def to_np_u32(t: torch.Tensor) -> np.ndarray:
    return t.cpu().numpy().view(np.uint32)


# ---------------------------------------------------------------------------------------
# uint32: bit-exact against the oracle, ragged / empty / tile-boundary sizes
# ---------------------------------------------------------------------------------------
@pytest.mark.parametrize("algo", ALGOS)
@pytest.mark.parametrize("exclusive", [False, True])
def test_u32_random_full_range_edge_sizes(ps, oracle, algo, exclusive):
    for n in EDGE_SIZES:
        x = oracle.random_u32(n, seed=1000 + n)  # full-range keys: totals wrap mod 2^32
        want, want_total = oracle.scan_u32(x, exclusive=exclusive)
        d_in = to_dev_u32(x)
        d_out = torch.full((n + 8,), 0x5A5A5A5A, dtype=torch.int32, device="cuda")  # canary tail
        total = torch.zeros(1, dtype=torch.int32, device="cuda")
        ps.scan(d_in, d_out[:n], algo=algo, exclusive=exclusive, total_out=total)
        ps.check()
        got = to_np_u32(d_out)
        assert np.array_equal(got[:n], want), f"{algo} n={n} excl={exclusive}"
        assert (got[n:] == 0x5A5A5A5A).all(), f"{algo} n={n}: wrote past the end"
        assert int(to_np_u32(total)[0]) == want_total
        # the input is untouched out of place
        assert np.array_equal(to_np_u32(d_in), x)


@pytest.mark.parametrize("algo", ALGOS)
def test_u32_in_place_and_carry(ps, oracle, algo):
    for n in [1, 4097, 8193, 250_001]:
        x = oracle.random_u32(n, seed=77 + n)
        carry = 0xFFFFFFF0  # forces wraparound immediately
        for exclusive in (False, True):
            want, want_total = oracle.scan_u32(x, exclusive=exclusive, carry=carry)
            buf = to_dev_u32(x)
            total = torch.zeros(1, dtype=torch.int32, device="cuda")
            ps.scan(buf, buf, algo=algo, exclusive=exclusive, carry=carry, total_out=total)
            ps.check()
            assert np.array_equal(to_np_u32(buf), want), f"{algo} in place n={n}"
            assert int(to_np_u32(total)[0]) == want_total


def test_reference_random_protocol_d3d12(ps, oracle):
    """GPUPrefixSumBase.h:135-161: sizes around one/two tiles, seed = size, values & 1023."""
    for algo in ALGOS:
        for n in [3072, 3073, 4100, 6143, 8200]:
            x = oracle.random_d3d12(n, seed=n)
            for exclusive in (False, True):
                d_out = ps.scan(to_dev_u32(x), algo=algo, exclusive=exclusive)
                ps.check()
                assert oracle.check_u32(x, to_np_u32(d_out), exclusive=exclusive)[0] == 0


def test_config0_2p20_random_vs_host_serial(ps, oracle):
    """BASELINE config[0]: uint32 inclusive, 2^20 random keys, CSDL vs the host serial scan."""
    n = 1 << 20
    x = oracle.random_u32(n, seed=1)
    for algo in ["csdl", "csdl_smem", "csdl_reg", "csdl_tma"]:
        got = to_np_u32(ps.scan(to_dev_u32(x), algo=algo))
        ps.check()
        errs, first = oracle.check_u32(x, got)
        assert errs == 0, f"{algo}: {errs} mismatches, first at {first}"


# ---------------------------------------------------------------------------------------
# the reference's own protocol: all-ones, closed form, sizes 4096..8192 (+ around 8192-key tiles)
# ---------------------------------------------------------------------------------------
@pytest.mark.parametrize("algo", ["csdl_blocked", "csdl_smem", "csdl_reg", "csdl_tma", "rts"])
@pytest.mark.parametrize("exclusive", [False, True])
def test_reference_all_ones_sweep(gps, algo, exclusive):
    from gpuprefixsums_b200.dispatchers import (ChainedScanDecoupledLookbackDispatcher,
                                                ReduceThenScanDispatcher)

    d = ReduceThenScanDispatcher(1 << 26) if algo == "rts" else ChainedScanDecoupledLookbackDispatcher(1 << 26, algo)
    sizes = list(range(4096, 8193, 3)) + list(range(8185, 8200)) + list(range(16380, 16390))
    passed, expected = (d.TestAllExclusive if exclusive else d.TestAllInclusive)(sizes=sizes, large=(26,))
    assert passed == expected


def test_emulated_deadlock_all_ones_sweep(gps):
    from gpuprefixsums_b200.dispatchers import EmulatedDeadlockingDispatcher

    d = EmulatedDeadlockingDispatcher(1 << 26)
    # includes the sizes on which the reference kernel hangs: (tiles-1) % 8 == 0, e.g. 9 tiles
    sizes = list(range(4096, 8193, 5)) + [8 * 4096 + 1, 9 * 4096, 16 * 4096 + 5, 17 * 4096, 64 * 4096 + 3]
    passed, expected = d.TestAllInclusive(sizes=sizes, large=(26,))
    assert passed == expected
    stats = d.stats.cpu().numpy()
    assert stats[1] > 0 and stats[2] > 0, f"stalled tiles must have been rescued by fallbacks: {stats}"


def test_emulated_deadlock_random_and_masks(ps, oracle):
    n = (1 << 22) + 77
    x = oracle.random_u32(n, seed=5)
    want, _ = oracle.scan_u32(x)
    for mask, spin in [(7, 4), (1, 1), (63, 8), (511, 2), (0xFFFFFFFF, 4)]:
        stats = torch.zeros(4, dtype=torch.int32, device="cuda")
        buf = to_dev_u32(x)
        ps.scan_emulated_deadlock(buf, buf, stall_mask=mask, max_spin=spin, stats=stats)  # in place
        ps.check()
        assert np.array_equal(to_np_u32(buf), want), f"mask={mask:#x} spin={spin}"
        if mask != 0xFFFFFFFF:
            assert int(stats[2]) > 0  # somebody inserted a recomputed aggregate
    with pytest.raises(Exception):
        ps.scan_emulated_deadlock(to_dev_u32(x), stall_mask=0)  # would stall every tile


# ---------------------------------------------------------------------------------------
# float32: tolerance against the float64 serial oracle
# ---------------------------------------------------------------------------------------
@pytest.mark.parametrize("algo", ALGOS)
@pytest.mark.parametrize("exclusive", [False, True])
def test_f32_against_float64_oracle(ps, oracle, algo, exclusive):
    for n in [1, 5, 4096, 4097, 8193, 100_003, 1 << 22]:
        x = oracle.random_f32(n, seed=31 + n)
        got = ps.scan(torch.from_numpy(x).cuda(), algo=algo, exclusive=exclusive).cpu().numpy()
        ps.check()
        errs, max_rel = oracle.check_f32(x, got, exclusive=exclusive, rtol=F32_RTOL, atol=F32_ATOL)
        assert errs == 0, f"{algo} n={n}: {errs} outside rtol={F32_RTOL}, max rel {max_rel:.3g}"


def test_rts_float_is_deterministic(ps, oracle):
    x = torch.from_numpy(oracle.random_f32(1 << 22, seed=9)).cuda()
    a = ps.scan(x, algo="rts").clone()
    for _ in range(3):
        assert torch.equal(ps.scan(x, algo="rts"), a)


# ---------------------------------------------------------------------------------------
# full size (2^28): closed form, properties, CUB and the reference's CUDA kernels
# ---------------------------------------------------------------------------------------
@pytest.mark.parametrize("algo", ["csdl_blocked", "csdl_smem", "csdl_reg", "csdl_tma", "rts", "csdl_fallback"])
def test_2p28_all_ones_closed_form(gps, ps, algo):
    n = 1 << 28
    buf = torch.empty(n, dtype=torch.int32, device="cuda")
    for exclusive in (False, True):
        gps.init_ones(buf)
        ps.scan(buf, buf, algo=algo, exclusive=exclusive)  # in place, the reference's protocol
        ps.check()
        assert gps.validate_ones(buf, exclusive) == 0


def test_2p28_random_vs_cub_and_properties(gps, ps, oracle):
    n = 1 << 28
    x = torch.empty(n, dtype=torch.int32, device="cuda")
    gps.init_random(x, seed=2024)  # same generator as the oracle
    head = oracle.random_u32(1 << 16, seed=2024)
    assert np.array_equal(to_np_u32(x[: 1 << 16]), head)
    ref = torch.empty_like(x)
    gps.cub_scan(x, ref)  # CUB validated against the oracle on a prefix, then used as device oracle
    assert oracle.check_u32(head, to_np_u32(ref[: 1 << 16]))[0] == 0
    out = torch.empty_like(x)
    # the whole 2^28-key array against the HOST serial oracle, word for word (0.2 s of CPU): the
    # default kernel directly, and CUB -- which then serves as the device-side oracle below
    want, want_total = oracle.scan_u32(to_np_u32(x))
    assert np.array_equal(to_np_u32(ref), want), "cub::DeviceScan vs host serial scan"
    ps.scan(x, out, algo="csdl")
    ps.check()
    assert np.array_equal(to_np_u32(out), want), "default CSDL vs host serial scan, all 2^28 keys"
    del want
    for algo in ["csdl_blocked", "csdl_smem", "csdl_reg", "csdl_tma", "rts", "csdl_fallback"]:
        total = torch.zeros(1, dtype=torch.int32, device="cuda")
        ps.scan(x, out, algo=algo, total_out=total)
        ps.check()
        assert gps.compare_u32(out, ref) == (0, n), algo
        assert int(total.item()) & 0xFFFFFFFF == want_total
        # size-independent properties: out[i] - out[i-1] == x[i] (mod 2^32), total == last
        assert torch.equal(out[1:] - out[:-1], x[1:])
        assert int(total.item()) == int(out[-1].item())
        # exclusive = inclusive - x
        exc = ps.scan(x, algo=algo, exclusive=True)
        ps.check()
        assert torch.equal(exc, out - x)
        del exc


REF_RTS_MAX_KEYS = 129 * 4096  # largest size the reference RTS gets right, see test_reference_rts_spine_bug


def test_reference_cuda_kernels_agree(gps, ps, oracle, ref_shim):
    """The reference's own CUDA scans (unmodified, rebuilt for sm_100a) on the same input.
    Reference CSDL / EmulatedDeadlock pack a 30-bit aggregate, so they are only exact while
    the grand total is below 2^30 (keys in [0,4) at 2^26); CUB is exact mod 2^32; reference
    RTS is only correct up to 129 tiles (its SpineScan bug, next test)."""
    n = 1 << 26
    small = torch.empty(n, dtype=torch.int32, device="cuda")
    gps.init_random(small, seed=7, mask=3)
    full = torch.empty(n, dtype=torch.int32, device="cuda")
    gps.init_random(full, seed=8)
    for exclusive in (False, True):
        mine_small = ps.scan(small, algo="csdl", exclusive=exclusive)
        mine_full = ps.scan(full, algo="csdl", exclusive=exclusive)
        ps.check()
        for algo, data, mine, count in [(ref_shim.CSDL, small, mine_small, n), (ref_shim.CUB, full, mine_full, n),
                                        (ref_shim.RTS, full, mine_full, REF_RTS_MAX_KEYS)]:
            work = data[:count].clone()
            torch.cuda.synchronize()
            assert ref_shim.scan_u32(algo, exclusive, work.data_ptr(), count) == 0
            torch.cuda.synchronize()
            assert gps.compare_u32(mine[:count], work) == (0, count), (algo, exclusive)
    # emulated deadlock (inclusive only in the reference)
    work = small.clone()
    torch.cuda.synchronize()
    assert ref_shim.scan_u32(ref_shim.EMULATED_DEADLOCK, False, work.data_ptr(), n) == 0
    torch.cuda.synchronize()
    mine = ps.scan_emulated_deadlock(small, stall_mask=7, max_spin=4)
    ps.check()
    assert gps.compare_u32(mine, work) == (0, n)


def test_reference_rts_spine_bug(gps, ps, oracle, ref_shim):
    """Documents a defect of the reference found while pinning the oracle: its RTS SpineScan
    adds s_red[WARPS - 1] (the total of ALL warps) instead of the preceding warps' total to
    every warp but the first (ReduceThenScan.cuh:92-93).  Spine entries of tiles >= 128 are
    therefore wrong (entry 128 is only right when it is the last one), and every key from
    tile 129 on comes out wrong: ReduceThenScan fails for n > 129 * 4096 = 528,384 keys.  The reference's own tests never see it: its large cases only
    validate the first 26..28 elements (ReduceThenScanDispatcher.cuh:58-61).  The oracle, the
    reference CSDL, cub::DeviceScan and every kernel of this repo agree with each other."""
    n = 131 * 4096
    x = torch.empty(n, dtype=torch.int32, device="cuda")
    gps.init_ones(x)
    work = x.clone()
    torch.cuda.synchronize()
    assert ref_shim.scan_u32(ref_shim.RTS, False, work.data_ptr(), n) == 0
    torch.cuda.synchronize()
    assert gps.validate_ones(work, False) > 0            # reference RTS: wrong
    assert gps.validate_ones(work[:REF_RTS_MAX_KEYS], False) == 0
    mine = ps.scan(x, algo="rts")
    ps.check()
    assert gps.validate_ones(mine, False) == 0           # ours: right
    work = x.clone()
    torch.cuda.synchronize()
    assert ref_shim.scan_u32(ref_shim.CSDL, False, work.data_ptr(), n) == 0
    torch.cuda.synchronize()
    assert gps.validate_ones(work, False) == 0           # reference CSDL: right


def test_float_2p28_tolerance(gps, ps, oracle):
    """BASELINE config[2]: float32 inclusive, 2^28 values in [0,1), tolerance-checked against
    the float64 serial oracle on the whole array."""
    n = 1 << 28
    x = torch.empty(n, dtype=torch.float32, device="cuda")
    gps.init_random(x, seed=11)
    hx = x.cpu().numpy()
    assert np.array_equal(hx[:4096], oracle.random_f32(4096, seed=11))
    for algo in ["csdl", "rts"]:
        got = ps.scan(x, algo=algo).cpu().numpy()
        ps.check()
        errs, max_rel = oracle.check_f32(hx, got, rtol=F32_RTOL, atol=F32_ATOL)
        print(f"float32 2^28 {algo}: max relative error {max_rel:.3e}")
        assert errs == 0, f"{algo}: {errs} outside tolerance, max rel {max_rel:.3g}"
    # emulated-deadlock path on the float input
    got = ps.scan_emulated_deadlock(x, stall_mask=7, max_spin=4).cpu().numpy()
    ps.check()
    errs, max_rel = oracle.check_f32(hx, got, rtol=F32_RTOL, atol=F32_ATOL)
    assert errs == 0, f"emulated deadlock: max rel {max_rel:.3g}"


# ---------------------------------------------------------------------------------------
# reduce, carry slots, host pipeline, errors
# ---------------------------------------------------------------------------------------
def test_reduce_matches_oracle(ps, oracle):
    for n in [0, 1, 4097, 1 << 20, (1 << 24) + 3]:
        x = oracle.random_u32(n, seed=n + 3)
        _, want = oracle.scan_u32(x)
        got = ps.reduce(to_dev_u32(x) if n else torch.empty(0, dtype=torch.int32, device="cuda"))
        assert int(to_np_u32(got)[0]) == want


@pytest.mark.parametrize("algo", ["csdl_blocked", "csdl_smem", "csdl_reg", "csdl_tma", "rts", "csdl_fallback"])
def test_carry_slots_fold_into_lookback(gps, ps, oracle, algo):
    """Two 'ranks' on one GPU: totals published as tagged slots, tile 0 sums the upstream ones."""
    from gpuprefixsums_b200.sharded import pack_slot

    n = 300_001
    a, b = oracle.random_u32(n, seed=1), oracle.random_u32(n, seed=2)
    slots = torch.zeros(16, dtype=torch.int64, device="cuda")
    ps2 = gps.PrefixSum(1 << 20)
    tag = 5
    ptrs = [slots.data_ptr() + 8 * i for i in range(2)]
    ps.reduce_publish(to_dev_u32(a), [ptrs[0]], tag)
    ps2.reduce_publish(to_dev_u32(b), [ptrs[1]], tag)
    torch.cuda.synchronize()
    words = slots.cpu().numpy().view(np.uint64)
    _, ta = oracle.scan_u32(a)
    _, tb = oracle.scan_u32(b)
    assert int(words[0]) == pack_slot(ta, tag) and int(words[1]) == pack_slot(tb, tag)
    # rank 1: polls both slots, carries only slot 0
    out = ps.scan_carry_slots(to_dev_u32(b), None, slots.data_ptr(), 2, 1, tag, algo=algo)
    ps.check()
    want, _ = oracle.scan_u32(b, carry=ta)
    assert np.array_equal(to_np_u32(out), want)
    ps2.close()


def test_scan_host_pipeline(gps, oracle):
    ps = gps.PrefixSum(1 << 24)
    ps.set_option("host_chunk_keys", 1 << 20)  # many chunks: exercises the device-side chaining
    for n in [0, 5, (1 << 20) - 1, (1 << 22) + 4099]:
        x = oracle.random_u32(n, seed=50 + n)
        h_in = torch.from_numpy(x.view(np.int32)).pin_memory()
        for exclusive in (False, True):
            h_out = ps.scan_host(h_in, exclusive=exclusive, carry=3)
            want, _ = oracle.scan_u32(x, exclusive=exclusive, carry=3)
            assert np.array_equal(h_out.numpy().view(np.uint32), want), (n, exclusive)
    # queued calls (wait=False) of different sizes and kinds share the staging rotation
    cases = []
    for k, n in enumerate([(1 << 20) + 3, 7, 1 << 21, (1 << 22) + 4099, 1 << 20, 12345]):
        x = oracle.random_u32(n, seed=70 + k)
        h_in = torch.from_numpy(x.view(np.int32)).pin_memory()
        h_out = torch.empty_like(h_in).pin_memory()
        ps.scan_host(h_in, h_out, exclusive=bool(k & 1), carry=k, wait=False)
        cases.append((x, h_in, h_out, bool(k & 1), k))
    ps.host_sync()
    ps.check()
    for x, _, h_out, exclusive, carry in cases:
        want, _ = oracle.scan_u32(x, exclusive=exclusive, carry=carry)
        assert np.array_equal(h_out.numpy().view(np.uint32), want), (x.size, exclusive)
    ps.close()


def test_error_behaviour(gps, ps):
    x = torch.zeros(64, dtype=torch.int32, device="cuda")
    with pytest.raises(gps.GpsError) as e:
        ps.scan(x[1:33], x[1:33])  # not 16-byte aligned
    assert e.value.code == -2
    with pytest.raises(gps.GpsError):
        ps.scan(x, algo=99)
    small = gps.PrefixSum(16)
    with pytest.raises(gps.GpsError):
        small.scan(x)  # exceeds ctx capacity
    small.close()
    with pytest.raises(TypeError):
        ps.scan(torch.zeros(8, dtype=torch.float64, device="cuda"))


# ---------------------------------------------------------------------------------------
# 64-bit sizes: more than 2^32 keys on one GPU (the 2-GPU shard of the 2^33-key case)
# ---------------------------------------------------------------------------------------
@pytest.mark.parametrize("algo", ["csdl", "csdl_reg", "rts", "csdl_tma"])
def test_more_than_2p32_keys(gps, algo):
    n = (1 << 32) + 12_345
    free, _ = torch.cuda.mem_get_info()
    if free < 2 * 4 * n + (2 << 30):
        pytest.skip("not enough device memory for 2 x 16 GiB")
    big = gps.PrefixSum(n)
    buf = torch.empty(n, dtype=torch.int32, device="cuda")
    gps.init_ones(buf)
    total = torch.zeros(1, dtype=torch.int32, device="cuda")
    big.scan(buf, buf, algo=algo, carry=7, total_out=total)  # in place, 64-bit indexing throughout
    big.check()
    assert gps.validate_ones(buf, False, first_value=7) == 0  # scan[i] == 7 + i + 1 (mod 2^32)
    assert (int(total.item()) & 0xFFFFFFFF) == (7 + n) & 0xFFFFFFFF
    big.close()
    del buf
    torch.cuda.empty_cache()


# ---------------------------------------------------------------------------------------
# scratch bookkeeping that replaces the reference's per-launch memsets
# ---------------------------------------------------------------------------------------
@pytest.mark.parametrize("algo", ["csdl_blocked", "csdl_smem", "csdl_reg", "csdl_tma", "csdl_fallback"])
def test_epoch_and_ticket_counter_wraparound(gps, oracle, algo):
    """Descriptors are epoch-tagged and the ticket counter is never reset: both must
    survive their wrap (2^30 launches / 2^32 tickets)."""
    ps = gps.PrefixSum(1 << 22)
    n = (1 << 21) + 4099
    x = oracle.random_u32(n, seed=123)
    want, _ = oracle.scan_u32(x)
    d_in = to_dev_u32(x)
    ps.set_option("debug_epoch", (1 << 30) - 4)
    ps.set_option("debug_counter", (1 << 32) - 700)  # a launch claims ~350-500 tickets here
    for i in range(8):
        out = ps.scan(d_in, algo=algo)
        ps.check()
        assert np.array_equal(to_np_u32(out), want), f"launch {i} after the jump"
    ps.close()


def test_single_thread_lookback(gps, oracle):
    """The reference's LookbackSingle (defined but unused there) as a selectable variant."""
    ps = gps.PrefixSum(1 << 24)
    ps.set_option("lookback", 1)
    n = (1 << 23) + 5
    x = oracle.random_u32(n, seed=77)
    want, _ = oracle.scan_u32(x)
    for algo in ("csdl_blocked", "csdl_smem", "csdl_reg", "csdl_tma", "csdl_fallback", "csdl"):
        out = ps.scan(to_dev_u32(x), algo=algo)
        ps.check()
        assert np.array_equal(to_np_u32(out), want), algo
    # descriptors 8 B + partials 16 B per 4096-key tile, wide descriptors 64 B per 36 KB tile, counters;
    # then the shadow descriptors / counter used by launches captured into a CUDA graph
    wide_words = -(-(4 << 28) // (256 * 9 * 16)) * 8
    a256 = lambda v: (v + 255) & ~255
    main = a256(a256(a256((1 << 16) * 8) + (1 << 16) * 16) + wide_words * 8) + 256
    assert gps._capi.lib.gps_scratch_bytes(1 << 28) == a256(a256(main + (1 << 16) * 8) + wide_words * 8) + 256
    ps.close()


@pytest.mark.parametrize("protocol", [0, 1, 2])
@pytest.mark.parametrize("tile_order", [0, 1])
def test_descriptor_protocols_and_tile_order(gps, oracle, protocol, tile_order):
    """atom.exch/ld.relaxed (default), st.release/ld.acquire and st.relaxed/ld.relaxed; ticket
    vs blockIdx tile order -- all bit-exact."""
    ps = gps.PrefixSum(1 << 24)
    ps.set_option("desc_protocol", protocol)
    ps.set_option("tile_order", tile_order)
    n = (1 << 23) + 777
    x = oracle.random_u32(n, seed=9)
    want, _ = oracle.scan_u32(x, exclusive=True)
    # the default kernel (blocked tile) and the Decoupled Fallback kernels included: the options
    # apply to every chained variant (tile_order 1: the register / smem Decoupled Fallback kernels run ticket-free too)
    for algo in ("csdl_blocked", "csdl_smem", "csdl_reg", "csdl_tma", "csdl_fallback", "csdl"):
        out = ps.scan(to_dev_u32(x), algo=algo, exclusive=True)
        ps.check()
        assert np.array_equal(to_np_u32(out), want), (algo, protocol, tile_order)
    for variant in (1, 2, 3):  # every fallback tile, not only the one the size picks
        ps.set_option("fallback_variant", variant)
        out = ps.scan(to_dev_u32(x), algo="csdl_fallback", exclusive=True)
        ps.check()
        assert np.array_equal(to_np_u32(out), want), (variant, protocol, tile_order)
    ps.close()


def test_caller_owned_scratch(gps, oracle):
    """Unity-style caller-owned temp storage (ReduceThenScan.cs:156-279): size
    query, ctx built on the caller's buffer, every algorithm bit-exact, undersized and
    misaligned buffers rejected."""
    n = (1 << 22) + 4099
    need = gps.PrefixSum.scratch_bytes(n)
    scratch = torch.full((need,), 0xA5, dtype=torch.uint8, device="cuda")  # dirty on purpose
    ps = gps.PrefixSum(n, scratch=scratch)
    x = oracle.random_u32(n, seed=12)
    want, _ = oracle.scan_u32(x)
    for algo in ["csdl", "csdl_blocked", "csdl_reg", "csdl_tma", "csdl_smem", "csdl_fallback", "rts"]:
        for _ in range(2):
            out = ps.scan(to_dev_u32(x), algo=algo)
            ps.check()
            assert np.array_equal(to_np_u32(out), want), algo
    x64 = np.random.default_rng(1).integers(0, 2**63, size=n // 2, dtype=np.uint64)
    w64, _ = oracle.scan_u64(x64)
    assert np.array_equal(ps.scan_wide(torch.from_numpy(x64.view(np.int64)).cuda()).cpu().numpy().view(np.uint64), w64)
    ps.close()
    with pytest.raises(gps.GpsError):
        gps.PrefixSum(n, scratch=scratch[: need - 1])
    with pytest.raises(gps.GpsError):
        gps.PrefixSum(n, scratch=torch.empty(need + 256, dtype=torch.uint8, device="cuda")[16:])


@pytest.mark.parametrize("variant", [1, 2, 3])
def test_fallback_variants_forced(gps, oracle, variant):
    """Decoupled Fallback on the register tile (1), the shared-memory-resident tile (2) and the
    blocked shared-memory tile (3),
    each forced at every size: both kinds, uint32 bit-exact / float32 in tolerance, in place,
    with and without emulated stalls (EmulatedDeadlocking.cuh:63-151)."""
    ps = gps.PrefixSum(1 << 23)
    ps.set_option("fallback_variant", variant)
    for n in [1, 4095, 4096, 4097, 6143, 6145, 9215, 9216, 9217, 9 * 4096, 9 * 9216, 100_003, (1 << 22) + 77]:
        x = oracle.random_u32(n, seed=600 + n)
        for exclusive in (False, True):
            want, _ = oracle.scan_u32(x, exclusive=exclusive, carry=11)
            got = ps.scan(to_dev_u32(x), algo="csdl_fallback", exclusive=exclusive, carry=11)
            ps.check()
            assert np.array_equal(to_np_u32(got), want), (n, exclusive)
            want0, _ = oracle.scan_u32(x, exclusive=exclusive)
            for mask, spin in [(7, 4), (1, 1), (63, 2)]:
                stats = torch.zeros(4, dtype=torch.int32, device="cuda")
                buf = to_dev_u32(x)
                ps.scan_emulated_deadlock(buf, buf, exclusive=exclusive, stall_mask=mask, max_spin=spin, stats=stats)
                ps.check()
                assert np.array_equal(to_np_u32(buf), want0), (n, exclusive, mask)
    xf = oracle.random_f32((1 << 22) + 5, seed=3)
    got = ps.scan_emulated_deadlock(torch.from_numpy(xf).cuda(), stall_mask=7, max_spin=4).cpu().numpy()
    ps.check()
    errs, max_rel = oracle.check_f32(xf, got, rtol=F32_RTOL, atol=F32_ATOL)
    assert errs == 0, max_rel
    ps.close()


@pytest.mark.parametrize("variant", [1, 2])
def test_rts_variants_forced(gps, oracle, variant):
    """Reduce-then-Scan with the propagate kernel on the register tile (1) and on the blocked
    shared-memory tile (2), each forced at every size; float32 deterministic in both."""
    ps = gps.PrefixSum(1 << 23)
    ps.set_option("rts_variant", variant)
    for n in [0, 1, 6143, 6144, 6145, 9215, 9216, 9217, 18433, 100_003, (1 << 22) + 77]:
        x = oracle.random_u32(n, seed=700 + n)
        for exclusive in (False, True):
            want, total = oracle.scan_u32(x, exclusive=exclusive, carry=3)
            tot = torch.zeros(1, dtype=torch.int32, device="cuda")
            got = ps.scan(to_dev_u32(x), algo="rts", exclusive=exclusive, carry=3, total_out=tot)
            ps.check()
            assert np.array_equal(to_np_u32(got), want), (n, exclusive)
            assert int(tot.cpu().numpy().view(np.uint32)[0]) == total
            if n:
                buf = to_dev_u32(x)
                ps.scan(buf, buf, algo="rts", exclusive=exclusive, carry=3)  # in place
                assert np.array_equal(to_np_u32(buf), want), ("in place", n, exclusive)
    xf = torch.from_numpy(oracle.random_f32((1 << 22) + 5, seed=8)).cuda()
    a = ps.scan(xf, algo="rts").clone()
    assert torch.equal(ps.scan(xf, algo="rts"), a)
    errs, max_rel = oracle.check_f32(xf.cpu().numpy(), a.cpu().numpy(), rtol=F32_RTOL, atol=F32_ATOL)
    assert errs == 0, max_rel
    ps.close()


def test_one_ctx_on_several_streams(gps, oracle):
    """Consecutive launches of one ctx on DIFFERENT streams (no synchronisation by the caller):
    the library orders them, so the shared ticket counter / descriptors stay consistent.
    (Found by compute-sanitizer: a sort on the default stream followed by the host pipeline
    on its own streams used to race on the ticket counter.)"""
    ps = gps.PrefixSum(1 << 22)
    n = (1 << 21) + 77
    x = oracle.random_u32(n, seed=91)
    want, _ = oracle.scan_u32(x)
    d = to_dev_u32(x)
    streams = [torch.cuda.Stream() for _ in range(3)]
    torch.cuda.synchronize()
    outs = []
    for it in range(12):
        st = streams[it % 3]
        with torch.cuda.stream(st):
            outs.append(ps.scan(d, algo=["csdl_reg", "csdl_smem", "rts", "csdl_blocked"][it % 4], stream=st))
            if it % 4 == 1:
                outs.append(ps.sort(d))  # current stream = st
    h_in = torch.from_numpy(x.view(np.int32)).pin_memory()
    h_out = ps.scan_host(h_in, wait=False)  # the pipeline's own streams
    ps.host_sync()
    torch.cuda.synchronize()
    ps.check()
    for o in outs:
        if o.numel() == n and not np.array_equal(to_np_u32(o), want):
            assert np.array_equal(to_np_u32(o), np.sort(x))
    assert np.array_equal(h_out.numpy().view(np.uint32), want)
    ps.close()


def test_cuda_graph_capture(gps, oracle):
    """Launches captured into a CUDA graph are replay-safe: Reduce-then-Scan keeps no launch
    state at all; the chained kernels, the wide scans and the sort run on a shadow scratch that
    the captured sequence clears first (constant epochs, zero ticket base).  Replays with new
    inputs stay bit-exact and plain launches on the same ctx are unaffected in between."""
    ps = gps.PrefixSum(1 << 22)
    n = (1 << 21) + 9
    d_in = torch.zeros(n, dtype=torch.int32, device="cuda")
    outs = {a: torch.zeros_like(d_in) for a in ["rts", "csdl", "csdl_reg", "csdl_blocked", "csdl_fallback"]}
    d64 = torch.zeros(n // 2, dtype=torch.int64, device="cuda")
    out64 = torch.zeros_like(d64)
    skeys = torch.zeros_like(d_in)
    for a, o in outs.items():  # warm-up outside the capture: function attributes, lazy scratch
        ps.scan(d_in, o, algo=a)
    ps.scan_wide(d64, out64)
    ps.sort(d_in, keys_out=skeys)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for k, (a, o) in enumerate(outs.items()):
            ps.scan(d_in, o, algo=a, exclusive=bool(k & 1), carry=5)
        ps.scan_wide(d64, out64, carry=9)
        ps.sort(d_in, keys_out=skeys)
    for seed in (1, 2, 3):
        x = oracle.random_u32(n, seed=seed)
        d_in.copy_(to_dev_u32(x))
        d64.copy_(d_in[: 2 * (n // 2)].view(torch.int64))
        g.replay()
        torch.cuda.synchronize()
        for k, (a, o) in enumerate(outs.items()):
            want, _ = oracle.scan_u32(x, exclusive=bool(k & 1), carry=5)
            assert np.array_equal(to_np_u32(o), want), (seed, a)
        w64, _ = oracle.scan_u64(x[: 2 * (n // 2)].view(np.uint64), carry=9)
        assert np.array_equal(out64.cpu().numpy().view(np.uint64), w64), seed
        assert np.array_equal(to_np_u32(skeys), np.sort(x)), seed
        # a plain launch between replays: host-side counters were not disturbed by the capture
        plain = ps.scan(d_in, algo="csdl_smem")
        want, _ = oracle.scan_u32(x)
        assert np.array_equal(to_np_u32(plain), want), seed
    ps.check()
    ps.close()


def test_occupancy_estimation_variant(gps, oracle):
    """The reference's csdldf_occ test variant: persistent CTAs drain the tickets, the ones whose first
    claim was a real tile count themselves; the scan stays bit-exact and later scans on the ctx are
    unaffected (ticket accounting: 2 x tiles claims)."""
    ps = gps.PrefixSum(1 << 24)
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    for n in [1, 6143, 6145, 100_003, (1 << 24) - 5]:
        x = oracle.random_u32(n, seed=n % 1000)
        for exclusive in (False, True):
            out, occ = ps.estimate_occupancy(to_dev_u32(x), exclusive=exclusive)
            ps.check()
            want, _ = oracle.scan_u32(x, exclusive=exclusive)
            assert np.array_equal(to_np_u32(out), want), (n, exclusive)
            tiles = -(-n // 6144)
            assert 1 <= occ <= tiles
            if tiles > 64 * sms:  # many more tiles than the device can hold: the count is the residency
                assert sms <= occ <= 16 * sms, occ
        got = ps.scan(to_dev_u32(x))  # the never-reset ticket counter is still in step
        ps.check()
        assert np.array_equal(to_np_u32(got), oracle.scan_u32(x)[0])
    ps.close()


def test_ticket_free_decoupled_fallback_default(gps, oracle):
    """GPS_ALGO_CSDL's mid-size kernels (round 2): Decoupled Fallback with tile = blockIdx.x, no ticket
    (2^19 < n <= 2^21 on the register tile, 2^22 < n <= 2^24 on the smem tile; "ticket_free" = 0 restores the
    plain ticketed kernels).  Both settings bit-exact on the range boundaries, ragged sizes, in place,
    exclusive, with a carry and a total, float32 in tolerance, and 40 launches back to back (epochs, no ticket
    counter to advance) interleaved with ticketed launches of other sizes."""
    ps = gps.PrefixSum((1 << 25) + 4096)
    sizes = [1, 1000, 2049, (1 << 17), (1 << 17) + 1, (1 << 18) + 33,  # one wave of the plain register tile: blockIdx order too
             (1 << 19), (1 << 19) + 1, (1 << 20) + 4099, 1 << 21, (1 << 21) + 1, 1 << 22, (1 << 22) + 1,
             3_800_000, 4_500_001, 1332 * 4096, 1332 * 4096 + 1,  # one wave of the <128,8> register tile, and just above it
             (1 << 23) + 777, (1 << 24) - 5, 1 << 24, (1 << 24) + 1, (1 << 25) + 1]
    for n in sizes:
        x = oracle.random_u32(n, seed=n % 1000)
        d = to_dev_u32(x)
        for tf in (1, 2, 0):  # 1: ticket-free with the warp-level rescue (default), 2: lean ticket-free kernels, 0: tickets
            ps.set_option("ticket_free", tf)
            for exclusive in (False, True):
                want, total = oracle.scan_u32(x, exclusive=exclusive, carry=11)
                tot = torch.zeros(1, dtype=torch.int32, device="cuda")
                out = ps.scan(d, exclusive=exclusive, carry=11, total_out=tot)
                ps.check()
                assert np.array_equal(to_np_u32(out), want), (n, tf, exclusive)
                assert int(tot.cpu().numpy().view(np.uint32)[0]) == total, (n, tf, exclusive)
            buf = d.clone()
            ps.scan(buf, buf)  # in place
            ps.check()
            assert np.array_equal(to_np_u32(buf), oracle.scan_u32(x)[0]), (n, tf, "in place")
    ps.set_option("ticket_free", 1)
    xf = oracle.random_f32((1 << 23) + 5, seed=3)
    outf = ps.scan(torch.from_numpy(xf).cuda()).cpu().numpy()
    ps.check()
    assert oracle.check_f32(xf, outf, rtol=2e-5, atol=1e-6)[0] == 0
    # back to back, sizes of both ticket-free ranges interleaved with ticketed kernels; every result checked
    rng = np.random.default_rng(5)
    cases = [(1 << 20) + 17, (1 << 23) + 9, (1 << 22), 100_003, (1 << 21), (1 << 24)]
    xs = {n: oracle.random_u32(n, seed=n % 77) for n in cases}
    ds = {n: to_dev_u32(xs[n]) for n in cases}
    outs = []
    for it in range(40):
        n = cases[int(rng.integers(0, len(cases)))]
        outs.append((n, ps.scan(ds[n], exclusive=bool(it & 1)), bool(it & 1)))
    ps.check()
    wants = {(n, e): oracle.scan_u32(xs[n], exclusive=e)[0] for n in cases for e in (False, True)}
    for n, o, e in outs:
        assert np.array_equal(to_np_u32(o), wants[(n, e)]), (n, e)
    ps.close()


def test_ticket_free_fallback_under_emulated_stalls(gps, oracle):
    """tile_order = 1 makes the emulated-deadlock harness run ticket-free as well: stalled tiles are rescued by
    their successors' fallbacks exactly as with tickets (register, smem and blocked tiles)."""
    ps = gps.PrefixSum(1 << 23)
    ps.set_option("tile_order", 1)
    for variant in (1, 2, 3):
        ps.set_option("fallback_variant", variant)
        for n in [(1 << 20) + 4099, (1 << 22) + 5] + ([(1 << 23) - 11] if variant == 3 else []):
            x = oracle.random_u32(n, seed=n % 91)
            for exclusive in (False, True):
                stats = torch.zeros(4, dtype=torch.int32, device="cuda")
                out = ps.scan_emulated_deadlock(to_dev_u32(x), exclusive=exclusive, stall_mask=7, max_spin=4, stats=stats)
                ps.check()
                assert np.array_equal(to_np_u32(out), oracle.scan_u32(x, exclusive=exclusive)[0]), (variant, n, exclusive)
                st = stats.cpu().numpy()
                assert st[1] > 0 and st[2] > 0, f"no fallbacks recorded: {st}"
    ps.close()


def test_ticket_free_scans_concurrently_on_two_streams(gps, oracle):
    """Two contexts on two streams, their ticket-free scans (one-wave register tiles, Decoupled Fallback smem tile) in flight at
    the same time: together the one-wave grids exceed what the device holds, so CTAs of one grid wait for slots the other grid
    occupies.  The Decoupled Fallback kernels need nothing for that; the one-wave kernels finish because CTAs are dispatched in
    index order (include/gps_b200.h, "ticket_free").  Every result is checked."""
    a, b = gps.PrefixSum(1 << 24), gps.PrefixSum(1 << 24)
    sa, sb = torch.cuda.Stream(), torch.cuda.Stream()
    sizes = [(1 << 21) + 77, 3_500_000, 5_000_000, (1 << 23) + 5, 1 << 20, 3_600_000]
    xs = {n: oracle.random_u32(n, seed=n % 101) for n in sizes}
    ds = {n: to_dev_u32(xs[n]) for n in sizes}
    wants = {n: oracle.scan_u32(xs[n])[0] for n in sizes}
    torch.cuda.synchronize()
    outs = []
    for rep in range(6):
        for i, n in enumerate(sizes):
            m = sizes[(i + 3) % len(sizes)]
            with torch.cuda.stream(sa):
                oa = a.scan(ds[n], stream=sa)
            with torch.cuda.stream(sb):
                ob = b.scan(ds[m], stream=sb)
            outs.append((n, oa))
            outs.append((m, ob))
    torch.cuda.synchronize()
    a.check()
    b.check()
    for n, o in outs:
        assert np.array_equal(to_np_u32(o), wants[n]), n
    a.close()
    b.close()


def test_late_cta_is_rescued_by_the_one_wave_kernels(gps, oracle):
    """What out-of-order dispatch would look like: the CTA of one tile starts ~40 ms late ("debug_delay_tile").  The
    ticket-free one-wave kernels (tile = blockIdx.x) do not wait for it: after ~10 ms of polling, a successor's look-back warp
    re-reduces that tile from the input and inserts its aggregate with CAS (warp-level Decoupled Fallback) -- results bit-exact,
    "rescued" reported.  With tickets ("ticket_free" = 0) and in place the scan simply waits: exact too, nothing rescued."""
    ps = gps.PrefixSum(1 << 23)
    for n, late in [((1 << 15) + 3, 3), ((1 << 20) + 4099, 17), (4_200_000, 1), (1 << 21, 300), ((1 << 20), 0)]:
        x = oracle.random_u32(n, seed=n % 977)
        d = to_dev_u32(x)
        for exclusive in (False, True):
            want, total = oracle.scan_u32(x, exclusive=exclusive, carry=5)
            ps.set_option("ticket_free", 1)
            ps.set_option("debug_delay_tile", late + 1)
            ps.get_option("rescued")  # clear
            tot = torch.zeros(1, dtype=torch.int32, device="cuda")
            out = ps.scan(d, exclusive=exclusive, carry=5, total_out=tot)
            ps.check()
            assert np.array_equal(to_np_u32(out), want), (n, late, exclusive)
            assert int(tot.cpu().numpy().view(np.uint32)[0]) == total
            assert ps.get_option("rescued") == 1, (n, late)
            assert ps.get_option("rescued") == 0  # reading clears it
            # in place: no rescue (the owner's stores could race with a re-read), the scan waits for the late CTA
            buf = d.clone()
            ps.scan(buf, buf, exclusive=exclusive, carry=5)
            ps.check()
            assert np.array_equal(to_np_u32(buf), want), (n, late, exclusive, "in place")
            assert ps.get_option("rescued") == 0
            # tickets: the late CTA holds a later ticket than... no: it holds ITS ticket and everybody behind it waits
            ps.set_option("ticket_free", 0)
            out = ps.scan(d, exclusive=exclusive, carry=5)
            ps.check()
            assert np.array_equal(to_np_u32(out), want), (n, late, exclusive, "tickets")
            assert ps.get_option("rescued") == 0
    # float32: the rescuer's sum and the owner's differ in rounding at most
    ps.set_option("ticket_free", 1)
    ps.set_option("debug_delay_tile", 9)
    xf = oracle.random_f32((1 << 20) + 77, seed=8)
    outf = ps.scan(torch.from_numpy(xf).cuda()).cpu().numpy()
    ps.check()
    assert oracle.check_f32(xf, outf, rtol=2e-5, atol=1e-6)[0] == 0
    assert ps.get_option("rescued") == 1
    ps.set_option("debug_delay_tile", 0)
    ps.close()
