"""GPU tests of the multi-GPU path.  They run on ONE GPU too: two shards of one array
live on the same device (single-process API: devices [0, 0]; multi-process API: two
gloo ranks sharing cuda:0, peer slots mapped through CUDA IPC), which exercises the same
carry-slot protocol, publish/poll kernels and flow control as a real NVLink pair.  With
two or more GPUs visible the ranks spread over them."""
import ctypes as C
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("algo", ["csdl", "csdl_reg", "rts", "csdl_tma"])
@pytest.mark.parametrize("carry_mode", [0, 1])
def test_single_process_sharded_scan(oracle, algo, carry_mode):
    import gpuprefixsums_b200 as gps
    from gpuprefixsums_b200 import _capi as capi

    ndev = torch.cuda.device_count()
    for G in (2, 3):
        devices = [i % ndev for i in range(G)]
        n_total = 3_000_017
        x = oracle.random_u32(n_total, seed=77)
        bounds = [gps.slice_bounds(n_total, G, r) for r in range(G)]
        d_in, d_out = [], []
        for r, (lo, hi) in enumerate(bounds):
            with torch.cuda.device(devices[r]):
                d_in.append(torch.from_numpy(x[lo:hi].view(np.int32)).cuda())
                d_out.append(torch.zeros(hi - lo, dtype=torch.int32, device="cuda"))
        torch.cuda.synchronize()
        grp = C.c_void_p()
        dev_arr = (C.c_int * G)(*devices)
        capi.check(capi.lib.gps_sharded_create(C.byref(grp), G, dev_arr, max(hi - lo for lo, hi in bounds)))
        try:
            ins = (C.c_void_p * G)(*[t.data_ptr() for t in d_in])
            outs = (C.c_void_p * G)(*[t.data_ptr() for t in d_out])
            ns = (C.c_uint64 * G)(*[hi - lo for lo, hi in bounds])
            ms = C.c_float(0)
            for exclusive in (False, True):
                for _ in range(3):  # repeated calls reuse the double-buffered slots
                    capi.check(capi.lib.gps_sharded_scan(grp, capi.ALGO_NAMES[algo], int(exclusive), capi.U32, ins,
                                                         outs, ns, carry_mode, C.byref(ms)))
                want, _ = oracle.scan_u32(x, exclusive=exclusive)
                got = np.concatenate([t.cpu().numpy().view(np.uint32) for t in d_out])
                assert np.array_equal(got, want), (G, algo, carry_mode, exclusive)
                assert ms.value > 0
        finally:
            capi.lib.gps_sharded_destroy(grp)


def _rank_main_f32(rank: int, world: int, port: int, n_total: int, mode: str, algo: str, q):
    """float32 slices: the carry is summed in slot order by tile 0 / the spine (sharded.py
    carry_from_totals); checked against the float64 serial oracle with the float32 tolerance."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch.distributed as dist

    torch.cuda.set_device(rank % torch.cuda.device_count())
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import gpuprefixsums_b200 as gps
        from tests import oracle_lib

        oracle = oracle_lib.load()
        lo, hi = gps.slice_bounds(n_total, world, rank)
        full = oracle.random_f32(n_total, seed=9)
        x = torch.from_numpy(full[lo:hi].copy()).cuda()
        sh = gps.ShardedScan(hi - lo, mode=mode)
        ok = True
        for step in range(3):
            exclusive = bool(step & 1)
            out = sh.scan(x, algo=algo, exclusive=exclusive)
            sh.check()
            got = np.zeros(n_total, dtype=np.float32)
            got[lo:hi] = out.cpu().numpy()
            # check this rank's slice inside the global array: the oracle compares position by position
            want, _ = oracle.scan_f32(full, exclusive=exclusive)
            bound = np.cumsum(np.abs(full.astype(np.float64)))
            err = np.abs(got[lo:hi].astype(np.float64) - want[lo:hi].astype(np.float64))
            ok = ok and bool(np.all(err <= 2e-5 * bound[lo:hi] + 1e-6))
        h_in = x.cpu().pin_memory()
        h_out = torch.empty_like(h_in).pin_memory()
        sh.scan_host(h_in, h_out, algo=algo)
        sh.synchronize_host()
        want, _ = oracle.scan_f32(full)
        err = np.abs(h_out.numpy().astype(np.float64) - want[lo:hi].astype(np.float64))
        ok = ok and bool(np.all(err <= 2e-5 * bound[lo:hi] + 1e-6))
        dist.barrier()
        sh.close()
        q.put((rank, ok))
    finally:
        dist.destroy_process_group()


def _rank_main(rank: int, world: int, port: int, n_total: int, mode: str, algo: str, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch.distributed as dist

    torch.cuda.set_device(rank % torch.cuda.device_count())
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import gpuprefixsums_b200 as gps
        from tests import oracle_lib

        oracle = oracle_lib.load()
        lo, hi = gps.slice_bounds(n_total, world, rank)
        x = torch.empty(hi - lo, dtype=torch.int32, device="cuda")
        gps.init_random(x, seed=5, first_index=lo)  # every rank generates only its own slice
        sh = gps.ShardedScan(hi - lo, mode=mode)
        ok = True
        for step in range(4):  # several steps: tag / parity rotation of the slots
            out = sh.scan(x, algo=algo, exclusive=bool(step & 1))
            sh.ps.check()
            full = oracle.random_u32(n_total, seed=5)
            want, _ = oracle.scan_u32(full, exclusive=bool(step & 1))
            ok = ok and np.array_equal(out.cpu().numpy().view(np.uint32), want[lo:hi])
        # host buffers in -> host buffers out, asynchronous and pipelined across calls
        h_in = x.cpu().pin_memory()
        h_outs = [torch.empty_like(h_in).pin_memory() for _ in range(5)]
        for k, h_out in enumerate(h_outs):
            sh.scan_host(h_in, h_out, algo=algo, exclusive=bool(k & 1))
        sh.synchronize_host()
        sh.ps.check()
        for k, h_out in enumerate(h_outs):
            want, _ = oracle.scan_u32(full, exclusive=bool(k & 1))
            ok = ok and np.array_equal(h_out.numpy().view(np.uint32), want[lo:hi])
        dist.barrier()
        sh.close()
        q.put((rank, ok))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("mode,algo,world,dtype", [("p2p", "csdl", 2, "u32"), ("p2p", "rts", 2, "u32"),
                                                   ("allgather", "csdl", 2, "u32"), ("p2p", "csdl", 3, "u32"),
                                                   ("p2p", "csdl", 2, "f32"), ("p2p", "rts", 3, "f32"),
                                                   ("allgather", "csdl_fallback", 3, "f32")])
def test_multi_process_sharded_scan(mode, algo, world, dtype):
    import torch.multiprocessing as mp

    n_total = 5_000_011
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    target = _rank_main if dtype == "u32" else _rank_main_f32
    procs = [ctx.Process(target=target, args=(r, world, port, n_total, mode, algo, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok in results), results


# ---------------------------------------------------------------------------------------
# gps_sharded_scan_host: one process, several GPUs, one host array per call
# ---------------------------------------------------------------------------------------
@pytest.mark.parametrize("G", [1, 2, 3, 8])
def test_sharded_scan_host_single_process(oracle, G):
    import gpuprefixsums_b200 as gps

    ndev = torch.cuda.device_count()
    devices = [i % ndev for i in range(G)]
    sh = gps.ShardedHostScan(devices, chunk_keys=1 << 15)  # small chunks: many links in the chain
    try:
        for n in [0, 1, 4097, (1 << 15), (1 << 15) + 1, 1_000_003, (1 << 21) + 5]:
            x = oracle.random_u32(n, seed=300 + n % 97)
            h_in = torch.from_numpy(x.view(np.int32)).pin_memory()
            for algo in ("csdl", "rts", "csdl_fallback"):
                for exclusive in (False, True):
                    h_out = torch.full((n,), -1, dtype=torch.int32).pin_memory()
                    total = torch.zeros(1, dtype=torch.int32).pin_memory()
                    sh.scan_host(h_in, h_out, algo=algo, exclusive=exclusive, carry=7, total_out=total)
                    want, want_total = oracle.scan_u32(x, exclusive=exclusive, carry=7)
                    assert np.array_equal(h_out.numpy().view(np.uint32), want), (G, n, algo, exclusive)
                    assert int(total.item()) & 0xFFFFFFFF == want_total
        # several calls in flight: copies of consecutive calls overlap, results stay separate
        n = 777_777
        xs = [oracle.random_u32(n, seed=40 + i) for i in range(6)]
        h_ins = [torch.from_numpy(x.view(np.int32)).pin_memory() for x in xs]
        h_outs = [torch.empty(n, dtype=torch.int32).pin_memory() for _ in xs]
        for i in range(6):
            sh.scan_host(h_ins[i], h_outs[i], exclusive=bool(i & 1), wait=False)
        sh.sync()
        for i in range(6):
            want, _ = oracle.scan_u32(xs[i], exclusive=bool(i & 1))
            assert np.array_equal(h_outs[i].numpy().view(np.uint32), want), i
        # float32 through the same chain
        xf = oracle.random_f32(1_500_001, seed=8)
        h_out = sh.scan_host(torch.from_numpy(xf).pin_memory(), torch.empty(xf.size, dtype=torch.float32).pin_memory())
        errs, max_rel = oracle.check_f32(xf, h_out.numpy(), rtol=2e-5, atol=1e-6)
        assert errs == 0, max_rel
    finally:
        sh.close()


def test_sharded_scan_host_with_a_late_cta(oracle):
    """A rescued tile (the warp-level Decoupled Fallback of the ticket-free one-wave kernels, here provoked with the
    "debug_delay_tile" hook in every chunk's scan) is informational for the multi-GPU host entry too: the call succeeds,
    the result is exact."""
    import gpuprefixsums_b200 as gps

    ndev = torch.cuda.device_count()
    sh = gps.ShardedHostScan([0, 1 % ndev], chunk_keys=1 << 16)
    try:
        sh.set_option("debug_delay_tile", 4)
        n = (1 << 18) + 4099  # 5 chunks of 32 tiles each
        x = oracle.random_u32(n, seed=17)
        h_in = torch.from_numpy(x.view(np.int32)).pin_memory()
        h_out = torch.full((n,), -1, dtype=torch.int32).pin_memory()
        sh.scan_host(h_in, h_out)  # waits; raises if the group reports an error
        assert np.array_equal(h_out.numpy().view(np.uint32), oracle.scan_u32(x)[0])
        sh.set_option("debug_delay_tile", 0)
    finally:
        sh.close()


def test_sharded_scan_host_io_routing(oracle):
    """The chunks may be routed through a subset of the group's GPUs, in any chain order (what the
    host-I/O calibration decides on boxes with unequal host paths); results do not change."""
    import gpuprefixsums_b200 as gps

    ndev = torch.cuda.device_count()
    G = 4
    sh = gps.ShardedHostScan([i % ndev for i in range(G)], chunk_keys=1 << 15)
    try:
        n = 3_000_001
        x = oracle.random_u32(n, seed=61)
        want, _ = oracle.scan_u32(x)
        h_in = torch.from_numpy(x.view(np.int32)).pin_memory()
        h_out = torch.empty(n, dtype=torch.int32).pin_memory()
        assert sh.host_io() == [0, 1, 2, 3]
        for io in ([2, 0], [3], [1, 3, 0], [0, 1, 2, 3]):
            sh.set_host_io(io)
            assert sh.host_io() == io
            for _ in range(2):
                h_out.fill_(-1)
                sh.scan_host(h_in, h_out)
                assert np.array_equal(h_out.numpy().view(np.uint32), want), io
        with pytest.raises(gps.GpsError):
            sh.set_host_io([1, 1])
        with pytest.raises(gps.GpsError):
            sh.set_host_io([4])
        # calibration: memcpy-only trials on the caller's buffers, then a scan through the winner
        io, rate = sh.calibrate(h_in, h_out)
        assert 1 <= len(io) <= G and rate > 0
        sh.scan_host(h_in, h_out)
        assert np.array_equal(h_out.numpy().view(np.uint32), want), ("after calibration", io)
        # too small to measure: every GPU stays in
        io, rate = sh.calibrate(h_in[:1000], h_out[:1000])
        assert io == [0, 1, 2, 3] and rate == 0
    finally:
        sh.close()


def test_carry_slot_timeout_is_reported_and_contributes_nothing(oracle):
    """A slot whose writer never shows up: the scan gives up after its own (here: tiny) budget,
    adds NOTHING for that slot -- never the mismatched word -- and the error is reported."""
    import gpuprefixsums_b200 as gps
    from gpuprefixsums_b200 import _capi as capi

    n = 100_003
    x = oracle.random_u32(n, seed=4)
    ps = gps.PrefixSum(n)
    ps.set_option("slot_wait_polls", 2000)
    # two slots: slot 0 carries the right tag, slot 1 a stale tag with a poisonous value
    slots = torch.tensor([(5 << 32) | 1000, (4 << 32) | 0xDEAD], dtype=torch.int64, device="cuda")
    for algo in ("csdl", "rts"):
        out = ps.scan_carry_slots(torch.from_numpy(x.view(np.int32)).cuda(), None, slots.data_ptr(), 2, 2, 5, algo=algo)
        with pytest.raises(gps.GpsError) as ei:
            ps.check()
        assert ei.value.code == capi.GPS_ERR_WATCHDOG
        want, _ = oracle.scan_u32(x, carry=1000)  # the stale slot contributed 0
        assert np.array_equal(out.cpu().numpy().view(np.uint32), want), algo
    # the ctx stays usable afterwards
    want, _ = oracle.scan_u32(x)
    assert np.array_equal(ps.scan(torch.from_numpy(x.view(np.int32)).cuda()).cpu().numpy().view(np.uint32), want)
    ps.check()
    ps.close()
