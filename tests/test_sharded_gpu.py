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


@pytest.mark.parametrize("mode,algo", [("p2p", "csdl"), ("p2p", "rts"), ("allgather", "csdl")])
def test_multi_process_sharded_scan(mode, algo):
    import torch.multiprocessing as mp

    world, n_total = 2, 5_000_011
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_rank_main, args=(r, world, port, n_total, mode, algo, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok in results), results
