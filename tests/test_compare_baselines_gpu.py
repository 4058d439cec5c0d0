"""Side-by-side throughput of the new kernels, the reference's own CUDA kernels (sm_100a
rebuild, oracle/_ref shim), cub::DeviceScan and the host serial scan, at 2^28 uint32 keys
with the reference's timing protocol.  Asserts only sanity (everything ran, the default
path beats the reference CSDL); the numbers go to gpurun_out/compare.json and are copied
into profiles/ and DESIGN.md by hand."""
import json
import os
import time

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _time_ours(ps, gps, algo, buf, batch, exclusive=False, emulated=False):
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    total = 0.0
    for i in range(batch + 1):
        gps.init_ones(buf)
        torch.cuda.synchronize()
        start.record()
        if emulated:
            ps.scan_emulated_deadlock(buf, buf, stall_mask=7, max_spin=4)
        else:
            ps.scan(buf, buf, algo=algo, exclusive=exclusive)
        stop.record()
        stop.synchronize()
        if i:
            total += start.elapsed_time(stop)
    ps.check()
    return total / batch


def test_compare_at_2p28(oracle, ref_shim):
    import gpuprefixsums_b200 as gps

    n, batch = 1 << 28, 50
    ps = gps.PrefixSum(n)
    buf = torch.empty(n, dtype=torch.int32, device="cuda")
    res = {"n": n, "protocol": "reference BatchTimingInclusive: all-ones, in place, events around one dispatch, "
                               f"mean of {batch} after one discarded iteration"}

    def rate(ms):
        return n / (ms * 1e-3) / 1e9

    for name, algo in [("ours_csdl (default: blocked smem-resident tile)", "csdl"), ("ours_csdl_reg", "csdl_reg"),
                       ("ours_csdl_tma", "csdl_tma"),
                       ("ours_rts", "rts"), ("ours_csdl_fallback (no stalls)", "csdl_fallback")]:
        res[name] = rate(_time_ours(ps, gps, algo, buf, batch))
    res["ours_csdl exclusive"] = rate(_time_ours(ps, gps, "csdl", buf, batch, exclusive=True))
    res["ours_emulated_deadlock (mask 7, spin 4)"] = rate(_time_ours(ps, gps, None, buf, batch, emulated=True))
    fbuf = torch.empty(n, dtype=torch.float32, device="cuda")
    res["ours_csdl float32"] = rate(_time_ours(ps, gps, "csdl", fbuf, batch))
    res["ours_rts float32"] = rate(_time_ours(ps, gps, "rts", fbuf, batch))
    del fbuf
    torch.cuda.synchronize()
    for name, algo in [("reference_csdl", ref_shim.CSDL), ("reference_rts (wrong beyond 129 tiles)", ref_shim.RTS),
                       ("reference_emulated_deadlock", ref_shim.EMULATED_DEADLOCK), ("cub_device_scan", ref_shim.CUB)]:
        ms = ref_shim.time_u32(algo, False, buf.data_ptr(), None, n, batch)
        assert ms > 0
        res[name] = rate(ms)
    # host serial scan (the oracle) on one core
    x = oracle.random_u32(n, seed=1)
    out = np.empty_like(x)
    best = min(_timed(lambda: oracle.scan_u32_mt(x, out, False, 0, 1)) for _ in range(3))
    res["host_serial_scan (1 core)"] = n / best / 1e9
    res["host_threads_available"] = os.cpu_count()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "compare.json"), "w") as f:
        json.dump(res, f, indent=1)
    print(json.dumps(res, indent=1))
    assert res["ours_csdl (default: blocked smem-resident tile)"] > res["reference_csdl"]
    assert res["ours_csdl (default: blocked smem-resident tile)"] > res["cub_device_scan"]
    ps.close()


def _timed(fn):
    t0 = time.perf_counter()
    fn()
    return time.perf_counter() - t0
