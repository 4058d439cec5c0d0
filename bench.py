#!/usr/bin/env python
"""bench.py -- headline benchmark: uint32 inclusive scan throughput (Gkeys/s) and the
HBM-roofline fraction of the scan kernel, on 1..8 B200s of one node.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
  (N > 1: launched by torchrun, one rank per GPU)

A "step" is one pass of the hot path over one batch of synthetic keys:
  * N == 1: BASELINE.json configs[1] -- uint32 inclusive scan of 2^28 random keys with the
    default CSDL kernel, out of place (1 GiB in + 1 GiB out, both far larger than the
    126 MB L2, so no flush is needed between iterations);
  * N  > 1: the sharded scan -- every rank holds a contiguous slice of 2^28 keys of one
    global array (weak scaling: N * 2^28 keys in total), pass 1 reduce + publish totals over
    NVLink P2P, pass 2 scan with the carry folded into tile 0's look-back.  No NCCL call
    inside the step in the default mode.  The same line also carries BASELINE.json configs[4]
    as `strong_2p33`: the 2^33-key sharded scan (2^33 / N keys per GPU) measured right after.
`value` is whole-job Gkeys/s with inputs resident in HBM; `e2e` is the same metric through the
host-buffer C-ABI call (pinned host memory, H2D + D2H inside the timed region) with its own
roofline: the box's host <-> device copy rate measured live by a memcpy-only probe on the same
GPUs.  `--impl reference` times the CPU oracle restatement of the reference's serial scan on all
host threads (the reference's CUDA directory has no CPU implementation; its kernels are the GPU
baselines reported in `reference_cuda` and by tests/test_compare_baselines_gpu.py); that arm
never loads the product library.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "uint32 inclusive scan throughput"
UNIT = "Gkeys/s"
LOG2_KEYS_PER_GPU = 28
BYTES_PER_KEY = 8  # algorithmic: 4 B read + 4 B written


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=500)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--algo", default="csdl", help="csdl | csdl_blocked | csdl_smem | csdl_reg | csdl_tma | rts | csdl_fallback")
    ap.add_argument("--log2-keys-per-gpu", type=int, default=LOG2_KEYS_PER_GPU)
    ap.add_argument("--log2-total", type=int, default=0, help="N>1: total keys of the MAIN measurement = 2^this (strong scaling)")
    ap.add_argument("--strong-log2-total", type=int, default=33,
                    help="N>1: extra strong-scaling point carried as `strong_2p33` (BASELINE configs[4]); 0 = skip")
    ap.add_argument("--strong-steps", type=int, default=20)
    ap.add_argument("--carry", default="p2p", choices=["p2p", "nccl"], help="N>1 carry exchange")
    ap.add_argument("--e2e-steps", type=int, default=0, help="0 = min(steps, 20)")
    ap.add_argument("--e2e-mode", default="auto", choices=["auto", "ranks", "single"],
                    help="N>1 end-to-end leg: every rank through ShardedScan.scan_host, rank 0 alone through "
                         "gps_sharded_scan_host on all GPUs, or both (auto: report the faster, list both)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--device-only", action="store_true",
                    help="development: only the device-resident measurement (no end-to-end legs, no 2^33 point); "
                         "prints a reduced line")
    ap.add_argument("--option", action="append", default=[], metavar="KEY=VALUE",
                    help="gps_set_option passthrough (tuning), e.g. poll_sleep_ns=100, tile_order=1")
    return ap.parse_args()


# ---------------------------------------------------------------------------------------
# helpers
# ---------------------------------------------------------------------------------------
def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, torch copy)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md, MEASURED_PEAKS.json absent)"


def ncu_traffic(key: str):
    """dram bytes per launch from the committed `ncu --set full` capture
    (profiles/roofline_traffic.json; each entry names the kernel and the commit it was taken at)."""
    try:
        with open(os.path.join(ROOT, "profiles", "roofline_traffic.json")) as f:
            ent = json.load(f).get(key)
    except Exception:
        return None, None
    if isinstance(ent, dict):
        return ent.get("dram_bytes_per_launch"), {k: v for k, v in ent.items() if k != "dram_bytes_per_launch"}
    return ent, None


def workload_config(args, n_gpus: int) -> dict:
    """The `config` object -- identical in both arms (the reference arm times a bounded sample of it)."""
    if n_gpus > 1 and args.log2_total:
        n_total = 1 << args.log2_total
    else:
        n_total = n_gpus << args.log2_keys_per_gpu
    return {"workload": workload_name(args, n_gpus), "keys_total": n_total, "keys_per_gpu": n_total // n_gpus,
            "in_place": False,
            "l2": "inputs (1 GiB in + 1 GiB out per GPU) far exceed the 126 MB L2; no flush needed"}


def workload_name(args, n_gpus: int) -> str:
    if n_gpus == 1:
        return f"uint32 inclusive scan, 2^{args.log2_keys_per_gpu} random keys, 1 B200, {args.algo} (BASELINE configs[1])"
    per_gpu = (f"2^{args.log2_total} keys in total ({n_gpus} slices of 2^{args.log2_total} / {n_gpus})" if args.log2_total
               else f"{n_gpus} x 2^{args.log2_keys_per_gpu} random keys")
    return (f"sharded uint32 inclusive scan, {per_gpu}, contiguous slices, "
            f"carry over {'NVLink P2P stores' if args.carry == 'p2p' else 'NCCL all_gather'} (BASELINE configs[4] shape)")


class ClockSampler:
    """nvidia-smi clocks/throttle reasons while the timed region runs (B200_PROFILING.md)."""

    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
              "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu_index = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.out = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                                          "-lms", "50", "-i", str(self.gpu_index)], stdout=self.out,
                                         stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None
        return self

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.out.close()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in open(self.path):
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                mx.append(float(parts[2]))
            except ValueError:
                continue
            for name, val in zip(names, parts[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.path)
        # samples under load: the upper half (idle samples before/after the region clock down)
        under_load = sorted(sm)[len(sm) // 2:] if sm else []
        return {"sm_mhz": statistics.median(under_load) if under_load else None,
                "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons), "samples": len(sm)}


def bind_to_gpu_cpus(index: int):
    """Pin this process to the CPUs NVML reports as local to GPU `index` BEFORE the pinned host
    buffers of the end-to-end leg are allocated and first touched, so they land on the GPU's
    NUMA node (a remote node costs ~15 % of the PCIe rate on these boxes).  Returns the CPU
    list or None; silently does nothing where NVML or the affinity call is unavailable."""
    try:
        import pynvml

        pynvml.nvmlInit()
        handle = pynvml.nvmlDeviceGetHandleByIndex(index)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(handle, (ncpu + 63) // 64)
        cpus = [i for i in range(ncpu) if (words[i // 64] >> (i % 64)) & 1]
        if cpus and len(cpus) < ncpu:
            os.sched_setaffinity(0, cpus)
            return cpus
        return None
    except Exception:
        return None


def cpu_baseline_serial(oracle, log2n: int) -> dict:
    """The oracle's serial scan (restating the reference's single-thread validation loop,
    Utility.hlsl:99-111) on 1 host core: bounded sample, best of 5 passes over 2^log2n keys."""
    import numpy as np

    n = 1 << log2n
    x = oracle.random_u32(n, seed=1)
    out = np.empty_like(x)
    best = float("inf")
    for _ in range(5):
        t0 = time.perf_counter()
        oracle.scan_u32_mt(x, out, False, 0, 1)
        best = min(best, time.perf_counter() - t0)
    return {"value": n / best / 1e9, "unit": UNIT, "cores": 1, "kind": "port",
            "sample": f"2^{log2n} random uint32 keys, inclusive, best of 5 passes, 1 of {os.cpu_count()} host threads",
            "host_threads_available": os.cpu_count()}


# ---------------------------------------------------------------------------------------
# reference arm: CPU oracle on all host threads.  Never imports gpuprefixsums_b200.
# ---------------------------------------------------------------------------------------
def run_reference(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import numpy as np

    from tests import oracle_lib

    oracle = oracle_lib.load()
    threads = os.cpu_count() or 1
    n = 1 << args.log2_keys_per_gpu  # one GPU's share of the workload per step
    x = oracle.random_u32(n, seed=1)
    out = np.empty_like(x)
    steps, warm = args.steps, args.warmup
    # bounded sample: keep the whole run under ~2 minutes on any host by shrinking the
    # per-step sample (a power of two of the same keys) if one step is too slow
    t0 = time.perf_counter()
    oracle.scan_u32_mt(x, out, False, 0, threads)
    one = time.perf_counter() - t0
    while n > (1 << 20) and one * (steps + warm) > 120.0:
        n >>= 1
        one /= 2
    x, out = x[:n], out[:n]
    for _ in range(warm):
        oracle.scan_u32_mt(x, out, False, 0, threads)
    t0 = time.perf_counter()
    for _ in range(steps):
        oracle.scan_u32_mt(x, out, False, 0, threads)
    dt = (time.perf_counter() - t0) / steps
    assert oracle.check_u32(x, out)[0] == 0, "threaded CPU scan disagrees with the serial oracle"
    value = n / dt / 1e9
    sample = (f"oracle port of the reference's serial scan, chunked over {threads} host threads; "
              f"{steps} steps of {n} random uint32 keys each (one GPU's share is 2^{args.log2_keys_per_gpu})")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": steps, "warmup": warm, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "config": workload_config(args, args.gpus),
        "sample_keys_per_step": n, "host_threads": threads,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    line["reference_cuda"] = reference_cuda_kernels(args.log2_keys_per_gpu)
    print(json.dumps(line), flush=True)


def reference_cuda_kernels(log2n: int):
    """Informational, not the arm's `value`: the reference's OWN CUDA kernels (oracle/_ref/
    libref_shim.so = the reference headers rebuilt for sm_100a, launched with its <8,4> shapes
    and its timing protocol incl. the per-launch descriptor memsets) on this box's GPU, so the
    driver's JSON also carries the GPU-to-GPU comparison.  The shim allocates and fills its own
    buffers (ref_bench_all): the product library is not loaded.  None without shim or GPU."""
    try:
        from tests import oracle_lib

        shim = oracle_lib.load_ref_shim()
        if shim is None or not hasattr(shim.lib, "ref_bench_all"):
            return None
        n = 1 << log2n
        ms = shim.bench_all(n, 50)
        if ms is None:
            return {"unavailable": "ref_bench_all failed (no GPU?)"}
        res = {"unit": UNIT, "keys": n, "input": "hashed keys in [0,4) (reference CSDL truncates totals >= 2^30)",
               "iters": 50}
        for name, t in zip(("csdl", "rts", "cub_device_scan"), ms):
            res[name] = n / (t * 1e-3) / 1e9 if t > 0 else None
        return res
    except Exception as e:  # no GPU / no shim: the CPU line stands on its own
        return {"unavailable": str(e)[:200]}


# ---------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------
def validate_slice(torch, dist, x, out, world, rank, dev):
    """Correctness of what was just timed, on the WHOLE slice (outside the timed region):
    (1) out[i] - out[i-1] == x[i] (mod 2^32) everywhere, (2) out[0] == carry + x[0] where the
    carry is recomputed independently (torch sums of every rank's slice, all_gathered),
    (3) the last output equals carry + slice total.  Together they pin every output."""
    n_local = x.numel()
    chunk = 1 << 26
    total = 0
    for lo_c in range(0, n_local, chunk):
        hi_c = min(n_local, lo_c + chunk)
        if hi_c - lo_c > 1:
            assert torch.equal(out[lo_c + 1:hi_c] - out[lo_c:hi_c - 1], x[lo_c + 1:hi_c]), "difference check failed"
        if lo_c:
            assert int(out[lo_c].item() - out[lo_c - 1].item() - x[lo_c].item()) % (1 << 32) == 0
        total += int(x[lo_c:hi_c].to(torch.int64).bitwise_and(0xFFFFFFFF).sum().item())
    total &= 0xFFFFFFFF
    carry = 0
    if world > 1:
        totals = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
        dist.all_gather(totals, torch.tensor([total], dtype=torch.int64, device=dev))
        carry = sum(int(t.item()) for t in totals[:rank]) & 0xFFFFFFFF
    if n_local:
        assert (int(out[0].item()) - carry - int(x[0].item())) % (1 << 32) == 0, "carry / first key check failed"
        assert (int(out[-1].item()) - carry - total) % (1 << 32) == 0, "slice total check failed"


def validate_host_scan(torch, h_in, h_out, dev):
    """A whole host array against its input: out[0] == in[0] and out[i] - out[i-1] == in[i] (mod
    2^32) for every i, checked on the device chunk by chunk -- together they pin every output."""
    n = h_in.numel()
    chunk = 1 << 26
    prev_last = None
    for lo in range(0, n, chunk):
        hi = min(n, lo + chunk)
        a = h_in[lo:hi].to(dev, non_blocking=True)
        b = h_out[lo:hi].to(dev, non_blocking=True)
        if hi - lo > 1:
            assert torch.equal(b[1:] - b[:-1], a[1:]), f"host scan: difference check failed in [{lo}, {hi})"
        first = int(b[0].item()) - int(a[0].item()) - (prev_last if prev_last is not None else 0)
        assert first % (1 << 32) == 0, f"host scan: chunk boundary check failed at {lo}"
        prev_last = int(b[-1].item())


def run_ours(args) -> None:
    import torch
    import torch.distributed as dist

    import gpuprefixsums_b200 as gps  # raises if the CUDA library is missing: no fallback

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torchrun --nproc-per-node {args.gpus}")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    host_group = None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
        host_group = dist.new_group(backend="gloo")  # host-only barrier: waiting ranks launch nothing on their GPUs

    cfg = workload_config(args, world)
    n_total = cfg["keys_total"]
    lo, hi = gps.slice_bounds(n_total, world, rank)
    n_local = hi - lo

    x = torch.empty(n_local, dtype=torch.int32, device=dev)
    out = torch.empty_like(x)
    gps.init_random(x, seed=1, first_index=lo)  # counter-based: the host never holds the array

    if world == 1:
        ps = gps.PrefixSum(n_local, dev)
        sh = None
        step = lambda: ps.scan(x, out, algo=args.algo)  # noqa: E731
        launches_per_step = 3 if args.algo == "rts" else 1
        checker = ps
    else:
        cap = n_local
        if args.strong_log2_total:
            s_lo, s_hi = gps.slice_bounds(1 << args.strong_log2_total, world, rank)
            cap = max(cap, s_hi - s_lo)
        sh = gps.ShardedScan(cap, mode=args.carry)
        step = lambda: sh.scan(x, out, algo=args.algo)  # noqa: E731
        launches_per_step = (2 + (3 if args.algo == "rts" else 1)) + (1 if args.carry == "nccl" else 0)
        checker = sh.ps

    for kv in args.option:
        key, _, val = kv.partition("=")
        checker.set_option(key, int(val))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v: float) -> float:
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def timed_steps(fn, steps, warm):
        for _ in range(warm):
            fn()
        checker.check()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        checker.check()
        return max_over_ranks(ms) / steps

    # ---- device-resident timing ------------------------------------------------------
    warm = max(args.warmup, 3)
    sampler = ClockSampler(local_rank).start() if rank == 0 else None
    ms_per_step = timed_steps(step, args.steps, warm)
    value = n_total / (ms_per_step * 1e-3) / 1e9
    validate_slice(torch, dist, x, out, world, rank, dev)
    if args.device_only:
        if sampler:
            sampler.stop()
        if rank == 0:
            print(json.dumps({"device_only": True, "n_gpus": world, "value": value, "unit": UNIT, "ms_per_step": ms_per_step,
                              "steps": args.steps, "options": args.option}), flush=True)
        if world > 1:
            dist.barrier()
            sh.close()
            dist.destroy_process_group()
        return

    # ---- end to end through the host-buffer call ----------------------------------------
    e2e_steps = args.e2e_steps or min(args.steps, 20)
    all_cpus = os.sched_getaffinity(0)
    gpu_cpus = bind_to_gpu_cpus(local_rank)
    h_in = torch.empty(n_local, dtype=torch.int32).pin_memory()
    h_in.copy_(x)
    h_out = torch.empty(n_local, dtype=torch.int32).pin_memory()

    def time_e2e(fn, sync, steps):
        for _ in range(2):
            fn()
        sync()
        barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            fn()
        sync()
        barrier()
        return max_over_ranks(time.perf_counter() - t0) / steps

    e2e_legs, e2e_extra = {}, {}
    if world == 1:
        s = time_e2e(lambda: ps.scan_host(h_in, h_out, algo=args.algo, wait=False), ps.host_sync, e2e_steps)
        e2e_legs["gps_scan_host"] = n_total / s / 1e9
        assert torch.equal(h_out, out.cpu()), "host-buffer path disagrees with the device path"
        e2e_path = ("gps_scan_host_async per step + one gps_host_sync (pinned host buffers, chunked H2D/scan/D2H "
                    "pipeline, staging buffers rotate across calls); all of h_out checked against the device result")
    else:
        if args.e2e_mode in ("auto", "ranks"):
            s = time_e2e(lambda: sh.scan_host(h_in, h_out, algo=args.algo), sh.synchronize_host, e2e_steps)
            e2e_legs["ranks_scan_host"] = n_total / s / 1e9
            assert torch.equal(h_out, out.cpu()), "host-buffer path disagrees with the device path"
        if args.e2e_mode in ("auto", "single"):
            # ONE process (rank 0) drives every GPU of the job through the C-ABI multi-GPU host call on
            # ONE host array of n_total keys; the other ranks wait at the barrier with idle GPUs
            barrier()
            if rank == 0:
                big_in = torch.empty(n_total, dtype=torch.int32).pin_memory()
                big_out = torch.empty(n_total, dtype=torch.int32).pin_memory()
                for r in range(world):  # the same global array, generated slice by slice on the device
                    r_lo, r_hi = gps.slice_bounds(n_total, world, r)
                    tmp = torch.empty(r_hi - r_lo, dtype=torch.int32, device=dev)
                    gps.init_random(tmp, seed=1, first_index=r_lo)
                    big_in[r_lo:r_hi].copy_(tmp)
                    del tmp
                hs = gps.ShardedHostScan(list(range(world)))
                # which GPUs carry the host traffic: measured here, memcpy only, on these very buffers
                io_gpus, io_rate = hs.calibrate(big_in, big_out)
                e2e_extra["single_process_io_gpus"] = io_gpus
                e2e_extra["single_process_io_calibrated_gbs_each_way"] = io_rate
                for _ in range(2):
                    hs.scan_host(big_in, big_out, algo=args.algo, wait=False)
                hs.sync()
                t0 = time.perf_counter()
                for _ in range(e2e_steps):
                    hs.scan_host(big_in, big_out, algo=args.algo, wait=False)
                hs.sync()
                s = (time.perf_counter() - t0) / e2e_steps
                e2e_legs["single_process_gps_sharded_scan_host"] = n_total / s / 1e9
                validate_host_scan(torch, big_in, big_out, dev)
                assert torch.equal(big_out[lo:hi], out.cpu()), "single-process host path disagrees with the device path"
                hs.close()
                del big_in, big_out
            dist.barrier(group=host_group)  # the other ranks wait HERE, on the host, with idle GPUs
        e2e_path = ("ranks_scan_host = ShardedScan.scan_host on every rank (pinned H2D, reduce + scan, D2H on three "
                    "streams, copy-in of step k+1 overlapping copy-out of step k); single_process_gps_sharded_scan_host = "
                    "rank 0 alone through the C-ABI call on one host array, chunks dealt round-robin to all GPUs, "
                    "running total chained GPU to GPU over NVLink; value = the faster of the legs that ran")

    # ---- the floor of the end-to-end leg, measured live: memcpy only, same buffers, same GPUs,
    # both directions at once on every rank (the step moves 4 B/key each way)
    s_up, s_down = torch.cuda.Stream(dev), torch.cuda.Stream(dev)

    piece = 1 << 24  # 64 MiB copies, the host pipeline's chunk size (one 1 GiB copy runs ~5 % slower)

    def copy_both():
        with torch.cuda.stream(s_up):
            for o in range(0, n_local, piece):
                x[o:o + piece].copy_(h_in[o:o + piece], non_blocking=True)
        with torch.cuda.stream(s_down):
            for o in range(0, n_local, piece):
                h_out[o:o + piece].copy_(out[o:o + piece], non_blocking=True)

    def copy_sync():
        s_up.synchronize()
        s_down.synchronize()

    torch.cuda.synchronize()
    floor_s = time_e2e(copy_both, copy_sync, 5)
    pcie_each_way = 4 * n_total / floor_s / 1e9  # GB/s per direction, whole job, every GPU carrying an equal share
    pcie_best = max(pcie_each_way, float(e2e_extra.get("single_process_io_calibrated_gbs_each_way", 0.0)))
    clocks = sampler.stop() if sampler else None
    if gpu_cpus:
        os.sched_setaffinity(0, all_cpus)  # the CPU baseline below sees every host core again
    del h_in, h_out

    # ---- BASELINE configs[4]: 2^33 keys in total, strong scaling, on the same ranks ---------
    strong = None
    if world > 1 and args.strong_log2_total and not args.log2_total:
        del x, out
        torch.cuda.empty_cache()
        st_total = 1 << args.strong_log2_total
        s_lo, s_hi = gps.slice_bounds(st_total, world, rank)
        sx = torch.empty(s_hi - s_lo, dtype=torch.int32, device=dev)
        sout = torch.empty_like(sx)
        gps.init_random(sx, seed=1, first_index=s_lo)
        s_ms = timed_steps(lambda: sh.scan(sx, sout, algo=args.algo), args.strong_steps, 3)
        validate_slice(torch, dist, sx, sout, world, rank, dev)
        strong = {"keys_total": st_total, "keys_per_gpu": s_hi - s_lo, "value": st_total / (s_ms * 1e-3) / 1e9,
                  "unit": UNIT, "per_gpu": st_total / world / (s_ms * 1e-3) / 1e9, "ms_per_step": s_ms,
                  "steps": args.strong_steps, "warmup": 3, "scaling": "strong", "validated": "every key of every slice",
                  "hbm_frac_of_measured_at_8B_per_key": BYTES_PER_KEY * (s_hi - s_lo) / (s_ms * 1e-3) / 1e9 / measured_peak()[0]}
        del sx, sout

    if rank == 0:
        peak, peak_src = measured_peak()
        # dominant kernel: the scan kernel.  N == 1 with CSDL: one launch per step, so the
        # event time per step IS the kernel's average launch duration.
        achieved = BYTES_PER_KEY * n_local / (ms_per_step * 1e-3) / 1e9
        traffic, traffic_src = ncu_traffic(args.algo if world == 1 else "sharded")
        roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                    "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                    "algorithmic_bytes_per_launch": BYTES_PER_KEY * n_local,
                    "note": ("per GPU; the sharded step moves 12 B/key (reduce 4 + scan 8) for 8 useful B/key"
                             if world > 1 else "8 B/key x 2^%d keys per launch" % args.log2_keys_per_gpu)}
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            from tests import oracle_lib  # checker only: the CPU baseline leg

            cpu = cpu_baseline_serial(oracle_lib.load(), args.log2_keys_per_gpu)
        e2e_value = max(e2e_legs.values())
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": warm, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "strong" if (world > 1 and args.log2_total) else "weak", "vs_baseline": None,
            "dtype": "u32", "data": "synthetic",
            "config": cfg, "algo": args.algo,
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 4 * n_total,
                    "d2h_bytes_per_step": 4 * n_total, "steps": e2e_steps, "legs": e2e_legs, **e2e_extra,
                    "host_cpus": f"{len(gpu_cpus)} GPU-local of {len(all_cpus)}" if gpu_cpus else f"all {len(all_cpus)}",
                    "path": e2e_path,
                    "roofline": {"bound": "pcie", "achieved": 4.0 * e2e_value,  # 4 B/key each way
                                 "peak": pcie_best, "unit": "GB/s per direction, whole job",
                                 "frac": 4.0 * e2e_value / pcie_best,
                                 "floor_gkeys": pcie_best / 4.0,
                                 "peak_all_gpus_equal_share": pcie_each_way,
                                 "peak_source": "measured live in this run, memcpy only, both directions at once: the better "
                                                "of (a) every rank copying its pinned slice in and a slice out (5 rounds, "
                                                "max over ranks) and (b) the host-I/O calibration of the single-process leg "
                                                "(the GPU subset with the best aggregate rate on this box); "
                                                "profiles/r02_pcie_aggregate.json and r02_pcie_combos.jsonl hold the box-wide probes"}},
            "gpu_launches": launches_per_step * args.steps,
            "roofline": roofline,
            "cpu_baseline": cpu,
        }
        if strong:
            line["strong_2p33"] = strong
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        sh.close()
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
