"""Sharded scan across GPUs, one process per GPU (``torch.distributed``).

New capability (the reference is single-GPU).  Rank g holds the
contiguous keys [g*n/G, (g+1)*n/G).  Per call:

  pass 1  ``gps_reduce_publish``    reduce the slice (4 B/key) and publish the tagged
                                    total to every rank's carry slots
  pass 2  ``gps_scan_carry_slots``  scan the slice (8 B/key); tile 0 (CSDL) or the spine
                                    (RTS) polls all G slots and seeds itself with the sum
                                    of the first `rank` totals -- the carry is applied in
                                    the same pass, no extra sweep over HBM

Carry exchange:
  * ``"p2p"``  the fold kernel stores {total|tag} straight into every peer's slot array
               over NVLink (peer memory mapped once with CUDA IPC); no host
               synchronisation and no collective call per step;
  * ``"nccl"`` a G-element ``all_gather`` of the tagged word (also used on the gloo/CPU
               test path for the host logic).
Slots are double-buffered by tag parity; since every rank polls ALL G slots, no rank can
finish step k before every rank has published step k, so a slot is never overwritten
while a peer still waits for its previous value.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch
import torch.distributed as dist

MAX_RANKS = 16  # gps::kMaxPublish


def carry_from_totals(totals_bits: list[int], rank: int, dtype: torch.dtype = torch.int32):
    """Host restatement of what tile 0 computes from the slots: the exclusive carry of
    `rank` (uint32 wraparound, or float32 summed in slot order).  Used by the CPU
    (gloo) tests of the multi-rank logic and by the host-staged debug mode."""
    if dtype == torch.float32:
        vals = torch.tensor(totals_bits[:rank], dtype=torch.int64).to(torch.int32).view(torch.float32)
        return float(vals.sum(dtype=torch.float32)) if rank else 0.0
    acc = 0
    for v in totals_bits[:rank]:
        acc = (acc + (v & 0xFFFFFFFF)) & 0xFFFFFFFF
    return acc


def pack_slot(value_bits: int, tag: int) -> int:
    return ((tag & 0xFFFFFFFF) << 32) | (value_bits & 0xFFFFFFFF)


def unpack_slot(word: int) -> tuple[int, int]:
    word &= 0xFFFFFFFFFFFFFFFF
    return word & 0xFFFFFFFF, word >> 32


def slice_bounds(n_total: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous slice of rank: sizes differ by at most one 4-key vector; every slice
    start is a multiple of 4 keys so device buffers stay 16-byte aligned."""
    vec = (n_total + 3) // 4
    base, rem = divmod(vec, world)
    start_vec = rank * base + min(rank, rem)
    end_vec = start_vec + base + (1 if rank < rem else 0)
    return min(start_vec * 4, n_total), min(end_vec * 4, n_total)


class ShardedScan:
    """Collective sharded scan; construct on every rank of ``group``."""

    def __init__(self, max_n_local: int, mode: str = "p2p", group=None):
        from . import _capi as capi
        from .scan import PrefixSum

        if mode not in ("p2p", "nccl", "allgather"):
            raise ValueError("mode must be 'p2p' or 'nccl' (alias 'allgather')")
        if mode == "allgather":
            mode = "nccl"
        self.capi = capi
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        if self.world > MAX_RANKS:
            raise ValueError(f"at most {MAX_RANKS} ranks")
        self.mode = mode
        self.device = torch.device("cuda", torch.cuda.current_device())
        self.ps = PrefixSum(max_n_local, self.device)
        self.tag = 0
        # 2 parities x MAX_RANKS tagged 64-bit slots, allocated with cudaMalloc so it can
        # be exported over CUDA IPC
        self._slots = C.c_void_p()
        capi.check(capi.lib.gps_malloc(C.byref(self._slots), 2 * MAX_RANKS * 8))
        capi.check(capi.lib.gps_memset(self._slots, 0, 2 * MAX_RANKS * 8, None))
        capi.check(capi.lib.gps_device_sync())
        self._peer_slots: list[int] = [0] * self.world
        self._opened: list[int] = []
        self._word = torch.zeros(1, dtype=torch.int64, device=self.device)
        self._gathered = torch.zeros(self.world, dtype=torch.int64, device=self.device)
        if mode == "p2p":
            self._exchange_ipc()

    def _exchange_ipc(self) -> None:
        capi = self.capi
        handle = C.create_string_buffer(capi.IPC_HANDLE_BYTES)
        capi.check(capi.lib.gps_ipc_export(self._slots, handle))
        every: list = [None] * self.world
        dist.all_gather_object(every, bytes(handle.raw), group=self.group)  # works on nccl and gloo
        for r in range(self.world):
            if r == self.rank:
                self._peer_slots[r] = self._slots.value
                continue
            ptr = C.c_void_p()
            capi.check(capi.lib.gps_ipc_open(every[r], C.byref(ptr)))
            self._peer_slots[r] = ptr.value
            self._opened.append(ptr.value)
        dist.barrier(group=self.group)

    def close(self) -> None:
        capi = self.capi
        for p in self._opened:
            capi.lib.gps_ipc_close(p)
        self._opened = []
        if self._slots.value:
            capi.lib.gps_free(self._slots)
            self._slots = C.c_void_p()
        self.ps.close()

    def scan(self, inp: torch.Tensor, out: Optional[torch.Tensor] = None, *, algo="csdl",
             exclusive: bool = False) -> torch.Tensor:
        """Scan this rank's slice in the context of the global array."""
        self.tag = (self.tag % 0xFFFFFFFE) + 1  # 1 .. 2^32-2, never 0
        tag, par = self.tag, self.tag & 1
        base = self._slots.value + par * MAX_RANKS * 8
        if self.mode == "p2p":
            dst = [self._peer_slots[r] + (par * MAX_RANKS + self.rank) * 8 for r in range(self.world)]
            self.ps.reduce_publish(inp, dst, tag)
        else:
            # publish into my own staging word, all_gather the words, then copy them into
            # the slot array the scan kernel polls
            self.ps.reduce_publish(inp, [self._word.data_ptr()], tag)
            if dist.get_backend(self.group) == "nccl":
                dist.all_gather_into_tensor(self._gathered, self._word, group=self.group)
            else:  # gloo (tests): stage the G words through the host
                words = [torch.zeros(1, dtype=torch.int64) for _ in range(self.world)]
                dist.all_gather(words, self._word.cpu(), group=self.group)
                self._gathered.copy_(torch.cat(words))
            self.capi.check(self.capi.lib.gps_memcpy_d2d(base, self._gathered.data_ptr(), self.world * 8,
                                                         torch.cuda.current_stream().cuda_stream))
        return self.ps.scan_carry_slots(inp, out, base, self.world, self.rank, tag, algo=algo,
                                        exclusive=exclusive)

    # ---- host buffers in, host buffers out, pipelined across calls ---------------------
    def scan_host(self, h_in: torch.Tensor, h_out: torch.Tensor, *, algo="csdl", exclusive: bool = False) -> None:
        """This rank's slice from (pinned) host memory to (pinned) host memory, ASYNCHRONOUS:
        H2D, reduce + scan, and D2H run on three streams over two device buffer pairs, so
        the copy-in of call k+1 overlaps the copy-out of call k (a scan needs its whole
        slice -- and every upstream total -- before it can write anything, so inside one call
        the two directions cannot overlap).  Call ``synchronize_host()`` before reading
        ``h_out`` or re-using ``h_in``.  The slot protocol is unchanged: all compute of a rank
        stays in order on one stream."""
        n = h_in.numel()
        if h_in.is_cuda or h_out.is_cuda or h_out.numel() != n:
            raise ValueError("scan_host takes two CPU tensors of equal size")
        if getattr(self, "_hp", None) is None or self._hp["x"][0].numel() < n or self._hp["x"][0].dtype != h_in.dtype:
            # earlier calls may still be copying into / scanning the old staging tensors on the side
            # streams: drain them before the caching allocator may hand those blocks out again
            self.synchronize_host()
            mk = lambda: torch.empty(n, dtype=h_in.dtype, device=self.device)  # noqa: E731
            ev = lambda: [torch.cuda.Event(), torch.cuda.Event()]  # noqa: E731
            self._hp = {"x": [mk(), mk()], "out": [mk(), mk()], "k": 0,
                        "s_h2d": torch.cuda.Stream(self.device), "s_cmp": torch.cuda.Stream(self.device),
                        "s_d2h": torch.cuda.Stream(self.device), "ev_h2d": ev(), "ev_cmp": ev(), "ev_d2h": ev()}
        hp = self._hp
        b = hp["k"] & 1
        first = hp["k"] < 2
        hp["k"] += 1
        x, out = hp["x"][b][:n], hp["out"][b][:n]
        with torch.cuda.stream(hp["s_h2d"]):
            if not first:
                hp["s_h2d"].wait_event(hp["ev_cmp"][b])  # x[b] is free once the scan two calls ago is done
            x.copy_(h_in, non_blocking=True)
            hp["ev_h2d"][b].record(hp["s_h2d"])
        with torch.cuda.stream(hp["s_cmp"]):
            hp["s_cmp"].wait_event(hp["ev_h2d"][b])
            if not first:
                hp["s_cmp"].wait_event(hp["ev_d2h"][b])  # out[b] is free once it went back to the host
            self.scan(x, out, algo=algo, exclusive=exclusive)
            hp["ev_cmp"][b].record(hp["s_cmp"])
        with torch.cuda.stream(hp["s_d2h"]):
            hp["s_d2h"].wait_event(hp["ev_cmp"][b])
            h_out.copy_(out, non_blocking=True)
            hp["ev_d2h"][b].record(hp["s_d2h"])

    def synchronize_host(self) -> None:
        """Wait for every queued scan_host call; raises GpsError(GPS_ERR_WATCHDOG) if a look-back or
        a carry slot of a peer never arrived (the outputs are then invalid)."""
        hp = getattr(self, "_hp", None)
        if hp is not None:
            for key in ("s_h2d", "s_cmp", "s_d2h"):
                hp[key].synchronize()
            self.ps.check(hp["s_cmp"])

    def check(self) -> None:
        """Synchronise the current stream and raise if a scan of this rank gave up waiting."""
        self.ps.check()


class ShardedHostScan:
    """ONE process, several GPUs, one host array per call (``gps_sharded_scan_host``): the array is
    cut into chunks dealt round-robin to the GPUs, every GPU pipelines copy-in / scan / copy-out
    and the chunks' scans chain their running total from GPU to GPU over NVLink.  The multi-GPU
    form of ``PrefixSum.scan_host``; what a C++ host binds (include/gps_b200.h)."""

    def __init__(self, devices: list[int], chunk_keys: int = 1 << 24):
        from . import _capi as capi

        self.capi = capi
        self.devices = list(devices)
        self._grp = C.c_void_p()
        arr = (C.c_int * len(self.devices))(*self.devices)
        capi.check(capi.lib.gps_sharded_create(C.byref(self._grp), len(self.devices), arr, int(chunk_keys)))
        self.set_option("host_chunk_keys", int(chunk_keys))

    def set_option(self, key: str, value: int) -> None:
        for i in range(len(self.devices)):
            ctx = self.capi.lib.gps_sharded_ctx(self._grp, i)
            self.capi.check(self.capi.lib.gps_set_option(ctx, key.encode(), int(value)))

    def scan_host(self, h_in: torch.Tensor, h_out: torch.Tensor, *, algo="csdl", exclusive: bool = False, carry=0,
                  total_out: Optional[torch.Tensor] = None, wait: bool = True) -> torch.Tensor:
        from .scan import _algo, _dtype_code, _kind, carry_bits

        if h_in.is_cuda or h_out.is_cuda or not h_in.is_contiguous() or not h_out.is_contiguous():
            raise ValueError("scan_host takes contiguous CPU tensors")
        if h_out.numel() != h_in.numel() or h_out.dtype != h_in.dtype:
            raise ValueError("h_out must match h_in")
        fn = self.capi.lib.gps_sharded_scan_host if wait else self.capi.lib.gps_sharded_scan_host_async
        self.capi.check(fn(self._grp, _algo(algo), _kind(exclusive), _dtype_code(h_in), h_in.data_ptr(), h_out.data_ptr(),
                           h_in.numel(), carry_bits(carry, h_in.dtype),
                           total_out.data_ptr() if total_out is not None else None))
        return h_out

    def sync(self) -> None:
        self.capi.check(self.capi.lib.gps_sharded_host_sync(self._grp))

    def calibrate(self, h_src: torch.Tensor, h_dst: torch.Tensor) -> tuple[list[int], float]:
        """Measure (memcpy only, on these pinned buffers; h_dst is overwritten with junk) which GPUs
        should carry the host traffic and route the chunks through them.  Returns (GPU indices into
        the group, GB/s per direction)."""
        rate = C.c_float(0)
        nbytes = min(h_src.numel() * h_src.element_size(), h_dst.numel() * h_dst.element_size())
        self.capi.check(self.capi.lib.gps_sharded_calibrate_host(self._grp, h_src.data_ptr(), h_dst.data_ptr(), nbytes,
                                                                 C.byref(rate)))
        return self.host_io(), float(rate.value)

    def host_io(self) -> list[int]:
        idx = (C.c_int * len(self.devices))()
        n = C.c_int(0)
        self.capi.check(self.capi.lib.gps_sharded_get_host_io(self._grp, idx, C.byref(n)))
        return list(idx[: n.value])

    def set_host_io(self, indices: list[int]) -> None:
        arr = (C.c_int * len(indices))(*indices)
        self.capi.check(self.capi.lib.gps_sharded_set_host_io(self._grp, arr, len(indices)))

    def close(self) -> None:
        if self._grp.value:
            self.capi.lib.gps_sharded_destroy(self._grp)
            self._grp = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
