"""Host-side Python surface over the C ABI: a context object working on torch tensors.

PyTorch is plumbing here (device memory, streams); every scan runs in
``lib/libgps_b200.so``.  The C++ dispatcher classes with the reference's names live in
``host/*.hpp``; ``dispatchers.py`` mirrors them for the Python tests.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _capi as capi
from ._capi import check, lib

_DTYPES = {torch.int32: capi.U32, torch.uint32: capi.U32, torch.float32: capi.F32}


def _dtype_code(t: torch.Tensor) -> int:
    try:
        return _DTYPES[t.dtype]
    except KeyError:
        raise TypeError(f"unsupported dtype {t.dtype}: uint32/int32 (bit pattern) or float32") from None


def _stream_ptr(stream: Optional[torch.cuda.Stream], device: torch.device) -> int:
    s = stream if stream is not None else torch.cuda.current_stream(device)
    return s.cuda_stream


def _algo(algo) -> int:
    return capi.ALGO_NAMES[algo] if isinstance(algo, str) else int(algo)


def _kind(exclusive: bool) -> int:
    return capi.EXCLUSIVE if exclusive else capi.INCLUSIVE


def carry_bits(value, dtype: torch.dtype) -> int:
    """uint32 value, or the bit pattern of a float32."""
    if dtype == torch.float32:
        return int(torch.tensor([float(value)], dtype=torch.float32).view(torch.int32).item()) & 0xFFFFFFFF
    return int(value) & 0xFFFFFFFF


class PrefixSum:
    """One ``gps_ctx``: scratch for scans of up to ``max_n`` keys on ``device``.

    Use one instance per stream; calls are asynchronous on the current torch stream.
    """

    def __init__(self, max_n: int, device: int | torch.device | None = None,
                 scratch: Optional[torch.Tensor] = None):
        if not torch.cuda.is_available():
            raise RuntimeError("gpuprefixsums_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        if dev.type != "cuda":
            raise ValueError("device must be a CUDA device")
        self.device = torch.device("cuda", dev.index if dev.index is not None else torch.cuda.current_device())
        self.max_n = int(max_n)
        self._ctx = C.c_void_p()
        # scratch: None = the library allocates; a uint8 CUDA tensor of at least
        # scratch_bytes(max_n) bytes = caller-owned temp storage (Unity-style API)
        self._scratch = scratch
        if scratch is None:
            check(lib.gps_create(C.byref(self._ctx), self.device.index, self.max_n))
        else:
            check(lib.gps_create_with_scratch(C.byref(self._ctx), self.device.index, self.max_n,
                                              scratch.data_ptr(), scratch.numel() * scratch.element_size()))

    @staticmethod
    def scratch_bytes(max_n: int) -> int:
        return int(lib.gps_scratch_bytes(int(max_n)))

    def close(self) -> None:
        if getattr(self, "_ctx", None) is not None and self._ctx.value:
            lib.gps_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- options -------------------------------------------------------------------
    def set_option(self, key: str, value: int) -> None:
        check(lib.gps_set_option(self._ctx, key.encode(), int(value)))

    def get_option(self, key: str) -> int:
        """Read an option back ("ticket_free", "pdl", "tile_order"; "rescued": 1 if a check() since the last query saw a
        rescued tile -- reading clears it)."""
        v = C.c_int64(0)
        check(lib.gps_get_option(self._ctx, key.encode(), C.byref(v)))
        return int(v.value)

    def check(self, stream: Optional[torch.cuda.Stream] = None) -> None:
        """Synchronise and raise if a look-back watchdog tripped since the last check."""
        check(lib.gps_check(self._ctx, _stream_ptr(stream, self.device)))

    # ---- scans ---------------------------------------------------------------------
    def _prep(self, inp: torch.Tensor, out: Optional[torch.Tensor]):
        if not inp.is_cuda or inp.device != self.device:
            raise ValueError(f"input must live on {self.device}")
        if not inp.is_contiguous():
            raise ValueError("input must be contiguous")
        if out is None:
            out = torch.empty_like(inp)
        if out.dtype != inp.dtype or out.numel() != inp.numel() or not out.is_contiguous() or out.device != inp.device:
            raise ValueError("output must match the input's dtype, size and device and be contiguous")
        return inp, out

    def scan(self, inp: torch.Tensor, out: Optional[torch.Tensor] = None, *, algo="csdl",
             exclusive: bool = False, carry=0, total_out: Optional[torch.Tensor] = None,
             stream: Optional[torch.cuda.Stream] = None) -> torch.Tensor:
        inp, out = self._prep(inp, out)
        check(lib.gps_scan(self._ctx, _algo(algo), _kind(exclusive), _dtype_code(inp), inp.data_ptr(),
                           out.data_ptr(), inp.numel(), carry_bits(carry, inp.dtype),
                           total_out.data_ptr() if total_out is not None else None,
                           _stream_ptr(stream, self.device)))
        return out

    def _wide_args(self, inp: torch.Tensor, struct4: bool, carry):
        if struct4:
            if inp.dtype not in (torch.int32, torch.uint32) or inp.numel() % 4:
                raise TypeError("struct4 needs an int32 tensor with a multiple of 4 values")
            dtype, n, host = capi.U32X4, inp.numel() // 4, (C.c_uint32 * 4)
            cval = host(*[int(v) & 0xFFFFFFFF for v in carry]) if carry is not None else None
        elif inp.dtype == torch.int64:
            dtype, n = capi.U64, inp.numel()
            cval = C.c_uint64(int(carry) & 0xFFFFFFFFFFFFFFFF) if carry is not None else None
        elif inp.dtype == torch.float64:
            dtype, n = capi.F64, inp.numel()
            cval = C.c_double(float(carry)) if carry is not None else None
        elif inp.dtype in (torch.int32, torch.uint32):  # blocked-tile kernel at 32 bits (comparison)
            dtype, n = capi.U32, inp.numel()
            cval = C.c_uint32(int(carry) & 0xFFFFFFFF) if carry is not None else None
        elif inp.dtype == torch.float32:
            dtype, n = capi.F32, inp.numel()
            cval = C.c_float(float(carry)) if carry is not None else None
        else:
            raise TypeError(f"scan_wide: unsupported dtype {inp.dtype}")
        return dtype, n, cval

    def scan_wide(self, inp: torch.Tensor, out: Optional[torch.Tensor] = None, *, algo="csdl",
                  exclusive: bool = False, struct4: bool = False, carry=None, total_out: Optional[torch.Tensor] = None,
                  stream: Optional[torch.cuda.Stream] = None) -> torch.Tensor:
        """uint64 (torch.int64 bit pattern), float64, or -- with struct4=True on an int32 tensor
        whose last dimension is 4 -- the four-lane struct scan.  algo: "csdl" (chained scan,
        default), "csdl_fallback" (chained scan with Decoupled Fallback) or "rts" (Reduce-then-Scan,
        deterministic for float64)."""
        if algo not in ("csdl", "rts", "csdl_fallback"):
            raise ValueError("scan_wide: algo must be 'csdl', 'csdl_fallback' or 'rts'")
        inp, out = self._prep(inp, out)
        dtype, n, cval = self._wide_args(inp, struct4, carry)
        check(lib.gps_scan_wide(self._ctx, capi.ALGO_NAMES[algo], _kind(exclusive), dtype, inp.data_ptr(), out.data_ptr(), n,
                                C.cast(C.byref(cval), C.c_void_p) if cval is not None else None,
                                total_out.data_ptr() if total_out is not None else None,
                                _stream_ptr(stream, self.device)))
        return out

    def scan_wide_emulated_deadlock(self, inp: torch.Tensor, out: Optional[torch.Tensor] = None, *, exclusive: bool = False,
                                    struct4: bool = False, carry=None, stall_mask: int = 7, max_spin: int = 4,
                                    stats: Optional[torch.Tensor] = None,
                                    stream: Optional[torch.cuda.Stream] = None) -> torch.Tensor:
        """The wide scans' Decoupled Fallback kernel with emulated stalls (see gps_b200.h)."""
        inp, out = self._prep(inp, out)
        dtype, n, cval = self._wide_args(inp, struct4, carry)
        check(lib.gps_scan_wide_emulated_deadlock(self._ctx, _kind(exclusive), dtype, inp.data_ptr(), out.data_ptr(), n,
                                                  C.cast(C.byref(cval), C.c_void_p) if cval is not None else None,
                                                  stall_mask & 0xFFFFFFFF, max_spin,
                                                  stats.data_ptr() if stats is not None else None,
                                                  _stream_ptr(stream, self.device)))
        return out

    def sort(self, keys: torch.Tensor, values: Optional[torch.Tensor] = None, *, end_bit: Optional[int] = None,
             key_type: str = "u32",
             descending: bool = False,
             keys_out: Optional[torch.Tensor] = None, values_out: Optional[torch.Tensor] = None,
             stream: Optional[torch.cuda.Stream] = None, fallback: bool = False, stall_mask: int = 0xFFFFFFFF,
             max_spin: int = 4, stats: Optional[torch.Tensor] = None):
        """Stable ascending LSD radix sort of 32- or 64-bit keys, optionally moving a 32-bit payload;
        returns keys or (keys, values).  key_type: "u32" (default: the bit pattern as unsigned,
        on the low `end_bit` bits), "i32" (signed order), "f32" (numeric order, float32 keys), or
        "u64" / "i64" / "f64" for 64-bit keys (int64 / float64 tensors).  end_bit defaults to the key width."""
        wide = key_type in ("u64", "i64", "f64")  # 64-bit keys: same call, 8 digit passes
        kt = {"u32": 0, "i32": 1, "f32": 2, "u64": 0, "i64": 1, "f64": 2}[key_type]
        if wide:
            want = (torch.float64,) if kt == 2 else (torch.int64,)
            if end_bit is None:
                end_bit = 64
        else:
            want = (torch.float32,) if kt == 2 else (torch.int32, torch.uint32)
            if end_bit is None:
                end_bit = 32
        if descending:
            kt |= 0x100
        if keys.dtype not in want or not keys.is_cuda or not keys.is_contiguous():
            raise TypeError(f"sort: keys must be a contiguous CUDA tensor of dtype {want}")
        ko = torch.empty_like(keys) if keys_out is None else keys_out
        ktmp = torch.empty_like(keys)
        if values is not None:
            if values.numel() != keys.numel() or values.element_size() != 4 or not values.is_contiguous():
                raise TypeError("sort: values must be 32-bit, contiguous, one per key")
            vo = torch.empty_like(values) if values_out is None else values_out
            vt = torch.empty_like(values)
            vptrs = (values.data_ptr(), vo.data_ptr(), vt.data_ptr())
        else:
            vo, vptrs = None, (None, None, None)
        fn_plain, fn_fb = (lib.gps_sort_keys64, lib.gps_sort_keys64_fallback) if wide else (lib.gps_sort_keys32,
                                                                                          lib.gps_sort_keys32_fallback)
        if fallback:  # Decoupled Fallback in the digit look-back (+ emulated stalls for the tests)
            check(fn_fb(self._ctx, kt, keys.data_ptr(), ko.data_ptr(), ktmp.data_ptr(), *vptrs,
                                               keys.numel(), int(end_bit), stall_mask & 0xFFFFFFFF, max_spin,
                                               stats.data_ptr() if stats is not None else None,
                                               _stream_ptr(stream, self.device)))
        else:
            check(fn_plain(self._ctx, kt, keys.data_ptr(), ko.data_ptr(), ktmp.data_ptr(), *vptrs, keys.numel(),
                           int(end_bit), _stream_ptr(stream, self.device)))
        return ko if values is None else (ko, vo)

    def scan_carry_slots(self, inp: torch.Tensor, out: Optional[torch.Tensor], slots_ptr: int, n_slots: int,
                         n_carry: int, tag: int, *, algo="csdl", exclusive: bool = False,
                         total_out: Optional[torch.Tensor] = None,
                         stream: Optional[torch.cuda.Stream] = None) -> torch.Tensor:
        inp, out = self._prep(inp, out)
        check(lib.gps_scan_carry_slots(self._ctx, _algo(algo), _kind(exclusive), _dtype_code(inp),
                                       inp.data_ptr(), out.data_ptr(), inp.numel(), slots_ptr, n_slots,
                                       n_carry, tag & 0xFFFFFFFF,
                                       total_out.data_ptr() if total_out is not None else None,
                                       _stream_ptr(stream, self.device)))
        return out

    def reduce(self, inp: torch.Tensor, total_out: Optional[torch.Tensor] = None,
               stream: Optional[torch.cuda.Stream] = None) -> torch.Tensor:
        if total_out is None:
            total_out = torch.empty(1, dtype=inp.dtype, device=inp.device)
        check(lib.gps_reduce(self._ctx, _dtype_code(inp), inp.data_ptr(), inp.numel(), total_out.data_ptr(),
                             _stream_ptr(stream, self.device)))
        return total_out

    def reduce_publish(self, inp: torch.Tensor, dst_slot_ptrs: list[int], tag: int,
                       total_out: Optional[torch.Tensor] = None,
                       stream: Optional[torch.cuda.Stream] = None) -> None:
        arr = (C.c_void_p * max(1, len(dst_slot_ptrs)))(*dst_slot_ptrs)
        check(lib.gps_reduce_publish(self._ctx, _dtype_code(inp), inp.data_ptr(), inp.numel(), arr,
                                     len(dst_slot_ptrs), tag & 0xFFFFFFFF,
                                     total_out.data_ptr() if total_out is not None else None,
                                     _stream_ptr(stream, self.device)))

    def scan_emulated_deadlock(self, inp: torch.Tensor, out: Optional[torch.Tensor] = None, *,
                               exclusive: bool = False, stall_mask: int = 7, max_spin: int = 4,
                               stats: Optional[torch.Tensor] = None,
                               stream: Optional[torch.cuda.Stream] = None) -> torch.Tensor:
        inp, out = self._prep(inp, out)
        check(lib.gps_scan_emulated_deadlock(self._ctx, _kind(exclusive), _dtype_code(inp), inp.data_ptr(),
                                             out.data_ptr(), inp.numel(), stall_mask & 0xFFFFFFFF, max_spin,
                                             stats.data_ptr() if stats is not None else None,
                                             _stream_ptr(stream, self.device)))
        return out

    def estimate_occupancy(self, inp: torch.Tensor, out: Optional[torch.Tensor] = None, *, exclusive: bool = False):
        """(scan result, estimated number of concurrently resident CTAs) -- the reference's csdldf_occ variant."""
        inp, out = self._prep(inp, out)
        occ = C.c_uint32(0)
        check(lib.gps_scan_estimate_occupancy(self._ctx, _kind(exclusive), _dtype_code(inp), inp.data_ptr(), out.data_ptr(),
                                              inp.numel(), C.byref(occ), _stream_ptr(None, self.device)))
        return out, int(occ.value)

    def scan_host(self, h_in: torch.Tensor, h_out: Optional[torch.Tensor] = None, *, algo="csdl",
                  exclusive: bool = False, carry=0, wait: bool = True) -> torch.Tensor:
        """Host buffers in, host buffers out (pinned recommended).  wait=False returns once the
        work is queued (consecutive calls then overlap their copies); host_sync() waits."""
        if h_in.is_cuda or not h_in.is_contiguous():
            raise ValueError("scan_host takes contiguous CPU tensors")
        if h_out is None:
            h_out = torch.empty_like(h_in, pin_memory=h_in.is_pinned())
        fn = lib.gps_scan_host if wait else lib.gps_scan_host_async
        check(fn(self._ctx, _algo(algo), _kind(exclusive), _dtype_code(h_in), h_in.data_ptr(),
                 h_out.data_ptr(), h_in.numel(), carry_bits(carry, h_in.dtype), None))
        return h_out

    def host_sync(self) -> None:
        check(lib.gps_host_sync(self._ctx))


# ---- harness helpers (UtilityKernels.cuh:15-38 equivalents) --------------------------
def init_ones(t: torch.Tensor) -> None:
    check(lib.gps_init_ones(t.data_ptr(), t.numel(), _dtype_code(t), _stream_ptr(None, t.device)))


def init_random(t: torch.Tensor, seed: int, mask: int = 0xFFFFFFFF, first_index: int = 0) -> None:
    check(lib.gps_init_random(t.data_ptr(), t.numel(), _dtype_code(t), seed, mask, first_index,
                              _stream_ptr(None, t.device)))


def validate_ones(t: torch.Tensor, exclusive: bool, first_value: int = 0) -> int:
    torch.cuda.current_stream(t.device).synchronize()
    errs = C.c_uint64(0)
    check(lib.gps_validate_ones(t.data_ptr(), t.numel(), _kind(exclusive), first_value, C.byref(errs)))
    return errs.value


def compare_u32(a: torch.Tensor, b: torch.Tensor) -> tuple[int, int]:
    torch.cuda.current_stream(a.device).synchronize()
    errs, first = C.c_uint64(0), C.c_uint64(0)
    check(lib.gps_compare_u32(a.data_ptr(), b.data_ptr(), a.numel(), C.byref(errs), C.byref(first)))
    return errs.value, first.value


def cub_scan(inp: torch.Tensor, out: torch.Tensor, exclusive: bool = False) -> torch.Tensor:
    """cub::DeviceScan baseline (libgps_cub.so) on the current stream."""
    rc = capi.cub_lib().gps_cub_scan(int(exclusive), _dtype_code(inp), inp.data_ptr(), out.data_ptr(),
                                     inp.numel(), _stream_ptr(None, inp.device))
    if rc != 0:
        raise RuntimeError("cub::DeviceScan failed")
    return out


def cub_sort(keys: torch.Tensor, keys_out: torch.Tensor, values: Optional[torch.Tensor] = None,
             values_out: Optional[torch.Tensor] = None, end_bit: Optional[int] = None) -> None:
    """cub::DeviceRadixSort baseline (libgps_cub.so) on the current stream; int64 tensors = uint64 keys."""
    wide = keys.dtype == torch.int64
    if end_bit is None:
        end_bit = 64 if wide else 32
    fn = capi.cub_lib().gps_cub_sort_u64 if wide else capi.cub_lib().gps_cub_sort_u32
    rc = fn(keys.data_ptr(), keys_out.data_ptr(),
                                         values.data_ptr() if values is not None else None,
                                         values_out.data_ptr() if values_out is not None else None,
                                         keys.numel(), end_bit, _stream_ptr(None, keys.device))
    if rc != 0:
        raise RuntimeError("cub::DeviceRadixSort failed")
