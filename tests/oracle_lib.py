"""Loader for the CPU oracle (oracle/scan_oracle.c) and the reference CUDA shim.

TEST INFRASTRUCTURE: only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline
legs import this.  The oracle is compiled on demand with gcc (seconds)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_SO = os.path.join(ORACLE_DIR, "_build", "liboracle_scan.so")
REF_SHIM_SO = os.path.join(ORACLE_DIR, "_ref", "libref_shim.so")

_u32p = np.ctypeslib.ndpointer(dtype=np.uint32, flags="C_CONTIGUOUS")
_f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")


def build(force: bool = False) -> str:
    src = os.path.join(ORACLE_DIR, "scan_oracle.c")
    if force or not os.path.exists(ORACLE_SO) or os.path.getmtime(ORACLE_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", ORACLE_DIR, "oracle"], stdout=subprocess.DEVNULL)
    return ORACLE_SO


class Oracle:
    def __init__(self, lib: C.CDLL):
        self.lib = lib
        lib.oracle_scan_u32.restype = C.c_uint32
        lib.oracle_scan_u32.argtypes = [_u32p, _u32p, C.c_uint64, C.c_int, C.c_uint32]
        lib.oracle_scan_u32_mt.restype = C.c_uint32
        lib.oracle_scan_u32_mt.argtypes = [_u32p, _u32p, C.c_uint64, C.c_int, C.c_uint32, C.c_int]
        lib.oracle_scan_f32_f64.restype = C.c_double
        lib.oracle_scan_f32_f64.argtypes = [_f32p, _f64p, C.c_void_p, C.c_uint64, C.c_int, C.c_double]
        lib.oracle_check_f32.restype = C.c_uint64
        lib.oracle_check_f32.argtypes = [_f32p, _f32p, C.c_uint64, C.c_int, C.c_double, C.c_double, C.c_double,
                                         C.POINTER(C.c_double)]
        lib.oracle_check_u32.restype = C.c_uint64
        lib.oracle_check_u32.argtypes = [_u32p, _u32p, C.c_uint64, C.c_int, C.c_uint32, C.POINTER(C.c_uint64)]
        _u64p = np.ctypeslib.ndpointer(dtype=np.uint64, flags="C_CONTIGUOUS")
        lib.oracle_scan_u64.restype = C.c_uint64
        lib.oracle_scan_u64.argtypes = [_u64p, _u64p, C.c_uint64, C.c_int, C.c_uint64]
        lib.oracle_scan_u32x4.restype = None
        lib.oracle_scan_u32x4.argtypes = [_u32p, _u32p, C.c_uint64, C.c_int, _u32p, _u32p]
        lib.oracle_check_f64.restype = C.c_uint64
        lib.oracle_check_f64.argtypes = [_f64p, _f64p, C.c_uint64, C.c_int, C.c_double, C.c_double, C.c_double,
                                         C.POINTER(C.c_double)]
        lib.oracle_sort_u32.restype = None
        lib.oracle_sort_u32.argtypes = [C.c_void_p] * 6 + [C.c_uint64, C.c_int]
        lib.oracle_sort_keys32.restype = None
        lib.oracle_sort_keys32.argtypes = [C.c_int] + [C.c_void_p] * 7 + [C.c_uint64]
        lib.oracle_validate_ones.restype = C.c_uint64
        lib.oracle_validate_ones.argtypes = [_u32p, C.c_uint64, C.c_int]
        lib.oracle_init_random_d3d12.restype = None
        lib.oracle_init_random_d3d12.argtypes = [_u32p, C.c_uint64, C.c_uint32]
        lib.oracle_init_random_u32.restype = None
        lib.oracle_init_random_u32.argtypes = [_u32p, C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint64]
        lib.oracle_init_random_f32.restype = None
        lib.oracle_init_random_f32.argtypes = [_f32p, C.c_uint64, C.c_uint64, C.c_uint64]

    # --- scans -------------------------------------------------------------------
    def scan_u32(self, x: np.ndarray, exclusive: bool = False, carry: int = 0):
        x = np.ascontiguousarray(x, dtype=np.uint32)
        out = np.empty_like(x)
        total = self.lib.oracle_scan_u32(x, out, x.size, int(exclusive), carry & 0xFFFFFFFF)
        return out, total

    def scan_u32_mt(self, x: np.ndarray, out: np.ndarray, exclusive: bool, carry: int, threads: int) -> int:
        return self.lib.oracle_scan_u32_mt(x, out, x.size, int(exclusive), carry & 0xFFFFFFFF, threads)

    def scan_f32(self, x: np.ndarray, exclusive: bool = False, carry: float = 0.0):
        x = np.ascontiguousarray(x, dtype=np.float32)
        out = np.empty(x.size, dtype=np.float64)
        total = self.lib.oracle_scan_f32_f64(x, out, None, x.size, int(exclusive), float(carry))
        return out, total

    def check_f32(self, x: np.ndarray, cand: np.ndarray, exclusive: bool = False, carry: float = 0.0,
                  rtol: float = 2e-5, atol: float = 1e-6):
        """(errors, max relative error) of a float32 candidate vs the float64 serial scan."""
        mx = C.c_double(0.0)
        errs = self.lib.oracle_check_f32(np.ascontiguousarray(x, dtype=np.float32),
                                         np.ascontiguousarray(cand, dtype=np.float32), x.size, int(exclusive),
                                         float(carry), rtol, atol, C.byref(mx))
        return errs, mx.value

    def check_u32(self, x: np.ndarray, cand: np.ndarray, exclusive: bool = False, carry: int = 0):
        first = C.c_uint64(0)
        errs = self.lib.oracle_check_u32(np.ascontiguousarray(x, dtype=np.uint32),
                                         np.ascontiguousarray(cand, dtype=np.uint32), x.size, int(exclusive),
                                         carry & 0xFFFFFFFF, C.byref(first))
        return errs, first.value

    def scan_u64(self, x: np.ndarray, exclusive: bool = False, carry: int = 0):
        x = np.ascontiguousarray(x, dtype=np.uint64)
        out = np.empty_like(x)
        total = self.lib.oracle_scan_u64(x, out, x.size, int(exclusive), carry & 0xFFFFFFFFFFFFFFFF)
        return out, total

    def scan_u32x4(self, x: np.ndarray, exclusive: bool = False, carry=(0, 0, 0, 0)):
        """x: flat uint32 view of n vec4 elements (4n values)."""
        x = np.ascontiguousarray(x, dtype=np.uint32)
        out = np.empty_like(x)
        total = np.zeros(4, dtype=np.uint32)
        self.lib.oracle_scan_u32x4(x, out, x.size // 4, int(exclusive), np.asarray(carry, dtype=np.uint32), total)
        return out, total

    def check_f64(self, x: np.ndarray, cand: np.ndarray, exclusive: bool = False, carry: float = 0.0,
                  rtol: float = 1e-13, atol: float = 1e-300):
        mx = C.c_double(0.0)
        errs = self.lib.oracle_check_f64(np.ascontiguousarray(x, dtype=np.float64),
                                         np.ascontiguousarray(cand, dtype=np.float64), x.size, int(exclusive),
                                         float(carry), rtol, atol, C.byref(mx))
        return errs, mx.value

    def sort_u32(self, keys: np.ndarray, vals: np.ndarray | None = None, end_bit: int = 32):
        """Stable LSD radix sort on the low end_bit bits; returns keys or (keys, vals)."""
        keys = np.ascontiguousarray(keys, dtype=np.uint32)
        ko, kt = np.empty_like(keys), np.empty_like(keys)
        ptr = lambda a: a.ctypes.data_as(C.c_void_p) if a is not None else None
        if vals is None:
            self.lib.oracle_sort_u32(ptr(keys), ptr(ko), ptr(kt), None, None, None, keys.size, end_bit)
            return ko
        vals = np.ascontiguousarray(vals, dtype=np.uint32)
        vo, vt = np.empty_like(vals), np.empty_like(vals)
        self.lib.oracle_sort_u32(ptr(keys), ptr(ko), ptr(kt), ptr(vals), ptr(vo), ptr(vt), keys.size, end_bit)
        return ko, vo

    def sort_keys32(self, keys: np.ndarray, vals: np.ndarray | None = None, key_type: str = "u32",
                    descending: bool = False):
        """keys: int32 / float32 / uint32 array; sorted by numeric order of that type (stable)."""
        kt = {"u32": 0, "i32": 1, "f32": 2}[key_type] | (0x100 if descending else 0)
        bits = np.ascontiguousarray(keys).view(np.uint32)
        ko, ktmp, kmap = np.empty_like(bits), np.empty_like(bits), np.empty_like(bits)
        ptr = lambda a: a.ctypes.data_as(C.c_void_p) if a is not None else None
        vo = vt = None
        if vals is not None:
            vals = np.ascontiguousarray(vals, dtype=np.uint32)
            vo, vt = np.empty_like(vals), np.empty_like(vals)
        self.lib.oracle_sort_keys32(kt, ptr(bits), ptr(ko), ptr(ktmp), ptr(kmap), ptr(vals), ptr(vo), ptr(vt), bits.size)
        ko = ko.view(keys.dtype)
        return ko if vals is None else (ko, vo)

    def sort_keys64(self, keys: np.ndarray, vals: np.ndarray | None = None, key_type: str = "u64",
                    descending: bool = False, end_bit: int = 64):
        """keys: uint64 / int64 / float64 array; stable sort by that type's order (on the low end_bit bits
        for uint64)."""
        kt = {"u64": 0, "i64": 1, "f64": 2}[key_type] | (0x100 if descending else 0)
        bits = np.ascontiguousarray(keys).view(np.uint64)
        ko, ktmp = np.empty_like(bits), np.empty_like(bits)
        ptr = lambda a: a.ctypes.data_as(C.c_void_p) if a is not None else None
        vo = vt = None
        if vals is not None:
            vals = np.ascontiguousarray(vals, dtype=np.uint32)
            vo, vt = np.empty_like(vals), np.empty_like(vals)
        self.lib.oracle_sort_keys64.restype = None
        self.lib.oracle_sort_keys64.argtypes = [C.c_int] + [C.c_void_p] * 6 + [C.c_uint64, C.c_int]
        self.lib.oracle_sort_keys64(kt, ptr(bits), ptr(ko), ptr(ktmp), ptr(vals), ptr(vo), ptr(vt), bits.size, end_bit)
        ko = ko.view(keys.dtype)
        return ko if vals is None else (ko, vo)

    def validate_ones(self, scan: np.ndarray, exclusive: bool) -> int:
        return self.lib.oracle_validate_ones(np.ascontiguousarray(scan, dtype=np.uint32), scan.size, int(exclusive))

    # --- inputs ------------------------------------------------------------------
    def random_u32(self, n: int, seed: int, mask: int = 0xFFFFFFFF, first_index: int = 0) -> np.ndarray:
        out = np.empty(n, dtype=np.uint32)
        self.lib.oracle_init_random_u32(out, n, seed, mask, first_index)
        return out

    def random_f32(self, n: int, seed: int, first_index: int = 0) -> np.ndarray:
        out = np.empty(n, dtype=np.float32)
        self.lib.oracle_init_random_f32(out, n, seed, first_index)
        return out

    def random_d3d12(self, n: int, seed: int) -> np.ndarray:
        out = np.empty(n, dtype=np.uint32)
        self.lib.oracle_init_random_d3d12(out, n, seed)
        return out


def load() -> Oracle:
    return Oracle(C.CDLL(build()))


class RefShim:
    """oracle/_ref/libref_shim.so: the reference's unmodified CUDA kernels on caller buffers."""

    CSDL, RTS, EMULATED_DEADLOCK, CUB = 0, 1, 2, 3

    def __init__(self, lib: C.CDLL):
        self.lib = lib
        lib.ref_scan_u32.restype = C.c_int
        lib.ref_scan_u32.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_uint32]
        lib.ref_cub_scan_f32.restype = C.c_int
        lib.ref_cub_scan_f32.argtypes = [C.c_int, C.c_void_p, C.c_uint32]
        lib.ref_init_one.restype = None
        lib.ref_init_one.argtypes = [C.c_void_p, C.c_uint32]
        lib.ref_validate.restype = C.c_uint32
        lib.ref_validate.argtypes = [C.c_int, C.c_void_p, C.c_uint32]
        lib.ref_time_u32.restype = C.c_double
        lib.ref_time_u32.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32]

        if hasattr(lib, "ref_bench_all"):
            lib.ref_bench_all.restype = C.c_int
            lib.ref_bench_all.argtypes = [C.c_uint32, C.c_uint32, C.POINTER(C.c_double)]

    def scan_u32(self, algo: int, exclusive: bool, d_ptr: int, size: int) -> int:
        return self.lib.ref_scan_u32(algo, int(exclusive), d_ptr, size)

    def bench_all(self, size: int, batch: int):
        """(csdl_ms, rts_ms, cub_ms) of the reference's kernels on the shim's own buffers."""
        out = (C.c_double * 3)()
        rc = self.lib.ref_bench_all(size, batch, out)
        return None if rc else tuple(out)

    def time_u32(self, algo: int, exclusive: bool, d_ptr: int, d_src: int | None, size: int, batch: int) -> float:
        return self.lib.ref_time_u32(algo, int(exclusive), d_ptr, d_src, size, batch)


def load_ref_shim() -> RefShim | None:
    if not os.path.exists(REF_SHIM_SO):
        return None
    return RefShim(C.CDLL(REF_SHIM_SO))
