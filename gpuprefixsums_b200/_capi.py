"""ctypes binding of the C ABI declared in ``include/gps_b200.h``.

This is plumbing for the tests and ``bench.py`` (PyTorch supplies device memory and
streams); the product is ``lib/libgps_b200.so``.  There is deliberately NO fallback: if
the shared library has not been built (``python -c 'import __graft_entry__ as g;
g.build()'`` or ``make -C gpuprefixsums_b200/csrc``) importing this module raises, and
every compute entry point fails with ``GPS_ERR_CUDA`` when no sm_100a GPU is present.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("GPS_B200_LIB") or os.path.join(_HERE, "lib", "libgps_b200.so")  # env: tuning builds
CUB_LIB_PATH = os.path.join(_HERE, "lib", "libgps_cub.so")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "gps_b200.h")

# enums (include/gps_b200.h)
GPS_OK = 0
GPS_ERR_INVALID, GPS_ERR_ALIGNMENT, GPS_ERR_CUDA = -1, -2, -3
GPS_ERR_WATCHDOG, GPS_ERR_MISMATCH, GPS_ERR_UNSUPPORTED = -4, -5, -6
ALGO_CSDL, ALGO_CSDL_REG, ALGO_CSDL_TMA, ALGO_CSDL_FALLBACK, ALGO_RTS, ALGO_CSDL_SMEM = 0, 1, 2, 3, 4, 5
ALGO_CSDL_BLOCKED = 6
INCLUSIVE, EXCLUSIVE = 0, 1
U32, F32, U64, F64, U32X4 = 0, 1, 2, 3, 4
NO_STALL = 0xFFFFFFFF
IPC_HANDLE_BYTES = 64

ALGO_NAMES = {
    "csdl": ALGO_CSDL,
    "csdl_reg": ALGO_CSDL_REG,
    "csdl_tma": ALGO_CSDL_TMA,
    "csdl_fallback": ALGO_CSDL_FALLBACK,
    "rts": ALGO_RTS,
    "csdl_smem": ALGO_CSDL_SMEM,
    "csdl_blocked": ALGO_CSDL_BLOCKED,
}

_vp, _u64, _u32, _i32 = C.c_void_p, C.c_uint64, C.c_uint32, C.c_int

# name -> (restype, argtypes); every symbol include/gps_b200.h declares
SIGNATURES = {
    "gps_abi_version": (_i32, []),
    "gps_last_error": (C.c_char_p, []),
    "gps_device_count": (_i32, [C.POINTER(_i32)]),
    "gps_device_info": (_i32, [_i32, C.c_char_p, C.POINTER(_i32), C.POINTER(_u64), C.POINTER(_u64)]),
    "gps_create": (_i32, [C.POINTER(_vp), _i32, _u64]),
    "gps_create_with_scratch": (_i32, [C.POINTER(_vp), _i32, _u64, _vp, _u64]),
    "gps_destroy": (None, [_vp]),
    "gps_check": (_i32, [_vp, _vp]),
    "gps_set_option": (_i32, [_vp, C.c_char_p, C.c_int64]),
    "gps_get_option": (_i32, [_vp, C.c_char_p, _vp]),
    "gps_scratch_bytes": (_u64, [_u64]),
    "gps_scan": (_i32, [_vp, _i32, _i32, _i32, _vp, _vp, _u64, _u32, _vp, _vp]),
    "gps_scan_wide": (_i32, [_vp, _i32, _i32, _i32, _vp, _vp, _u64, _vp, _vp, _vp]),
    "gps_scan_wide_emulated_deadlock": (_i32, [_vp, _i32, _i32, _vp, _vp, _u64, _vp, _u32, _u32, _vp, _vp]),
    "gps_sort_scratch_bytes": (_u64, [_u64]),
    "gps_sort_keys32": (_i32, [_vp, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _u64, _i32, _vp]),
    "gps_sort_keys32_fallback": (_i32, [_vp, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _u64, _i32, _u32, _u32, _vp, _vp]),
    "gps_sort_keys64": (_i32, [_vp, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _u64, _i32, _vp]),
    "gps_sort_keys64_fallback": (_i32, [_vp, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _u64, _i32, _u32, _u32, _vp, _vp]),
    "gps_sort_u32": (_i32, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _u64, _i32, _vp]),
    "gps_scan_carry_slots": (_i32, [_vp, _i32, _i32, _i32, _vp, _vp, _u64, _vp, _u32, _u32, _u32, _vp, _vp]),
    "gps_reduce": (_i32, [_vp, _i32, _vp, _u64, _vp, _vp]),
    "gps_reduce_publish": (_i32, [_vp, _i32, _vp, _u64, C.POINTER(_vp), _u32, _u32, _vp, _vp]),
    "gps_scan_emulated_deadlock": (_i32, [_vp, _i32, _i32, _vp, _vp, _u64, _u32, _u32, _vp, _vp]),
    "gps_scan_estimate_occupancy": (_i32, [_vp, _i32, _i32, _vp, _vp, _u64, C.POINTER(_u32), _vp]),
    "gps_scan_host": (_i32, [_vp, _i32, _i32, _i32, _vp, _vp, _u64, _u32, _vp]),
    "gps_scan_host_async": (_i32, [_vp, _i32, _i32, _i32, _vp, _vp, _u64, _u32, _vp]),
    "gps_host_sync": (_i32, [_vp]),
    "gps_init_ones": (_i32, [_vp, _u64, _i32, _vp]),
    "gps_init_random": (_i32, [_vp, _u64, _i32, _u64, _u32, _u64, _vp]),
    "gps_validate_ones": (_i32, [_vp, _u64, _i32, _u32, C.POINTER(_u64)]),
    "gps_validate_ones_wide": (_i32, [_vp, _u64, _i32, _i32, C.POINTER(_u64)]),
    "gps_sort_fingerprint": (_i32, [_vp, _u64, C.POINTER(_u64), C.POINTER(_u64)]),
    "gps_sort_fingerprint64": (_i32, [_vp, _u64, C.POINTER(_u64), C.POINTER(_u64)]),
    "gps_compare_u32": (_i32, [_vp, _vp, _u64, C.POINTER(_u64), C.POINTER(_u64)]),
    "gps_compare_f32": (_i32, [_vp, _vp, _vp, _u64, C.c_float, C.c_float, C.POINTER(_u64), C.POINTER(C.c_float)]),
    "gps_set_device": (_i32, [_i32]),
    "gps_malloc": (_i32, [C.POINTER(_vp), _u64]),
    "gps_free": (_i32, [_vp]),
    "gps_malloc_host": (_i32, [C.POINTER(_vp), _u64]),
    "gps_free_host": (_i32, [_vp]),
    "gps_memcpy_h2d": (_i32, [_vp, _vp, _u64, _vp]),
    "gps_memcpy_d2h": (_i32, [_vp, _vp, _u64, _vp]),
    "gps_memcpy_d2d": (_i32, [_vp, _vp, _u64, _vp]),
    "gps_memset": (_i32, [_vp, _i32, _u64, _vp]),
    "gps_stream_sync": (_i32, [_vp]),
    "gps_device_sync": (_i32, []),
    "gps_event_create": (_i32, [C.POINTER(_vp)]),
    "gps_event_destroy": (_i32, [_vp]),
    "gps_event_record": (_i32, [_vp, _vp]),
    "gps_event_elapsed_ms": (_i32, [_vp, _vp, C.POINTER(C.c_float)]),
    "gps_ipc_export": (_i32, [_vp, C.c_char_p]),
    "gps_ipc_open": (_i32, [C.c_char_p, C.POINTER(_vp)]),
    "gps_ipc_close": (_i32, [_vp]),
    "gps_sharded_create": (_i32, [C.POINTER(_vp), _i32, C.POINTER(_i32), _u64]),
    "gps_sharded_destroy": (None, [_vp]),
    "gps_sharded_scan": (_i32, [_vp, _i32, _i32, _i32, C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_u64), _i32,
                                C.POINTER(C.c_float)]),
    "gps_sharded_ctx": (_vp, [_vp, _i32]),
    "gps_sharded_scan_host": (_i32, [_vp, _i32, _i32, _i32, _vp, _vp, _u64, _u32, _vp]),
    "gps_sharded_scan_host_async": (_i32, [_vp, _i32, _i32, _i32, _vp, _vp, _u64, _u32, _vp]),
    "gps_sharded_host_sync": (_i32, [_vp]),
    "gps_sharded_calibrate_host": (_i32, [_vp, _vp, _vp, _u64, C.POINTER(C.c_float)]),
    "gps_sharded_set_host_io": (_i32, [_vp, C.POINTER(_i32), _i32]),
    "gps_sharded_get_host_io": (_i32, [_vp, C.POINTER(_i32), C.POINTER(_i32)]),
}


class GpsError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"gps error {code}: {message}")
        self.code = code


def _load(path: str) -> C.CDLL:
    if not os.path.exists(path):
        raise ImportError(
            f"{path} is missing: build the CUDA library first (make -C gpuprefixsums_b200/csrc, or "
            "__graft_entry__.build()). There is no CPU fallback."
        )
    return C.CDLL(path, mode=C.RTLD_GLOBAL)


lib = _load(LIB_PATH)
for _name, (_res, _args) in SIGNATURES.items():
    _fn = getattr(lib, _name)  # AttributeError here = header and library out of sync
    _fn.restype = _res
    _fn.argtypes = _args

_cub = None


def cub_lib() -> C.CDLL:
    """libgps_cub.so: cub::DeviceScan comparison baseline (library code, not ours)."""
    global _cub
    if _cub is None:
        _cub = _load(CUB_LIB_PATH)
        _cub.gps_cub_scan.restype = _i32
        _cub.gps_cub_scan.argtypes = [_i32, _i32, _vp, _vp, _u64, _vp]
        _cub.gps_cub_sort_u32.restype = _i32
        _cub.gps_cub_sort_u32.argtypes = [_vp, _vp, _vp, _vp, _u64, _i32, _vp]
        _cub.gps_cub_sort_u64.restype = _i32
        _cub.gps_cub_sort_u64.argtypes = [_vp, _vp, _vp, _vp, _u64, _i32, _vp]
    return _cub


def check(rc: int) -> None:
    if rc != GPS_OK:
        raise GpsError(rc, (lib.gps_last_error() or b"").decode())
