"""gpuprefixsums_b200 -- B200-native (sm_100a) device-wide prefix sums.

The product is the C-ABI shared library ``lib/libgps_b200.so`` (CUDA kernels in
``csrc/``, interface in ``include/gps_b200.h``) plus the C++ dispatcher classes in
``host/`` that keep the reference's operator surface.  This Python package is the thin
ctypes plumbing used by the tests and ``bench.py``.  Importing it requires the built
library; there is no CPU fallback anywhere.
"""
from . import _capi  # noqa: F401  (raises ImportError if the library is not built)
from ._capi import (ALGO_CSDL, ALGO_CSDL_BLOCKED, ALGO_CSDL_FALLBACK, ALGO_CSDL_REG, ALGO_CSDL_SMEM, ALGO_CSDL_TMA, ALGO_RTS, EXCLUSIVE, F32,
                    INCLUSIVE, NO_STALL, U32, GpsError)
from .scan import PrefixSum, compare_u32, cub_scan, cub_sort, init_ones, init_random, validate_ones  # noqa: F401
from .sharded import ShardedHostScan, ShardedScan, slice_bounds  # noqa: F401

__all__ = [
    "PrefixSum", "ShardedScan", "ShardedHostScan", "slice_bounds", "GpsError", "init_ones", "init_random", "validate_ones",
    "compare_u32", "cub_scan", "cub_sort", "ALGO_CSDL", "ALGO_CSDL_REG", "ALGO_CSDL_TMA", "ALGO_CSDL_SMEM", "ALGO_CSDL_BLOCKED", "ALGO_CSDL_FALLBACK", "ALGO_RTS",
    "INCLUSIVE", "EXCLUSIVE", "U32", "F32", "NO_STALL",
]
