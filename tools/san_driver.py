"""compute-sanitizer target: every kernel family of libgps_b200.so at small ragged sizes through
the C ABI only (no torch, so the sanitizer sees just our launches).
  compute-sanitizer --tool {memcheck,racecheck,synccheck} python tools/san_driver.py"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gpuprefixsums_b200 import _capi as capi  # noqa: E402

lib, check = capi.lib, capi.check
N = 150_003
ctx = C.c_void_p()
check(lib.gps_create(C.byref(ctx), 0, 1 << 20))
bufs = [C.c_void_p() for _ in range(6)]
for b in bufs:
    check(lib.gps_malloc(C.byref(b), (1 << 20) * 4))
a, b, c, d, e, f = bufs
check(lib.gps_init_random(a, 1 << 20, capi.U32, 7, 0xFFFFFFFF, 0, None))
for algo in (capi.ALGO_CSDL_BLOCKED, capi.ALGO_CSDL_SMEM, capi.ALGO_CSDL_REG, capi.ALGO_CSDL_TMA, capi.ALGO_CSDL_FALLBACK,
             capi.ALGO_RTS):
    for kind in (capi.INCLUSIVE, capi.EXCLUSIVE):
        for n in (N, 1 << 20):  # 1 << 20 > 2^17: the smem kernel proper needs n > 2^21 via ALGO_CSDL only
            check(lib.gps_scan(ctx, algo, kind, capi.U32, a, b, n, 5, None, None))
check(lib.gps_scan(ctx, capi.ALGO_CSDL, capi.INCLUSIVE, capi.U32, a, b, 1000, 0, None, None))  # tiny variant
# forced kernel variants: Decoupled Fallback on the smem / blocked tiles (with emulated stalls,
# in place), RTS propagate on the blocked tile
stats = C.c_void_p()
check(lib.gps_malloc(C.byref(stats), 64))
check(lib.gps_memset(stats, 0, 64, None))
for variant in (1, 2, 3):
    check(lib.gps_set_option(ctx, b"fallback_variant", variant))
    for kind in (capi.INCLUSIVE, capi.EXCLUSIVE):
        check(lib.gps_scan(ctx, capi.ALGO_CSDL_FALLBACK, kind, capi.U32, a, b, N, 1, None, None))
        check(lib.gps_scan_emulated_deadlock(ctx, kind, capi.U32, a, b, N, 7, 2, stats, None))
        check(lib.gps_memcpy_d2d(f, a, N * 4, None))
        check(lib.gps_scan_emulated_deadlock(ctx, kind, capi.U32, f, f, N, 3, 1, stats, None))  # in place
check(lib.gps_set_option(ctx, b"fallback_variant", 0))
for variant in (1, 2):
    check(lib.gps_set_option(ctx, b"rts_variant", variant))
    for kind in (capi.INCLUSIVE, capi.EXCLUSIVE):
        check(lib.gps_scan(ctx, capi.ALGO_RTS, kind, capi.F32, a, b, N, 0, None, None))
check(lib.gps_set_option(ctx, b"rts_variant", 0))
check(lib.gps_scan(ctx, capi.ALGO_CSDL, capi.INCLUSIVE, capi.F32, a, b, N, 0, None, None))
for dtype, es in ((capi.U64, 8), (capi.F64, 8), (capi.U32X4, 16), (capi.U32, 4)):
    for algo in (capi.ALGO_CSDL, capi.ALGO_RTS):
        for kind in (capi.INCLUSIVE, capi.EXCLUSIVE):
            check(lib.gps_scan_wide(ctx, algo, kind, dtype, a, b, (N * 4) // es - 1, None, None, None))
            check(lib.gps_scan_wide(ctx, algo, kind, dtype, a, a, 37, None, None, None))  # in place, one partial tile
check(lib.gps_init_random(a, 1 << 20, capi.U32, 9, 0xFFFFFFFF, 0, None))
for n in (N, 10_240, 3):
    check(lib.gps_sort_u32(ctx, a, b, c, None, None, None, n, 32, None))
    check(lib.gps_sort_u32(ctx, a, b, c, a, d, e, n, 32, None))
    check(lib.gps_sort_u32(ctx, a, b, c, None, None, None, n, 8, None))
    check(lib.gps_sort_keys32(ctx, 2, a, b, c, a, d, e, n, 32, None))  # float32 key maps, pairs
    check(lib.gps_sort_keys32(ctx, 1, a, b, c, None, None, None, n, 32, None))
# host-buffer pipeline, queued calls
import ctypes
hin, hout = C.c_void_p(), C.c_void_p()
check(lib.gps_malloc_host(C.byref(hin), N * 4))
check(lib.gps_malloc_host(C.byref(hout), N * 4))
ctypes.memset(hin.value, 1, N * 4)
check(lib.gps_set_option(ctx, b"host_chunk_keys", 1 << 15))
for k in range(3):
    check(lib.gps_scan_host_async(ctx, capi.ALGO_CSDL, k & 1, capi.U32, hin, hout, N - k, k, None))
check(lib.gps_host_sync(ctx))
check(lib.gps_check(ctx, None))
inv, sums = C.c_uint64(0), (C.c_uint64 * 2)()
check(lib.gps_sort_fingerprint(b, 3, C.byref(inv), sums))
lib.gps_destroy(ctx)
for x in bufs:
    lib.gps_free(x)
print("san_driver done")
