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
# end of round 2: ticket-free launches -- the default dispatch at these sizes (one-wave register tile with blockIdx order),
# its ticketed counterpart, and the Decoupled Fallback kernels with tile = blockIdx.x (register / smem tile), with stalls
for tf in (1, 0):
    check(lib.gps_set_option(ctx, b"ticket_free", tf))
    for kind in (capi.INCLUSIVE, capi.EXCLUSIVE):
        for n in (N, 1 << 20, 2049):
            check(lib.gps_scan(ctx, capi.ALGO_CSDL, kind, capi.U32, a, b, n, 3, None, None))
# the warp-level rescue of the one-wave kernels: one tile's CTA starts late, a successor re-reduces it (tiny / small / mid tile)
check(lib.gps_set_option(ctx, b"ticket_free", 1))
for n, late in ((40_000, 3), (N, 9), (1 << 20, 200)):
    check(lib.gps_set_option(ctx, b"debug_delay_tile", late + 1))
    check(lib.gps_scan(ctx, capi.ALGO_CSDL, capi.INCLUSIVE, capi.U32, a, b, n, 3, None, None))
    check(lib.gps_scan(ctx, capi.ALGO_CSDL, capi.EXCLUSIVE, capi.F32, a, b, n, 0, None, None))
check(lib.gps_set_option(ctx, b"debug_delay_tile", 0))
check(lib.gps_check(ctx, None))
check(lib.gps_set_option(ctx, b"tile_order", 1))
for variant in (1, 2):
    check(lib.gps_set_option(ctx, b"fallback_variant", variant))
    for kind in (capi.INCLUSIVE, capi.EXCLUSIVE):
        check(lib.gps_scan(ctx, capi.ALGO_CSDL_FALLBACK, kind, capi.U32, a, b, 1 << 20, 1, None, None))
        check(lib.gps_scan_emulated_deadlock(ctx, kind, capi.U32, a, b, N, 7, 2, stats, None))
        check(lib.gps_memcpy_d2d(f, a, N * 4, None))
        check(lib.gps_scan_emulated_deadlock(ctx, kind, capi.F32, f, f, N, 3, 1, stats, None))  # in place
check(lib.gps_set_option(ctx, b"fallback_variant", 0))
check(lib.gps_set_option(ctx, b"tile_order", 0))
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
# round 2: Decoupled Fallback for the wide scans (production variant, emulated stalls, in place)
for dtype, es in ((capi.U64, 8), (capi.F64, 8), (capi.U32X4, 16), (capi.U32, 4)):
    for kind in (capi.INCLUSIVE, capi.EXCLUSIVE):
        nel = (N * 4) // es - 1
        check(lib.gps_scan_wide(ctx, capi.ALGO_CSDL_FALLBACK, kind, dtype, a, b, nel, None, None, None))
        check(lib.gps_scan_wide_emulated_deadlock(ctx, kind, dtype, a, b, nel, None, 7, 2, stats, None))
        check(lib.gps_memcpy_d2d(f, a, N * 4, None))
        check(lib.gps_scan_wide_emulated_deadlock(ctx, kind, dtype, f, f, nel, None, 1, 1, stats, None))  # in place
# round 2: occupancy-estimation variant (persistent CTAs)
occ = C.c_uint32(0)
check(lib.gps_scan_estimate_occupancy(ctx, capi.INCLUSIVE, capi.U32, a, b, N, C.byref(occ), None))
check(lib.gps_scan_estimate_occupancy(ctx, capi.EXCLUSIVE, capi.F32, a, b, 1 << 20, C.byref(occ), None))
# round 2: L2 cache policies on the tile loads / stores, sharded-step options
for hints in (1, 2, 3):
    check(lib.gps_set_option(ctx, b"l2_hints", hints))
    for algo in (capi.ALGO_CSDL_SMEM, capi.ALGO_CSDL_BLOCKED):
        check(lib.gps_scan(ctx, algo, capi.INCLUSIVE, capi.U32, a, b, 1 << 20, 0, None, None))
check(lib.gps_set_option(ctx, b"l2_hints", -1))
slots = C.c_void_p()
check(lib.gps_malloc(C.byref(slots), 64 * 8))
check(lib.gps_memset(slots, 0, 64 * 8, None))
for l2 in (0, 7):
    check(lib.gps_set_option(ctx, b"sharded_l2", l2))
    for tag in (1, 2):
        dst = (C.c_void_p * 1)(slots.value + (tag & 1) * 16 * 8)
        check(lib.gps_reduce_publish(ctx, capi.U32, a, 1 << 20, dst, 1, tag, None, None))
        check(lib.gps_scan_carry_slots(ctx, capi.ALGO_CSDL_BLOCKED, capi.INCLUSIVE, capi.U32, a, b, 1 << 20,
                                       slots.value + (tag & 1) * 16 * 8, 1, 0, tag, None, None))
check(lib.gps_set_option(ctx, b"sharded_l2", 0))
check(lib.gps_init_random(a, 1 << 20, capi.U32, 9, 0xFFFFFFFF, 0, None))
# round 2: the sort with Decoupled Fallback in the digit look-back (no stalls / emulated stalls), partial digit
for n in (N, 10_240):
    check(lib.gps_sort_keys32_fallback(ctx, 0, a, b, c, None, None, None, n, 32, 0xFFFFFFFF, 4, None, None))
    check(lib.gps_sort_keys32_fallback(ctx, 0, a, b, c, a, d, e, n, 32, 7, 2, stats, None))
    check(lib.gps_sort_keys32_fallback(ctx, 1, a, b, c, None, None, None, n, 32, 1, 1, stats, None))
    check(lib.gps_sort_u32(ctx, a, b, c, None, None, None, n, 13, None))
for n in (N, 10_240, 3):
    check(lib.gps_sort_u32(ctx, a, b, c, None, None, None, n, 32, None))
    check(lib.gps_sort_u32(ctx, a, b, c, a, d, e, n, 32, None))
    check(lib.gps_sort_u32(ctx, a, b, c, None, None, None, n, 8, None))
    check(lib.gps_sort_keys32(ctx, 2, a, b, c, a, d, e, n, 32, None))  # float32 key maps, pairs
    check(lib.gps_sort_keys32(ctx, 1, a, b, c, None, None, None, n, 32, None))
# round 2: 64-bit keys (keys / pairs, key maps, partial last digit, Decoupled Fallback with emulated stalls)
sort64_sizes = (N // 2, 6144, 5121, 3)
for n in sort64_sizes:
    check(lib.gps_sort_keys64(ctx, 0, a, b, c, None, None, None, n, 64, None))
    check(lib.gps_sort_keys64(ctx, 0, a, b, c, a, d, e, n, 64, None))
    check(lib.gps_sort_keys64(ctx, 0, a, b, c, None, None, None, n, 45, None))
    check(lib.gps_sort_keys64(ctx, 2 | 0x100, a, b, c, a, d, e, n, 64, None))  # float64 maps, descending, pairs
    check(lib.gps_sort_keys64_fallback(ctx, 0, a, b, c, a, d, e, n, 64, 7, 2, stats, None))
    check(lib.gps_sort_keys64_fallback(ctx, 1, a, b, c, None, None, None, n, 64, 0xFFFFFFFF, 4, None, None))
if len(sys.argv) > 1 and sys.argv[1] == "sort64":  # only the 64-bit sort (and everything above it)
    check(lib.gps_check(ctx, None))
    print("san_driver done")
    sys.exit(0)
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
# round 2: the multi-GPU host entry with three "GPUs" on device 0, I/O routing, calibration
grp = C.c_void_p()
devs = (C.c_int * 3)(0, 0, 0)
check(lib.gps_sharded_create(C.byref(grp), 3, devs, 1 << 15))
for i in range(3):
    check(lib.gps_set_option(lib.gps_sharded_ctx(grp, i), b"host_chunk_keys", 1 << 13))
tot = C.c_uint32(0)
for k in range(3):
    check(lib.gps_sharded_scan_host_async(grp, capi.ALGO_CSDL, k & 1, capi.U32, hin, hout, N - k, k, None))
check(lib.gps_sharded_host_sync(grp))
io = (C.c_int * 2)(2, 0)
check(lib.gps_sharded_set_host_io(grp, io, 2))
check(lib.gps_sharded_scan_host(grp, capi.ALGO_RTS, capi.INCLUSIVE, capi.F32, hin, hout, N, 0, None))
rate = C.c_float(0)
check(lib.gps_sharded_calibrate_host(grp, hin, hout, N * 4, C.byref(rate)))
check(lib.gps_sharded_scan_host(grp, capi.ALGO_CSDL_FALLBACK, capi.INCLUSIVE, capi.U32, hin, hout, N, 0, None))
lib.gps_sharded_destroy(grp)
inv, sums = C.c_uint64(0), (C.c_uint64 * 2)()
check(lib.gps_sort_fingerprint(b, 3, C.byref(inv), sums))
lib.gps_destroy(ctx)
for x in bufs:
    lib.gps_free(x)
print("san_driver done")
