// gps_config.h -- compile-time kernel configurations baked into libgps_b200.so.
// Chosen by measurement on B200 (tools/gps_tune sweeps the alternatives; the numbers
// behind each choice are in DESIGN.md and profiles/).
#pragma once

// CSDL register path: NT threads, V uint4 per thread -> tile = NT * V * 4 keys
// Measured at 2^28 uint32 keys on B200 (profiles/r01_tune_sweep.txt): <128,12> 785-803,
// <128,10> 796, <64,24> 782-812, <128,16> 782-790, <256,8> 752, <256,4> (the reference's
// shape) 662 Gkeys/s.  24 KB tiles, 7 CTAs/SM -> 168 KB of loads in flight per SM.
#ifndef GPS_REG_NT
#define GPS_REG_NT 128
#endif
#ifndef GPS_REG_V
#define GPS_REG_V 12
#endif
// below this many keys the scan is launch/latency bound and smaller tiles win
// (2^20 keys: <256,4> 6.4 us vs <128,12> 8.2 us; equal at 2^21..2^23; <128,12> ahead from 2^24)
// (round 2, with programmatic dependent launch: 2^22 keys <256,4> 8.7 us, smem<128,8> 10.7 us; from
// 2^23 the smem-resident tile is ahead, 12.9 vs 16.2 us)
#ifndef GPS_REG_SMALL_MAX_N
#define GPS_REG_SMALL_MAX_N (1u << 22)
#endif
#ifndef GPS_REG_SMALL_NT
#define GPS_REG_SMALL_NT 256
#endif
#ifndef GPS_REG_SMALL_V
#define GPS_REG_SMALL_V 4
#endif
// launch-latency regime (profiles/r01_size_sweep.txt): up to 2^17 keys 8 KB tiles <128,4> finish
// in 4.6 us where <256,4> takes 5.1 (14.2 vs 12.8 Gkeys/s at 2^16); from 2^18 <256,4> is ahead
#ifndef GPS_REG_TINY_MAX_N
#define GPS_REG_TINY_MAX_N (1u << 17)
#endif
#ifndef GPS_REG_TINY_NT
#define GPS_REG_TINY_NT 128
#endif
#ifndef GPS_REG_TINY_V
#define GPS_REG_TINY_V 4
#endif
// look-back windows (32 descriptors each) fetched per L2 round trip: 1 beat 2 and 4
// (803 / 757 / 530 Gkeys/s on <128,12> resp. <256,4>): the nearest tiles decide
#ifndef GPS_LOOKBACK_WIN
#define GPS_LOOKBACK_WIN 1
#endif

// CSDL TMA path: NT threads x 16 keys per tile, STAGES shared-memory tile buffers
#ifndef GPS_TMA_NT
#define GPS_TMA_NT 512
#endif
#ifndef GPS_TMA_STAGES
#define GPS_TMA_STAGES 2
#endif
// 0 = as many CTAs per SM as fit
#ifndef GPS_TMA_MAX_CTAS_PER_SM
#define GPS_TMA_MAX_CTAS_PER_SM 0
#endif

// Reduce-then-Scan tiles
#ifndef GPS_RTS_NT
#define GPS_RTS_NT 128
#endif
#ifndef GPS_RTS_V
#define GPS_RTS_V 12
#endif

// CSDL with the tile resident in shared memory: one TMA bulk load per CTA, NT threads,
// V uint4 per thread.  <128,8> 837, <128,12> 806-832, <256,8> 821-825, <128,16> 795-810,
// <64,16> 764, <256,4> 619-680 Gkeys/s; rows per pass-2 group 4 (2: same, 8: 640)
#ifndef GPS_SMEM_NT
#define GPS_SMEM_NT 128
#endif
#ifndef GPS_SMEM_V
#define GPS_SMEM_V 8
#endif

// wide elements (uint64 / float64 / vec4<u32>), gps_wide.cuh: NT threads x V (odd) 16-byte
// vectors per thread, tile resident in shared memory; REGS = register budget per thread.
// Measured at 1 GiB in + 1 GiB out on B200 (profiles/r01_wide_sweep.txt), useful GB/s for
// uint64 / vec4<u32>: <256,9> 6508 / 6178 (default), <256,11> 6372 / 6208, <128,15> 6458 / 6039,
// <128,13> 6454 / 5864, <128,9> 6423 / 5357, <128,7> 6045 / 4596, <64,21> 6427 / 5141.
#ifndef GPS_WIDE_NT
#define GPS_WIDE_NT 256
#endif
#ifndef GPS_WIDE_V
#define GPS_WIDE_V 9
#endif
#ifndef GPS_WIDE_REGS
#define GPS_WIDE_REGS 40
#endif
// register budget of the wide Decoupled Fallback kernel
#ifndef GPS_WIDE_FB_REGS
#define GPS_WIDE_FB_REGS 40
#endif

// radix sort digit pass (gps_sort.cuh): NT threads x KPT keys per thread per tile.
// Measured at 2^28 random keys on B200 (profiles/r01_sort_sweep.txt), Gkeys/s keys / pairs:
// <256,16> LB 4: 50.0 / 44.9 (default); LB 1: 48.7 / 42.4; LB 8: 45.7 / 42.3; LB 16: 37.8 / 38.3;
// <256,12>: 38.8 / 36.8; <512,8>: 34.4 (keys); <384,12>: 37.0 (keys);
// cub::DeviceRadixSort (CCCL 2.8.2, onesweep) on the same box: 44.5 / 35.2.
// Later: dynamic shared memory, <256,40> tiles, conflict-free histogram: 65.8 / 50.0; ranks packed
// two per register so that 3 CTAs fit per SM (84 registers): keys <256,40> 74.4 (<256,32> r64 72.5,
// <256,44> 69.4, <256,48> r96 69.9), pairs <256,32> 55.7 (<256,28> 53.9, <256,40> r128 50.0)
#ifndef GPS_SORT_NT
#define GPS_SORT_NT 256
#endif
#ifndef GPS_SORT_KPT
#define GPS_SORT_KPT 40
#endif
#ifndef GPS_SORT_REGS
#define GPS_SORT_REGS 84
#endif
// descriptors fetched per look-back round trip (per digit)
#ifndef GPS_SORT_LB
#define GPS_SORT_LB 4
#endif
// pairs: own tile shape (keys and payload both live in shared memory)
#ifndef GPS_SORT_PAIRS_NT
#define GPS_SORT_PAIRS_NT GPS_SORT_NT
#endif
#ifndef GPS_SORT_PAIRS_KPT
#define GPS_SORT_PAIRS_KPT 32
#endif
#ifndef GPS_SORT_PAIRS_REGS
#define GPS_SORT_PAIRS_REGS 84
#endif

// 64-bit keys (gps_sort_keys64): about half as many keys per thread, so that the keys take about the same
// registers and the tile about the same shared memory as for 32-bit keys.  Measured at 2^27 random uint64
// keys on B200 (tools/gpu_r2_p.sh, profiles/r02_sort64_bench.txt), Gkeys/s keys / pairs:
// <256,24> / <256,20> r84: 24.9 / 20.3 (default); <256,20> / <256,16> r84: 23.6 / 19.0;
// <256,16> / <256,12> r64: 23.2 / 18.2; <256,28> / <256,24> r96: 22.7 / 18.8;
// cub::DeviceRadixSort on the same keys: 12.3 / 15.1
#ifndef GPS_SORT64_NT
#define GPS_SORT64_NT 256
#endif
#ifndef GPS_SORT64_KPT
#define GPS_SORT64_KPT 24
#endif
#ifndef GPS_SORT64_PAIRS_KPT
#define GPS_SORT64_PAIRS_KPT 20
#endif
#ifndef GPS_SORT64_REGS
#define GPS_SORT64_REGS 84
#endif

// CSDL on the blocked shared-memory tile (gps_csdl_blocked.cuh): NT threads x V (odd) uint4
// per thread, TMA bulk load + per-warp TMA bulk store
#ifndef GPS_BLOCKED_NT
#define GPS_BLOCKED_NT 256
#endif
#ifndef GPS_BLOCKED_V
#define GPS_BLOCKED_V 9
#endif
// below this many keys GPS_ALGO_CSDL keeps the lane-strided smem kernel (16 KB tiles fill the
// machine sooner: 2^26 keys 766 vs 738, 2^27 817 vs 806 Gkeys/s in a burst; 2^28 840 vs 844,
// 2^29 853 vs 863, 2^30 850 vs 871 -- profiles/r01_size_sweep.txt)
#ifndef GPS_BLOCKED_MIN_N
#define GPS_BLOCKED_MIN_N (1u << 27)
#endif

// GPS_ALGO_CSDL resolves to: 1 = register path, 2 = TMA ring, 3 = lane-strided smem-resident
// tile, 4 = blocked smem-resident tile.  Chosen by measurement at 2^28 keys (DESIGN.md
// sections 5 and 7): in a 30-launch burst blocked 840, smem 835, register 820, TMA ring 659
// Gkeys/s; over 4000 back-to-back launches (board at its power limit) blocked 822, smem 783,
// register 754 -- the blocked tile executes 0.43 x the instructions
#ifndef GPS_CSDL_DEFAULT_VARIANT
#define GPS_CSDL_DEFAULT_VARIANT 4
#endif

// Ticket-free Decoupled Fallback as GPS_ALGO_CSDL's kernel where it measured faster than plain CSDL with an
// atomic ticket ("ticket_free" = 1, the default; 0 = plain kernels everywhere).  A CTA of a Decoupled Fallback
// kernel never depends on another CTA making progress, so tile = blockIdx.x is deadlock-free under ANY dispatch
// order and the ~0.7 us ticket round trip at the head of every CTA disappears.  us per launch, back to back with
// programmatic dependent launch (tools/gpu_r2_s.sh, gpu_r2_t.sh; profiles/r02_ticket_free.txt), plain + ticket ->
// fallback + blockIdx: 2^20 5.98 -> 5.65, 2^21 7.27 -> 6.57 (register tile <256,4>); 2^23 13.56 -> 12.15,
// 2^24 27.2 -> 26.2, 2^25 48.9 -> 48.1 (smem tile <128,8>); it loses at 2^22 (9.04 -> 9.74), below 2^19 (the
// fallback kernel's extra barriers) and from 2^26 (88.0 -> 89.3), which keep the plain kernels.  Through the
// library inside a CUDA graph (tools/ticket_free_ab.py): 2^20 5.57 -> 5.17, 2^21 7.05 -> 6.20, 2^23 13.25 -> 12.01,
// 3 * 2^22 18.48 -> 16.35, 2^24 24.9 -> 24.9 (cold L2 33.0 -> 32.3), 2^25 46.9 -> 47.0: the upper bound is 2^24.
// Small sizes come first (gps_api.cu, dispatch_scan_tk): a grid of at most ONE wave of the small register-tile
// kernels runs the PLAIN kernel with tile = blockIdx.x -- equally deadlock-free, and cheaper than the fallback
// kernel (2^20: 5.14 us, 2^21: 6.41); the register-tile range below then only applies on devices too small for that.
// "ticket_free" = 1 (default) launches the one-wave kernels WITH the warp-level rescue (gps_common.cuh rescue_tile_warp:
// unconditional forward progress out of place, +0.07..0.25 us), = 2 the lean ones.
// A third register tile, <128,8> (56 registers: 9 CTAs per SM, one wave = 1332 tiles = 5.4 M keys on 148 SMs), takes the
// sizes between one wave of <256,4> (888 tiles = 3.6 M keys) and its own wave, blockIdx order: 3.8 M keys 8.30 -> 8.07 us,
// 2^22 9.01 -> 8.44, 4.5 M 9.58 -> 8.59, 5 M 9.45 -> 9.02 (tools/gpu_r2_ab.sh; profiles/r02_ticket_free.txt G)
#ifndef GPS_REG_MID_NT
#define GPS_REG_MID_NT 128
#endif
#ifndef GPS_REG_MID_V
#define GPS_REG_MID_V 8
#endif
#ifndef GPS_TICKET_FREE_REG_MIN_N
#define GPS_TICKET_FREE_REG_MIN_N ((1u << 19) + 1u)
#endif
#ifndef GPS_TICKET_FREE_REG_MAX_N
#define GPS_TICKET_FREE_REG_MAX_N (1u << 21)
#endif
#ifndef GPS_TICKET_FREE_SMEM_MAX_N  // the smem tile takes over above GPS_REG_SMALL_MAX_N
#define GPS_TICKET_FREE_SMEM_MAX_N (1u << 24)
#endif

// spin budget of GPS_ALGO_CSDL_FALLBACK (no emulated stalls): polls of a not-ready
// window before the CTA recomputes the blocking tile itself
#ifndef GPS_FALLBACK_DEFAULT_SPIN
#define GPS_FALLBACK_DEFAULT_SPIN 64
#endif

// L2 cache policies on the tile loads (evict_first) and stores (evict_last) of the smem-resident
// CSDL kernels, option "l2_hints" = -1 (auto): applied while the output is at most this many
// percent of the device's L2 (and at least MIN percent); which bits auto uses
// Measured (tools/l2_hint_sweep.py, profiles/r02_midsize.txt): evict_first loads help only where
// input + output are about the size of the L2 -- 2^24 keys: 26.0 -> 24.4 us (the output then stays
// in L2 and is overwritten there by the next scan); 2^23: 13.3 -> 15.0 us (worse), 2^25 and up:
// neutral to -3 %; evict_last stores lose everywhere.  Hence a band around 2^24.
#ifndef GPS_L2_HINT_MIN_OUT_PCT
#define GPS_L2_HINT_MIN_OUT_PCT 38
#endif
#ifndef GPS_L2_HINT_MAX_OUT_PCT
#define GPS_L2_HINT_MAX_OUT_PCT 76
#endif
#ifndef GPS_L2_HINT_AUTO_BITS
#define GPS_L2_HINT_AUTO_BITS 1
#endif

// programmatic dependent launch inside the sharded step (gps_reduce_publish + gps_scan_carry_slots):
// 0 = plain launches, 1 = only the one-block fold, 2 = fold and scan (option "sharded_pdl")
#ifndef GPS_SHARDED_PDL_DEFAULT
#define GPS_SHARDED_PDL_DEFAULT 1  // measured on 2 GPUs: 0.4740 / 0.4727 / 0.4804 ms per step for 0 / 1 / 2
#endif

// L2 policies of the sharded step (option "sharded_l2", bits: 1 = reduce keeps the front of the slice,
// 2 = scan stores evict_first, 4 = scan loads evict_first) and how much of the front is kept
#ifndef GPS_SHARDED_L2_DEFAULT
#define GPS_SHARDED_L2_DEFAULT 0
#endif
#ifndef GPS_SHARDED_L2_KEEP_MB
#define GPS_SHARDED_L2_KEEP_MB 96
#endif
