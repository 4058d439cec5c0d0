// gps_csdl_tma.cuh -- Chained Scan with Decoupled Lookback, TMA path (sm_100a).
//
// Same algorithm and descriptor protocol as gps_csdl_reg.cuh (replacing
// ChainedScanDecoupledLookback.cuh:79-139 of the reference) but the tile never moves
// through the LSU on its way in or out of the SM:
//   * persistent CTAs claim tiles with an atomic counter and keep a ring of STAGES
//     shared-memory tile buffers; one elected thread issues `cp.async.bulk`
//     global->shared copies (SASS UBLKCP) for the tiles it claimed, completion tracked
//     by one mbarrier per stage, so loads for the next tiles are always in flight while
//     the current tile is scanned / waits on its look-back;
//   * every thread owns 16 CONSECUTIVE keys (64 B) of the tile: one serial pass + ONE
//     warp scan per 512 keys (the register path needs one warp scan per 128 keys); the
//     four 16-byte chunks are read in a per-thread rotated order so that every
//     quarter-warp LDS.128/STS.128 phase hits 32 distinct banks, and rotated back in
//     registers;
//   * results are written back into the same stage, fenced to the async proxy and
//     stored with one `cp.async.bulk` shared->global per tile.
// Forward progress: a tile is only ever claimed by a CTA that is already running, and
// a CTA processes its claims in increasing order, so the smallest unfinished tile can
// always complete -- no co-residency assumption on the grid.
#pragma once
#include "gps_common.cuh"

namespace gps {

template <int NT, int STAGES>
struct TmaCfg {
    static constexpr int IPT = 16;                       // keys per thread
    static constexpr int CHUNKS = IPT / 4;               // 16-byte chunks per thread
    static constexpr uint32_t TILE_KEYS = NT * IPT;
    static constexpr uint32_t TILE_BYTES = TILE_KEYS * 4;
    static constexpr uint32_t RING_BYTES = TILE_BYTES * STAGES;
};

// rotate left by `rot` chunks: out[m] = in[(m + rot) & 3], done with selects so that all
// register indices stay static
__device__ __forceinline__ void rotate4(uint4 (&r)[4], uint32_t rot)
{
    const bool b0 = rot & 1u, b1 = rot & 2u;
    uint4 a[4];
#pragma unroll
    for (int m = 0; m < 4; ++m) {
        const uint4 p = r[(m + 1) & 3], q = r[m];
        a[m] = make_uint4(b0 ? p.x : q.x, b0 ? p.y : q.y, b0 ? p.z : q.z, b0 ? p.w : q.w);
    }
#pragma unroll
    for (int m = 0; m < 4; ++m) {
        const uint4 p = a[(m + 2) & 3], q = a[m];
        r[m] = make_uint4(b1 ? p.x : q.x, b1 ? p.y : q.y, b1 ? p.z : q.z, b1 ? p.w : q.w);
    }
}

template <typename T, int NT, int STAGES, bool INCLUSIVE>
__global__ void __launch_bounds__(NT) csdl_tma_kernel(const __grid_constant__ ScanParams p)
{
    using Cfg = TmaCfg<NT, STAGES>;
    constexpr int NW = NT / 32;
    constexpr int IPT = Cfg::IPT;
    static_assert(Cfg::CHUNKS == 4, "rotation network is written for 4 chunks");

    extern __shared__ __align__(128) unsigned char ring[];
    __shared__ __align__(8) uint64_t s_full[STAGES];
    __shared__ uint32_t s_tile[STAGES];
    __shared__ T s_warp[NW];
    __shared__ T s_prefix;

    const uint32_t tid = threadIdx.x;
    const uint32_t lane = tid & 31u;
    const uint32_t warp = tid >> 5;
    const uint32_t num_tiles = p.num_tiles;
    const uint32_t full_tiles = (uint32_t)(p.n / Cfg::TILE_KEYS);
    const unsigned char* in_bytes = reinterpret_cast<const unsigned char*>(p.in);
    unsigned char* out_bytes = reinterpret_cast<unsigned char*>(p.out);

    // producer (thread 0): claim a tile for `stage`; full tiles are fetched by TMA, the
    // partial last tile and the "no more work" marker only complete the barrier phase
    auto produce = [&](uint32_t stage) {
        const uint32_t t = claim_tile(p);
        s_tile[stage] = t;
        if (t < full_tiles) {
            mbar_arrive_expect_tx(&s_full[stage], Cfg::TILE_BYTES);
            bulk_load(ring + (size_t)stage * Cfg::TILE_BYTES,
                      in_bytes + (uint64_t)t * Cfg::TILE_BYTES, Cfg::TILE_BYTES, &s_full[stage]);
        } else {
            mbar_arrive(&s_full[stage]);
        }
    };

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; ++s) mbar_init(&s_full[s], 1);
        fence_mbar_init();
    }
    __syncthreads();
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < STAGES - 1; ++s) produce((uint32_t)s);
    }

    // rotation that makes the 64-byte-stride LDS.128/STS.128 conflict-free: within a
    // quarter warp threads {0,2,4,6} (and {1,3,5,7}) share a 16-bank half, so they start
    // at different chunks
    const uint32_t rot = (tid >> 1) & 3u;

    const bool profiler = p.prof != nullptr && tid == 0;
    for (uint32_t it = 0;; ++it) {
        long long tprof[6] = {0, 0, 0, 0, 0, 0};
        LookbackStats lbs;
        if (profiler) tprof[0] = tprof[1] = clock64();
        const uint32_t stage = it % STAGES;
        const uint32_t parity = (it / STAGES) & 1u;
        if (!mbar_wait_bounded(&s_full[stage], parity)) {  // never expected; do not hang the GPU
            atomicOr(p.status, STATUS_WATCHDOG);
            return;
        }
        const uint32_t tile = s_tile[stage];
        if (tile >= num_tiles) break;
        if (profiler) tprof[2] = clock64();
        const bool full = tile < full_tiles;
        unsigned char* buf = ring + (size_t)stage * Cfg::TILE_BYTES;
        uint4* mine = reinterpret_cast<uint4*>(buf + (size_t)tid * (IPT * 4));
        const uint64_t key0 = (uint64_t)tile * Cfg::TILE_KEYS + (uint64_t)tid * IPT;

        T x[IPT];
        if (full) {
            uint4 r[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) r[j] = mine[(j + rot) & 3u];
            // r[j] holds chunk (j + rot); chunk m is r[(m - rot) & 3] = rotate by (4 - rot)
            rotate4(r, (4u - rot) & 3u);
#pragma unroll
            for (int m = 0; m < 4; ++m) {
                x[4 * m + 0] = from_bits<T>(r[m].x);
                x[4 * m + 1] = from_bits<T>(r[m].y);
                x[4 * m + 2] = from_bits<T>(r[m].z);
                x[4 * m + 3] = from_bits<T>(r[m].w);
            }
        } else {
            const uint32_t* in1 = reinterpret_cast<const uint32_t*>(p.in);
#pragma unroll
            for (int i = 0; i < IPT; ++i)
                x[i] = (key0 + i < p.n) ? from_bits<T>(in1[key0 + i]) : T(0);
        }

        T tsum = x[0];
#pragma unroll
        for (int i = 1; i < IPT; ++i) tsum += x[i];
        const T inc = warp_inclusive_scan(tsum, lane);
        const T lane_excl = warp_exclusive_of(inc, tsum, lane);
        if (lane == 31) s_warp[warp] = inc;
        __syncthreads();  // (A)

        T warp_prefix, aggregate;
        block_scan_warp_totals<T, NW>(s_warp, warp, lane, warp_prefix, aggregate);

        if (warp == 0) {
            T prefix;
            if (tile == 0) {
                prefix = acquire_carry<T>(p, lane);
                if (profiler) tprof[3] = clock64();
            } else {
                if (lane == 0)
                    desc_publish(p, p.desc + tile, pack_desc(to_bits(aggregate), p.epoch, ST_AGGREGATE));
                if (profiler) tprof[3] = clock64();
                prefix = lookback_warp<T>(p, tile, lane, &lbs);
            }
            if (profiler) tprof[4] = clock64();
            if (lane == 0) {
                const T inclusive = prefix + aggregate;
                desc_publish(p, p.desc + tile, pack_desc(to_bits(inclusive), p.epoch, ST_INCLUSIVE));
                s_prefix = prefix;
                if (tile == num_tiles - 1) write_total<T>(p, inclusive);
            }
        }
        __syncthreads();  // (B)

        T run = (s_prefix + warp_prefix) + lane_excl;
#pragma unroll
        for (int i = 0; i < IPT; ++i) {
            const T in_i = x[i];
            if (INCLUSIVE) {
                run += in_i;
                x[i] = run;
            } else {
                x[i] = run;
                run += in_i;
            }
        }

        if (full) {
            uint4 r[4];
#pragma unroll
            for (int m = 0; m < 4; ++m)
                r[m] = make_uint4(to_bits(x[4 * m + 0]), to_bits(x[4 * m + 1]), to_bits(x[4 * m + 2]),
                                  to_bits(x[4 * m + 3]));
            rotate4(r, rot);  // r[j] = chunk (j + rot)
#pragma unroll
            for (int j = 0; j < 4; ++j) mine[(j + rot) & 3u] = r[j];
            fence_proxy_async_smem();
        } else {
            uint32_t* out1 = reinterpret_cast<uint32_t*>(p.out);
#pragma unroll
            for (int i = 0; i < IPT; ++i)
                if (key0 + i < p.n) out1[key0 + i] = to_bits(x[i]);
        }
        __syncthreads();  // (C) all result chunks are in the stage and fenced

        if (tid == 0) {
            if (full) {
                bulk_store(out_bytes + (uint64_t)tile * Cfg::TILE_BYTES, buf, Cfg::TILE_BYTES);
                bulk_commit();
            }
            // the stage of iteration it-1 is refilled for iteration it+STAGES-1 once its
            // store has finished READING shared memory (all but the newest group)
            bulk_wait_read<1>();
            produce((it + STAGES - 1) % STAGES);
            if (profiler) {
                tprof[5] = clock64();
                prof_add(p, tile, tprof, lbs);
            }
        }
    }
    if (tid == 0) bulk_wait<0>();
}

// ---------------------------------------------------------------------------------
// TMA path, version 3: early refill.
//
// LOADS load buffers + ONE store buffer per CTA.  The moment every thread has pulled its
// 16 keys of the tile into registers (the barrier after the thread sums), the load buffer
// is handed straight back to the TMA engine for the tile this CTA will process LOADS
// iterations later -- so a load is in flight during the whole scan / look-back / store of
// the current tile even with a single load buffer.  Results go to the separate store
// buffer and leave with one bulk store.  A dedicated thread in the LAST warp drives the
// TMA engine so warp 0's look-back is never held up by a bulk-group wait.
// STAGES (template) = LOADS + 1 buffers in shared memory.
// ---------------------------------------------------------------------------------
template <typename T, int NT, int STAGES, bool INCLUSIVE>
__global__ void __launch_bounds__(NT) csdl_tma3_kernel(const __grid_constant__ ScanParams p)
{
    using Cfg = TmaCfg<NT, STAGES>;
    constexpr int NW = NT / 32;
    constexpr int IPT = Cfg::IPT;
    constexpr int LOADS = STAGES - 1;
    static_assert(LOADS >= 1, "at least one load buffer and one store buffer");

    extern __shared__ __align__(128) unsigned char ring[];  // [LOADS load buffers][store buffer]
    __shared__ __align__(8) uint64_t s_full[LOADS];
    __shared__ uint32_t s_tile[LOADS];
    __shared__ T s_warp[2][NW];
    __shared__ T s_prefix;

    const uint32_t tid = threadIdx.x;
    const uint32_t lane = tid & 31u;
    const uint32_t warp = tid >> 5;
    const bool tma_thread = tid == (uint32_t)(NT - 32);  // lane 0 of the last warp
    const uint32_t num_tiles = p.num_tiles;
    const uint32_t full_tiles = (uint32_t)(p.n / Cfg::TILE_KEYS);
    const unsigned char* in_bytes = reinterpret_cast<const unsigned char*>(p.in);
    unsigned char* out_bytes = reinterpret_cast<unsigned char*>(p.out);
    unsigned char* store_buf = ring + (size_t)LOADS * Cfg::TILE_BYTES;
    const uint32_t rot = (tid >> 1) & 3u;

    auto produce = [&](uint32_t stage) {
        const uint32_t t = claim_tile(p);
        s_tile[stage] = t;
        if (t < full_tiles) {
            mbar_arrive_expect_tx(&s_full[stage], Cfg::TILE_BYTES);
            bulk_load(ring + (size_t)stage * Cfg::TILE_BYTES,
                      in_bytes + (uint64_t)t * Cfg::TILE_BYTES, Cfg::TILE_BYTES, &s_full[stage]);
        } else {
            mbar_arrive(&s_full[stage]);
        }
    };

    if (tma_thread) {
#pragma unroll
        for (int s = 0; s < LOADS; ++s) mbar_init(&s_full[s], 1);
        fence_mbar_init();
    }
    __syncthreads();
    if (tma_thread) {
#pragma unroll
        for (int s = 0; s < LOADS; ++s) produce((uint32_t)s);
    }

    for (uint32_t it = 0;; ++it) {
        const uint32_t stage = it % LOADS;
        if (!mbar_wait_bounded(&s_full[stage], (it / LOADS) & 1u)) {
            atomicOr(p.status, STATUS_WATCHDOG);
            return;
        }
        const uint32_t tile = s_tile[stage];
        if (tile >= num_tiles) break;
        const bool full = tile < full_tiles;
        const uint64_t key0 = (uint64_t)tile * Cfg::TILE_KEYS + (uint64_t)tid * IPT;

        T x[IPT];
        if (full) {
            const uint4* mine =
                reinterpret_cast<const uint4*>(ring + (size_t)stage * Cfg::TILE_BYTES + (size_t)tid * (IPT * 4));
            uint4 r[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) r[j] = mine[(j + rot) & 3u];
            rotate4(r, (4u - rot) & 3u);
#pragma unroll
            for (int m = 0; m < 4; ++m) {
                x[4 * m + 0] = from_bits<T>(r[m].x);
                x[4 * m + 1] = from_bits<T>(r[m].y);
                x[4 * m + 2] = from_bits<T>(r[m].z);
                x[4 * m + 3] = from_bits<T>(r[m].w);
            }
        } else {
            const uint32_t* in1 = reinterpret_cast<const uint32_t*>(p.in);
#pragma unroll
            for (int i = 0; i < IPT; ++i) x[i] = (key0 + i < p.n) ? from_bits<T>(in1[key0 + i]) : T(0);
        }
        T tsum = x[0];
#pragma unroll
        for (int i = 1; i < IPT; ++i) tsum += x[i];
        const T inc = warp_inclusive_scan(tsum, lane);
        const T lane_excl = warp_exclusive_of(inc, tsum, lane);
        T* sw = s_warp[it & 1u];
        if (lane == 31) sw[warp] = inc;
        __syncthreads();  // (A) warp totals visible; every thread is done READING the load buffer

        if (tma_thread) {
            produce(stage);        // hand the buffer back: load for iteration it + LOADS
            bulk_wait_read<0>();   // the previous tile's store has left the store buffer
        }
        T warp_prefix, aggregate;
        block_scan_warp_totals<T, NW>(sw, warp, lane, warp_prefix, aggregate);
        if (warp == 0) {
            T prefix;
            if (tile == 0) {
                prefix = acquire_carry<T>(p, lane);
            } else {
                if (lane == 0)
                    desc_publish(p, p.desc + tile, pack_desc(to_bits(aggregate), p.epoch, ST_AGGREGATE));
                prefix = lookback_warp<T>(p, tile, lane);
            }
            if (lane == 0) {
                const T inclusive = prefix + aggregate;
                desc_publish(p, p.desc + tile, pack_desc(to_bits(inclusive), p.epoch, ST_INCLUSIVE));
                s_prefix = prefix;
                if (tile == num_tiles - 1) write_total<T>(p, inclusive);
            }
        }
        __syncthreads();  // (B) prefix visible; store buffer is free

        T run = (s_prefix + warp_prefix) + lane_excl;
#pragma unroll
        for (int i = 0; i < IPT; ++i) {
            const T in_i = x[i];
            if (INCLUSIVE) {
                run += in_i;
                x[i] = run;
            } else {
                x[i] = run;
                run += in_i;
            }
        }
        if (full) {
            uint4* mine = reinterpret_cast<uint4*>(store_buf + (size_t)tid * (IPT * 4));
            uint4 r[4];
#pragma unroll
            for (int m = 0; m < 4; ++m)
                r[m] = make_uint4(to_bits(x[4 * m + 0]), to_bits(x[4 * m + 1]), to_bits(x[4 * m + 2]),
                                  to_bits(x[4 * m + 3]));
            rotate4(r, rot);
#pragma unroll
            for (int j = 0; j < 4; ++j) mine[(j + rot) & 3u] = r[j];
            fence_proxy_async_smem();
        } else {
            uint32_t* out1 = reinterpret_cast<uint32_t*>(p.out);
#pragma unroll
            for (int i = 0; i < IPT; ++i)
                if (key0 + i < p.n) out1[key0 + i] = to_bits(x[i]);
        }
        __syncthreads();  // (C) results are in the store buffer and fenced
        if (tma_thread && full) {
            bulk_store(out_bytes + (uint64_t)tile * Cfg::TILE_BYTES, store_buf, Cfg::TILE_BYTES);
            bulk_commit();
        }
    }
    if (tma_thread) bulk_wait<0>();
}

// ---------------------------------------------------------------------------------
// TMA path, version 2: "aggregate ahead".
//
// A persistent CTA that handles its claimed tiles strictly one after the other delays
// the AGGREGATE of tile i+1 until the look-back of tile i has finished -- and every
// successor of tile i+1 in the grid waits for that aggregate (measured: deeper prefetch
// made version 1 SLOWER).  Version 2 splits the work per tile in two phases and runs
// them software-pipelined:
//   phase A(tile): wait for the TMA load, pull the thread's 16 keys into registers,
//                  thread/warp/block reduce, PUBLISH THE AGGREGATE;
//   phase B(tile): look-back, add the prefix, write the results back into the stage,
//                  TMA store.
// Order per CTA:  A(t0)  A(t1) B(t0)  A(t2) B(t1)  ...   so an aggregate is visible as
// soon as its data has landed, and each look-back has had a whole tile period of head
// start by the time the CTA blocks on it.  The keys of the tile between its A and B
// phase live in registers (two tiles per thread at the hand-over).
// ---------------------------------------------------------------------------------
template <typename T, int NT, int STAGES, bool INCLUSIVE>
__global__ void __launch_bounds__(NT) csdl_tma2_kernel(const __grid_constant__ ScanParams p)
{
    using Cfg = TmaCfg<NT, STAGES>;
    constexpr int NW = NT / 32;
    constexpr int IPT = Cfg::IPT;
    static_assert(STAGES >= 3, "needs a stage each for: current tile, next tile, refill");

    extern __shared__ __align__(128) unsigned char ring[];
    __shared__ __align__(8) uint64_t s_full[STAGES];
    __shared__ uint32_t s_tile[STAGES];
    __shared__ T s_warp[2][NW];
    __shared__ T s_prefix;

    const uint32_t tid = threadIdx.x;
    const uint32_t lane = tid & 31u;
    const uint32_t warp = tid >> 5;
    const uint32_t num_tiles = p.num_tiles;
    const uint32_t full_tiles = (uint32_t)(p.n / Cfg::TILE_KEYS);
    const unsigned char* in_bytes = reinterpret_cast<const unsigned char*>(p.in);
    unsigned char* out_bytes = reinterpret_cast<unsigned char*>(p.out);
    const uint32_t rot = (tid >> 1) & 3u;

    auto produce = [&](uint32_t stage) {
        const uint32_t t = claim_tile(p);
        s_tile[stage] = t;
        if (t < full_tiles) {
            mbar_arrive_expect_tx(&s_full[stage], Cfg::TILE_BYTES);
            bulk_load(ring + (size_t)stage * Cfg::TILE_BYTES,
                      in_bytes + (uint64_t)t * Cfg::TILE_BYTES, Cfg::TILE_BYTES, &s_full[stage]);
        } else {
            mbar_arrive(&s_full[stage]);
        }
    };

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; ++s) mbar_init(&s_full[s], 1);
        fence_mbar_init();
    }
    __syncthreads();
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < STAGES - 1; ++s) produce((uint32_t)s);
    }

    // per-thread state of a tile between its phase A and phase B
    struct Held {
        T x[IPT];
        T lane_excl, warp_prefix, aggregate;
        uint32_t tile;
    };

    // phase A; returns false (block-uniform) when the claim was past the end.  `slot`
    // selects the s_warp buffer so consecutive phases never share one.
    auto phase_a = [&](uint32_t it, Held& h) -> bool {
        const uint32_t stage = it % STAGES;
        if (!mbar_wait_bounded(&s_full[stage], (it / STAGES) & 1u)) {
            atomicOr(p.status, STATUS_WATCHDOG);
            h.tile = 0xFFFFFFFFu;
            return false;
        }
        h.tile = s_tile[stage];
        if (h.tile >= num_tiles) return false;
        const bool full = h.tile < full_tiles;
        if (full) {
            const uint4* mine =
                reinterpret_cast<const uint4*>(ring + (size_t)stage * Cfg::TILE_BYTES + (size_t)tid * (IPT * 4));
            uint4 r[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) r[j] = mine[(j + rot) & 3u];
            rotate4(r, (4u - rot) & 3u);
#pragma unroll
            for (int m = 0; m < 4; ++m) {
                h.x[4 * m + 0] = from_bits<T>(r[m].x);
                h.x[4 * m + 1] = from_bits<T>(r[m].y);
                h.x[4 * m + 2] = from_bits<T>(r[m].z);
                h.x[4 * m + 3] = from_bits<T>(r[m].w);
            }
        } else {
            const uint32_t* in1 = reinterpret_cast<const uint32_t*>(p.in);
            const uint64_t key0 = (uint64_t)h.tile * Cfg::TILE_KEYS + (uint64_t)tid * IPT;
#pragma unroll
            for (int i = 0; i < IPT; ++i) h.x[i] = (key0 + i < p.n) ? from_bits<T>(in1[key0 + i]) : T(0);
        }
        T tsum = h.x[0];
#pragma unroll
        for (int i = 1; i < IPT; ++i) tsum += h.x[i];
        const T inc = warp_inclusive_scan(tsum, lane);
        h.lane_excl = warp_exclusive_of(inc, tsum, lane);
        T* sw = s_warp[it & 1u];
        if (lane == 31) sw[warp] = inc;
        __syncthreads();
        block_scan_warp_totals<T, NW>(sw, warp, lane, h.warp_prefix, h.aggregate);
        if (warp == 0 && lane == 0 && h.tile != 0)
            desc_publish(p, p.desc + h.tile, pack_desc(to_bits(h.aggregate), p.epoch, ST_AGGREGATE));
        return true;
    };

    Held cur;
    bool have_cur = phase_a(0, cur);
    for (uint32_t it = 0; have_cur; ++it) {
        Held nxt;
        const bool have_next = phase_a(it + 1, nxt);

        // ---- phase B of the current tile
        const uint32_t stage = it % STAGES;
        const uint32_t tile = cur.tile;
        const bool full = tile < full_tiles;
        if (warp == 0) {
            const T prefix = tile == 0 ? acquire_carry<T>(p, lane) : lookback_warp<T>(p, tile, lane);
            if (lane == 0) {
                const T inclusive = prefix + cur.aggregate;
                desc_publish(p, p.desc + tile, pack_desc(to_bits(inclusive), p.epoch, ST_INCLUSIVE));
                s_prefix = prefix;
                if (tile == num_tiles - 1) write_total<T>(p, inclusive);
            }
        }
        __syncthreads();

        T run = (s_prefix + cur.warp_prefix) + cur.lane_excl;
#pragma unroll
        for (int i = 0; i < IPT; ++i) {
            const T in_i = cur.x[i];
            if (INCLUSIVE) {
                run += in_i;
                cur.x[i] = run;
            } else {
                cur.x[i] = run;
                run += in_i;
            }
        }
        unsigned char* buf = ring + (size_t)stage * Cfg::TILE_BYTES;
        if (full) {
            uint4* mine = reinterpret_cast<uint4*>(buf + (size_t)tid * (IPT * 4));
            uint4 r[4];
#pragma unroll
            for (int m = 0; m < 4; ++m)
                r[m] = make_uint4(to_bits(cur.x[4 * m + 0]), to_bits(cur.x[4 * m + 1]),
                                  to_bits(cur.x[4 * m + 2]), to_bits(cur.x[4 * m + 3]));
            rotate4(r, rot);
#pragma unroll
            for (int j = 0; j < 4; ++j) mine[(j + rot) & 3u] = r[j];
            fence_proxy_async_smem();
        } else {
            uint32_t* out1 = reinterpret_cast<uint32_t*>(p.out);
            const uint64_t key0 = (uint64_t)tile * Cfg::TILE_KEYS + (uint64_t)tid * IPT;
#pragma unroll
            for (int i = 0; i < IPT; ++i)
                if (key0 + i < p.n) out1[key0 + i] = to_bits(cur.x[i]);
        }
        __syncthreads();  // results are in the stage and fenced; s_prefix may be rewritten

        if (tid == 0) {
            if (full) {
                bulk_store(out_bytes + (uint64_t)tile * Cfg::TILE_BYTES, buf, Cfg::TILE_BYTES);
                bulk_commit();
            }
            // refill the stage of iteration it-1 (its store is the second newest group)
            // with the claim for iteration it+STAGES-1
            bulk_wait_read<1>();
            produce((it + STAGES - 1) % STAGES);
        }
        cur = nxt;
        have_cur = have_next;
    }
    if (tid == 0) bulk_wait<0>();
}

}  // namespace gps
