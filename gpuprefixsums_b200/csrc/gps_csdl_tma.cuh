// gps_csdl_tma.cuh -- Chained Scan with Decoupled Lookback, TMA path (sm_100a).
//
// Same algorithm and descriptor protocol as gps_csdl_reg.cuh (replacing
// ChainedScanDecoupledLookback.cuh:79-139 of the reference) but the tile never moves
// through the LSU on its way in or out of the SM:
//   * persistent CTAs claim tiles with an atomic counter; one elected thread issues a
//     `cp.async.bulk` global->shared copy (SASS UBLKCP) per claimed tile, completion
//     tracked by an mbarrier, into a ring of STAGES shared-memory tile buffers;
//   * every thread owns 16 CONSECUTIVE keys (64 B) of the tile: one serial pass + ONE
//     warp scan per 512 keys (the register path needs one warp scan per 128 keys); the
//     four 16-byte chunks are read in a per-thread rotated order so that every
//     quarter-warp LDS.128/STS.128 phase hits 32 distinct banks, and rotated back in
//     registers with selects;
//   * the tile AGGREGATE is published as early as possible -- thread sums, one REDUX per
//     warp, one barrier, one REDUX in warp 0 -- and the warp/block scan the outputs need
//     is done by the other warps while warp 0 walks the look-back;
//   * results are written back into the same stage, fenced to the async proxy and leave
//     with one `cp.async.bulk` shared->global per tile.
//
// Measured design rules (B200, 2^28 keys; numbers in DESIGN.md / profiles/):
//   * claim -> load -> aggregate must be as short as possible: every tile that is claimed
//     but not yet published is a hole all later tiles spin on.  Prefetching a CTA's NEXT
//     tile early (deeper rings, early refill, aggregate-ahead pipelines were all tried)
//     lengthens the holes and LOWERS throughput; overlap comes from several CTAs per SM
//     instead.  STAGES = 2 is a load/store ping-pong: the bulk store of tile i drains
//     from one buffer while the load of tile i+1 fills the other.
//   * descriptors are published with atom.exch (performed at L2 at once), not st.release.
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

// WIN: look-back windows per round trip; PROF: compile the phase timers in (tuning only)
template <typename T, int NT, int STAGES, bool INCLUSIVE, int WIN = 2, bool PROF = false>
__global__ void __launch_bounds__(NT, min_ctas_per_sm(NT)) csdl_tma_kernel(const __grid_constant__ ScanParams p)
{
    using Cfg = TmaCfg<NT, STAGES>;
    constexpr int NW = NT / 32;
    constexpr int IPT = Cfg::IPT;
    static_assert(Cfg::CHUNKS == 4, "rotation network is written for 4 chunks");
    static_assert(STAGES >= 2, "load/store ping-pong needs two buffers");

    extern __shared__ __align__(128) unsigned char ring[];
    __shared__ __align__(8) uint64_t s_full[STAGES];
    __shared__ uint32_t s_tile[STAGES];
    __shared__ T s_wsum[2][NW];  // per-warp sums, double-buffered across iterations
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
    const bool profiler = PROF && p.prof != nullptr && tid == 0;

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

        // ---- aggregate first: thread sums (4 independent chains), REDUX per warp
        T tsum;
        {
            T c0 = (x[0] + x[1]) + (x[2] + x[3]);
            T c1 = (x[4] + x[5]) + (x[6] + x[7]);
            T c2 = (x[8] + x[9]) + (x[10] + x[11]);
            T c3 = (x[12] + x[13]) + (x[14] + x[15]);
            tsum = (c0 + c1) + (c2 + c3);
        }
        const T wsum = warp_reduce(tsum);
        T* sw = s_wsum[it & 1u];
        if (lane == 0) sw[warp] = wsum;
        __syncthreads();  // (A)

        // every warp: its exclusive prefix within the tile + the tile aggregate
        const T wraw = lane < (uint32_t)NW ? sw[lane] : T(0);
        const T aggregate = warp_reduce(wraw);
        if (warp == 0 && lane == 0 && tile != 0)
            desc_publish(p, p.desc + tile, pack_desc(to_bits(aggregate), p.epoch, ST_AGGREGATE));
        if (profiler) tprof[3] = clock64();

        // scan work that does not need the tile prefix (overlaps warp 0's look-back for
        // the other warps; warp 0 does it before its first poll)
        const T winc = warp_inclusive_scan(wraw, lane);
        const T warp_prefix = __shfl_sync(kFullMask, warp_exclusive_of(winc, wraw, lane), (int)warp);
        const T inc = warp_inclusive_scan(tsum, lane);
        const T lane_excl = warp_exclusive_of(inc, tsum, lane);

        if (warp == 0) {
            const T prefix = tile == 0 ? acquire_carry<T>(p, lane)
                                       : lookback_warp<T, WIN>(p, tile, lane, PROF ? &lbs : nullptr);
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

}  // namespace gps
