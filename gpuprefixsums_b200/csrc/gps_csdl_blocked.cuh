// gps_csdl_blocked.cuh -- Chained Scan with Decoupled Lookback on a BLOCKED shared-memory tile.
//
// Fourth variant of the CSDL tile pipeline (same algorithm, descriptor protocol, ticket and
// carry handling as gps_csdl_smem.cuh; replaces ChainedScanDecoupledLookback.cuh:79-139 of the
// reference).  It came out of the wide-element work (gps_wide.cuh) and differs from the
// lane-strided smem kernel in how a thread walks the tile:
//   * one TMA bulk load brings the tile into shared memory (as before);
//   * thread t owns the V consecutive uint4 [t*V, (t+1)*V) of the tile (V odd: the LDS.128 /
//     STS.128 of a quarter-warp hit 8 distinct 16-byte bank groups, no padding needed), sums
//     them, and ONE warp scan of the thread totals per tile replaces the V per-row warp scans;
//   * pass 2 scans the thread's run serially, in place in shared memory, and each warp's
//     contiguous 32*V uint4 leave with one TMA bulk store (UBLKCP.G.S).
// 52 M warp-instructions per 2^28 keys instead of 120 M.  In a 30-launch burst both kernels
// run at the same speed (840 vs 835 Gkeys/s: HBM bound); what the lower instruction count buys
// is POWER: the board sits at its power limit while scanning, and over 500 / 4000 back-to-back
// launches this kernel holds 837 / 822 Gkeys/s where the lane-strided one drops to 814 / 783
// (tools/sustained_bench.py).
#pragma once
#include "gps_common.cuh"

namespace gps {

// ---- the blocked tile: pieces shared by the chained scan, RTS propagate and the fallback kernel
template <typename T, int NT, int V>
struct BlockedTile {
    static_assert(V % 2 == 1, "V must be odd: conflict-free LDS.128 at stride V*16 bytes");
    static constexpr int NW = NT / 32;
    static constexpr uint32_t TILE_VEC = NT * V;
    static constexpr uint32_t TILE_KEYS = TILE_VEC * 4;
    static constexpr uint32_t TILE_BYTES = TILE_KEYS * 4;
    static constexpr uint32_t WARP_VEC = 32 * V;

    // called by ONE thread: arm the mbarrier and start the TMA bulk load of a full tile
    __device__ __forceinline__ static void issue_load(uint4* s_data, uint64_t* s_bar, const void* in, uint32_t tile,
                                                      uint64_t n, uint32_t flags = 0u)
    {
        if (((uint64_t)tile + 1u) * TILE_KEYS <= n) {
            mbar_arrive_expect_tx(s_bar, TILE_BYTES);
            tile_bulk_load(flags, s_data, reinterpret_cast<const unsigned char*>(in) + (uint64_t)tile * TILE_BYTES,
                           TILE_BYTES, s_bar);
        }
    }
    // all threads: wait for the TMA (full tile) or fill this thread's own uint4 with guarded
    // loads (partial last tile, keys at index >= n read as 0).  false = the TMA never landed.
    __device__ __forceinline__ static bool arrive(uint4* s_data, uint64_t* s_bar, const void* in, uint32_t tile,
                                                  uint64_t n, bool full, uint32_t v0)
    {
        if (full) return mbar_wait_bounded(s_bar, 0u);
        const uint32_t* in1 = reinterpret_cast<const uint32_t*>(in);
#pragma unroll 1
        for (int k = 0; k < V; ++k) {
            const uint64_t e = ((uint64_t)tile * TILE_VEC + v0 + k) * 4u;
            uint4 r;
            r.x = e + 0 < n ? in1[e + 0] : 0u;
            r.y = e + 1 < n ? in1[e + 1] : 0u;
            r.z = e + 2 < n ? in1[e + 2] : 0u;
            r.w = e + 3 < n ? in1[e + 3] : 0u;
            s_data[v0 + k] = r;
        }
        return true;
    }
    // pass 1: sum of the thread's V consecutive uint4
    __device__ __forceinline__ static T thread_sum(const uint4* s_data, uint32_t v0)
    {
        T acc = T(0);
#pragma unroll
        for (int k = 0; k < V; ++k) {
            const uint4 r = s_data[v0 + k];
            acc += (from_bits<T>(r.x) + from_bits<T>(r.y)) + (from_bits<T>(r.z) + from_bits<T>(r.w));
        }
        return acc;
    }
    // pass 2: serial scan of the thread's run in place in shared memory (`run` = exclusive
    // prefix of the thread's first key), then the warp's contiguous 32*V uint4 leave with one
    // TMA bulk store (full tile) or a guarded coalesced copy (partial tile)
    template <bool INCLUSIVE>
    __device__ __forceinline__ static void finish(uint4* s_data, uint32_t v0, T run, void* out, uint32_t tile,
                                                  uint64_t n, bool full, uint32_t warp, uint32_t lane, uint32_t flags = 0u)
    {
#pragma unroll
        for (int k = 0; k < V; ++k) {
            const uint4 r = s_data[v0 + k];
            const T x0 = from_bits<T>(r.x), x1 = from_bits<T>(r.y), x2 = from_bits<T>(r.z), x3 = from_bits<T>(r.w);
            T o0, o1, o2, o3;
            if (INCLUSIVE) {
                o0 = run + x0;
                o1 = o0 + x1;
                o2 = o1 + x2;
                o3 = o2 + x3;
                run = o3;
            } else {
                o0 = run;
                o1 = o0 + x0;
                o2 = o1 + x1;
                o3 = o2 + x2;
                run = o3 + x3;
            }
            s_data[v0 + k] = make_uint4(to_bits(o0), to_bits(o1), to_bits(o2), to_bits(o3));
        }
        const uint64_t warp_vec0 = (uint64_t)tile * TILE_VEC + (uint64_t)warp * WARP_VEC;
        if (full) {
            fence_proxy_async_smem();  // generic-proxy writes -> visible to the TMA store
            __syncwarp();
            if (lane == 0) {
                tile_bulk_store(flags, reinterpret_cast<unsigned char*>(out) + warp_vec0 * 16u, s_data + warp * WARP_VEC,
                                WARP_VEC * 16u);
                bulk_commit();
                bulk_wait_read<0>();  // shared memory must outlive the engine's read
            }
        } else {
            __syncwarp();
            const uint32_t* s_k = reinterpret_cast<const uint32_t*>(s_data + warp * WARP_VEC);
            uint32_t* out1 = reinterpret_cast<uint32_t*>(out);
            const uint64_t e0 = warp_vec0 * 4u;
#pragma unroll 1
            for (uint32_t i = lane; i < WARP_VEC * 4u; i += 32u)
                if (e0 + i < n) out1[e0 + i] = s_k[i];
        }
    }
};

template <typename T, int NT, int V, bool INCLUSIVE, int WIN = 1, int REGS = 40>
__global__ void __launch_bounds__(NT, min_ctas_per_sm(NT, REGS))
    csdl_blocked_kernel(const __grid_constant__ ScanParams p)
{
    using Tile = BlockedTile<T, NT, V>;
    constexpr int NW = Tile::NW;
    __shared__ __align__(128) uint4 s_data[Tile::TILE_VEC];
    __shared__ __align__(8) uint64_t s_bar;
    __shared__ T s_warp[NW];
    __shared__ T s_prefix;
    __shared__ uint32_t s_tile;

    const uint32_t tid = threadIdx.x;
    const uint32_t lane = tid & 31u;
    const uint32_t warp = tid >> 5;
    pdl_entry();

    if (tid == 0) {
        mbar_init(&s_bar, 1);
        fence_mbar_init();
        const uint32_t t = (p.flags & PF_STATIC_TILE) ? blockIdx.x : claim_tile(p);
        s_tile = t;
        Tile::issue_load(s_data, &s_bar, p.in, t, p.n, p.flags);
    }
    __syncthreads();
    const uint32_t tile = s_tile;
    if (tile >= p.num_tiles) {  // ticket accounting out of step (an earlier launch failed): never index past the descriptors
        if (threadIdx.x == 0) atomicOr(p.status, STATUS_WATCHDOG);
        return;
    }
    const bool full = ((uint64_t)tile + 1u) * Tile::TILE_KEYS <= p.n;
    const uint32_t v0 = tid * V;  // this thread's first uint4 within the tile
    if (!Tile::arrive(s_data, &s_bar, p.in, tile, p.n, full, v0)) {  // never expected; do not hang the GPU
        atomicOr(p.status, STATUS_WATCHDOG);
        return;
    }

    // ---- pass 1: thread total -> one warp scan per tile -> aggregate first
    const T acc = Tile::thread_sum(s_data, v0);
    const T inc = warp_inclusive_scan(acc, lane);
    const T thread_excl = warp_exclusive_of(inc, acc, lane);
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    T warp_prefix = T(0), aggregate = T(0);
#pragma unroll
    for (int j = 0; j < NW; ++j) {
        const T w = s_warp[j];
        if ((uint32_t)j < warp) warp_prefix += w;
        aggregate += w;
    }

    if (warp == 0) {
        if (lane == 0 && tile != 0)
            desc_publish(p, p.desc + tile, pack_desc(to_bits(aggregate), p.epoch, ST_AGGREGATE));
        T prefix;
        if (tile == 0)
            prefix = acquire_carry<T>(p, lane);
        else
            prefix = (p.flags & PF_SKIP_LOOKBACK) ? T(0) : lookback_warp<T, WIN, GPS_LOOKBACK_MAXWIN>(p, tile, lane);
        if (lane == 0) {
            const T inclusive = prefix + aggregate;
            desc_publish(p, p.desc + tile, pack_desc(to_bits(inclusive), p.epoch, ST_INCLUSIVE));
            s_prefix = prefix;
            if (tile == p.num_tiles - 1) write_total<T>(p, inclusive);
        }
    }
    __syncthreads();

    Tile::template finish<INCLUSIVE>(s_data, v0, (s_prefix + warp_prefix) + thread_excl, p.out, tile, p.n, full, warp,
                                     lane, p.flags);
}

}  // namespace gps
