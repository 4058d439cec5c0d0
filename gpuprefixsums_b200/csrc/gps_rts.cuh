// gps_rts.cuh -- three-kernel Reduce-then-Scan (second algorithm) and device reduce.
//
// Replaces ReduceThenScan.cuh:15-161 of the reference (Reduce, SpineScan,
// Propagate{In,Ex}clusive):
//   kernel 1  rts_reduce     one CTA per tile, tile sum -> partials[tile].  Tiles are visited
//                            BACK TO FRONT, so when the pass ends the front of the array is
//                            what is left in the 126 MB L2 -- exactly what kernel 3 reads first
//   kernel 2  rts_spine      ONE block turns the tile sums into exclusive prefixes (+ carry),
//                            8192 per step, coalesced, next chunk's loads in flight while the
//                            current one is scanned (the reference's SpineScan walks a
//                            4096-wide window the same way but gets the cross-warp offset
//                            wrong beyond 128 tiles, ReduceThenScan.cuh:56-104)
//   kernel 3  rts_propagate  one CTA per tile: load, block scan, add partials[tile], store --
//                            the CSDL tile pipeline minus descriptors and look-back
// Deterministic for float32 (fixed association order).  DRAM traffic 12 B/key.
#pragma once
#include "gps_csdl_blocked.cuh"
#include "gps_tile_reg.cuh"

namespace gps {

struct RtsParams {
    const void* in;
    void* out;
    uint64_t n;
    void* partials;            // one key-sized value per tile: sums, then exclusive prefixes
    uint32_t num_tiles;        // ceil(n / TILE_KEYS)
    uint32_t carry_bits;
    const uint64_t* slots;
    uint32_t n_slots;
    uint32_t n_carry;
    uint32_t slot_tag;
    uint32_t slot_poll_limit;
    void* total_out;           // nullable
    PublishList publish;       // tagged total for downstream consumers
    uint32_t* status;
    uint32_t keep_tiles;       // reduce: the first keep_tiles tiles are loaded with an L2 evict_last policy (the
                               // scan that follows reads them first), the others evict_first; 0 = no policies
};

template <typename T, int NT>
__device__ __forceinline__ T block_reduce(T v, T* s_red, uint32_t warp, uint32_t lane)
{
    constexpr int NW = NT / 32;
    v = warp_reduce(v);
    if (lane == 0) s_red[warp] = v;
    __syncthreads();
    v = warp_reduce(lane < (uint32_t)NW ? s_red[lane] : T(0));
    __syncthreads();
    return v;  // valid in every thread
}

// ---- kernel 1: per-tile sums, back to front ---------------------------------------
template <typename T, int NT, int V>
__global__ void __launch_bounds__(NT, min_ctas_per_sm(NT, 16 + 4 * V))
    rts_reduce_kernel(const __grid_constant__ RtsParams p)
{
    pdl_entry();
    using Tile = RegTile<T, NT, V>;
    __shared__ T s_red[NT / 32];
    const uint32_t lane = lane_id();
    const uint32_t warp = threadIdx.x >> 5;
    const uint32_t tile = p.num_tiles - 1u - blockIdx.x;
    const uint64_t vbase = Tile::vec_base(tile, warp, lane);
    Tile t;
    if (((uint64_t)tile + 1u) * Tile::TILE_KEYS <= p.n) {
        if (p.keep_tiles)  // block-uniform
            t.load_full_policy(p.in, vbase, tile < p.keep_tiles ? l2_policy_evict_last() : l2_policy_evict_first());
        else
            t.load_full(p.in, vbase);
    } else {
        t.load_partial(p.in, vbase, p.n);
    }
    const T sum = block_reduce<T, NT>(t.thread_sum(), s_red, warp, lane);
    if (threadIdx.x == 0) reinterpret_cast<T*>(p.partials)[tile] = sum;
}

// (Round 2 tried to fold the total inside the reduce kernel -- every CTA adds its tile sum to an
// accumulator, fences and takes a number; the last one publishes -- to save the fold launch of
// gps_reduce_publish: the fence + returning atomic lengthen every CTA's lifetime and the reduce
// pass dropped from 7.2 to 5.5 TB/s (sharded step 0.475 -> 0.522 ms on 2 GPUs).  Reverted: the
// partials + one-block fold below stay.)
// ---- kernel 2: spine -- exclusive scan of the tile sums (+ carry) -------------------
constexpr int kSpineThreads = 1024;

template <typename T>
__device__ __forceinline__ ScanParams carry_view(const RtsParams& p)
{
    ScanParams s{};
    s.carry_bits = p.carry_bits;
    s.slots = p.slots;
    s.n_slots = p.n_slots;
    s.n_carry = p.n_carry;
    s.slot_tag = p.slot_tag;
    s.slot_poll_limit = p.slot_poll_limit;
    s.status = p.status;
    return s;
}

// One block walks the tile sums in chunks of kSpineThreads * kSpineItems with fully
// coalesced accesses (thread t owns kSpineItems consecutive values of the chunk), the loads of
// the next chunk in flight while the current one is scanned; the running total is carried
// in a register.  43,691 tile sums (2^28 keys) take 6 chunks.
constexpr int kSpineItems = 8;

template <typename T>
__global__ void __launch_bounds__(kSpineThreads) rts_spine_kernel(const __grid_constant__ RtsParams p)
{
    pdl_entry();
    constexpr int NW = kSpineThreads / 32;
    constexpr uint32_t CHUNK = kSpineThreads * kSpineItems;
    __shared__ T s_warp[2][NW];
    __shared__ T s_carry;
    const uint32_t lane = lane_id();
    const uint32_t warp = threadIdx.x >> 5;
    T* partials = reinterpret_cast<T*>(p.partials);

    if (warp == 0) {
        const ScanParams cv = carry_view<T>(p);
        const T c = acquire_carry<T>(cv, lane);
        if (lane == 0) s_carry = c;
    }
    auto load_chunk = [&](uint32_t base, T (&v)[kSpineItems]) {
        const uint32_t i0 = base + threadIdx.x * kSpineItems;
#pragma unroll
        for (int i = 0; i < kSpineItems; ++i) v[i] = (i0 + i < p.num_tiles) ? partials[i0 + i] : T(0);
    };
    T cur[kSpineItems], nxt[kSpineItems];
    load_chunk(0, cur);
    __syncthreads();
    T running = s_carry;
    uint32_t it = 0;
    for (uint32_t base = 0; base < p.num_tiles; base += CHUNK, ++it) {
        if (base + CHUNK < p.num_tiles) load_chunk(base + CHUNK, nxt);
        T tsum = T(0);
#pragma unroll
        for (int i = 0; i < kSpineItems; ++i) tsum += cur[i];
        const T inc = warp_inclusive_scan(tsum, lane);
        const T lane_excl = warp_exclusive_of(inc, tsum, lane);
        T* sw = s_warp[it & 1u];
        if (lane == 31) sw[warp] = inc;
        __syncthreads();
        T warp_prefix, total;
        block_scan_warp_totals<T, NW>(sw, warp, lane, warp_prefix, total);
        T run = (running + warp_prefix) + lane_excl;
        const uint32_t i0 = base + threadIdx.x * kSpineItems;
#pragma unroll
        for (int i = 0; i < kSpineItems; ++i) {
            if (i0 + i < p.num_tiles) partials[i0 + i] = run;
            run += cur[i];
        }
        running += total;
#pragma unroll
        for (int i = 0; i < kSpineItems; ++i) cur[i] = nxt[i];
    }
    if (threadIdx.x == 0) {
        if (p.total_out) *reinterpret_cast<T*>(p.total_out) = running;
        publish_total<T>(p.publish, running);
    }
}

// Fold of the tile sums into one total, optionally published to (peer) carry slots with
// st.relaxed.sys -- the second half of gps_reduce / gps_reduce_publish.
template <typename T>
__global__ void __launch_bounds__(kSpineThreads) rts_fold_kernel(const __grid_constant__ RtsParams p)
{
    pdl_entry();
    __shared__ T s_red[kSpineThreads / 32];
    const uint32_t lane = lane_id();
    const uint32_t warp = threadIdx.x >> 5;
    const T* partials = reinterpret_cast<const T*>(p.partials);
    T acc = T(0);
    for (uint32_t i = threadIdx.x; i < p.num_tiles; i += kSpineThreads) acc += partials[i];
    acc = block_reduce<T, kSpineThreads>(acc, s_red, warp, lane);
    if (threadIdx.x == 0) {
        if (p.total_out) *reinterpret_cast<T*>(p.total_out) = acc;
        publish_total<T>(p.publish, acc);
    }
}

// ---- kernel 3: propagate ----------------------------------------------------------
template <typename T, int NT, int V, bool INCLUSIVE>
__global__ void __launch_bounds__(NT, min_ctas_per_sm(NT, 24 + 4 * V))
    rts_propagate_kernel(const __grid_constant__ RtsParams p)
{
    pdl_entry();
    using Tile = RegTile<T, NT, V>;
    constexpr int NW = Tile::NW;
    __shared__ T s_warp[NW];
    const uint32_t lane = lane_id();
    const uint32_t warp = threadIdx.x >> 5;
    const uint32_t tile = blockIdx.x;
    const uint64_t vbase = Tile::vec_base(tile, warp, lane);
    const bool full = ((uint64_t)tile + 1u) * Tile::TILE_KEYS <= p.n;
    const T tile_prefix = reinterpret_cast<const T*>(p.partials)[tile];
    Tile t;
    if (full)
        t.load_full(p.in, vbase);
    else
        t.load_partial(p.in, vbase, p.n);
    t.scan_warp(lane);
    if (lane == 0) s_warp[warp] = t.warp_total;
    __syncthreads();
    T warp_prefix, aggregate;
    block_scan_warp_totals<T, NW>(s_warp, warp, lane, warp_prefix, aggregate);
    const T prefix = tile_prefix + warp_prefix;
    if (full)
        t.template store_full<INCLUSIVE>(p.out, vbase, prefix);
    else
        t.template store_partial<INCLUSIVE>(p.out, vbase, p.n, prefix);
}

// ---- kernel 3 on the blocked shared-memory tile (gps_csdl_blocked.cuh): TMA bulk load, one warp
// scan per tile, per-warp TMA bulk store.  Pairs with rts_reduce_kernel<T, NT, V> (same tile
// extent; the reduce may visit the keys in any order).
template <typename T, int NT, int V, bool INCLUSIVE, int REGS = 40>
__global__ void __launch_bounds__(NT, min_ctas_per_sm(NT, REGS))
    rts_propagate_blocked_kernel(const __grid_constant__ RtsParams p)
{
    pdl_entry();
    using Tile = BlockedTile<T, NT, V>;
    constexpr int NW = Tile::NW;
    __shared__ __align__(128) uint4 s_data[Tile::TILE_VEC];
    __shared__ __align__(8) uint64_t s_bar;
    __shared__ T s_warp[NW];
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const uint32_t tile = blockIdx.x;
    if (tid == 0) {
        mbar_init(&s_bar, 1);
        fence_mbar_init();
        Tile::issue_load(s_data, &s_bar, p.in, tile, p.n);
    }
    __syncthreads();
    const bool full = ((uint64_t)tile + 1u) * Tile::TILE_KEYS <= p.n;
    const uint32_t v0 = tid * V;
    const T tile_prefix = reinterpret_cast<const T*>(p.partials)[tile];
    if (!Tile::arrive(s_data, &s_bar, p.in, tile, p.n, full, v0)) {
        atomicOr(p.status, STATUS_WATCHDOG);
        return;
    }
    const T acc = Tile::thread_sum(s_data, v0);
    const T inc = warp_inclusive_scan(acc, lane);
    const T thread_excl = warp_exclusive_of(inc, acc, lane);
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    T warp_prefix = T(0);
#pragma unroll
    for (int j = 0; j < NW; ++j)
        if ((uint32_t)j < warp) warp_prefix += s_warp[j];
    Tile::template finish<INCLUSIVE>(s_data, v0, (tile_prefix + warp_prefix) + thread_excl, p.out, tile, p.n, full, warp,
                                     lane);
}

}  // namespace gps
