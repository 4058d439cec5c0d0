// gps_rts.cuh -- three-kernel Reduce-then-Scan (second algorithm) and device reduce.
//
// Replaces ReduceThenScan.cuh:15-161 of the reference (Reduce, SpineScan,
// Propagate{In,Ex}clusive).  B200-first restructuring: instead of one partial per
// 4096-key tile (65,536 partials at 2^28, scanned by a single 1024-thread block in 16
// serial iterations, ReduceThenScan.cuh:56-104), the input is cut into P contiguous
// PARTITIONS, P = a multiple of the SM count (one persistent CTA each):
//   kernel 1  rts_reduce     CTA p streams its partition, one partial per partition
//   kernel 2  rts_spine      ONE block turns the P partials into exclusive prefixes
//                            (+ carry), single pass, no loop
//   kernel 3  rts_propagate  CTA p re-streams its partition tile by tile, carrying the
//                            running prefix in a register -- no descriptors, no look-back
// Deterministic for float32 (fixed association order).  DRAM traffic 12 B/key.
#pragma once
#include "gps_tile_reg.cuh"

namespace gps {

struct RtsParams {
    const void* in;
    void* out;
    uint64_t n;
    void* partials;            // P key-sized values: sums after kernel 1, exclusive prefixes after kernel 2
    uint32_t num_tiles;        // ceil(n / TILE_KEYS)
    uint32_t tiles_per_part;   // ceil(num_tiles / P)
    uint32_t num_parts;        // P = gridDim.x of kernels 1 and 3
    uint32_t carry_bits;
    const uint64_t* slots;
    uint32_t n_slots;
    uint32_t n_carry;
    uint32_t slot_tag;
    void* total_out;           // nullable
    PublishList publish;       // tagged total for downstream consumers
    uint32_t* status;
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

// ---- kernel 1: per-partition sums ------------------------------------------------
template <typename T, int NT, int V>
__global__ void __launch_bounds__(NT) rts_reduce_kernel(const __grid_constant__ RtsParams p)
{
    using Tile = RegTile<T, NT, V>;
    __shared__ T s_red[NT / 32];
    const uint32_t lane = lane_id();
    const uint32_t warp = threadIdx.x >> 5;
    const uint32_t part = blockIdx.x;
    const uint32_t first = part * p.tiles_per_part;
    const uint32_t last = min(first + p.tiles_per_part, p.num_tiles);

    T acc = T(0);
    for (uint32_t tile = first; tile < last; ++tile) {
        const uint64_t vbase = Tile::vec_base(tile, warp, lane);
        Tile t;
        if (((uint64_t)tile + 1u) * Tile::TILE_KEYS <= p.n)
            t.load_full(p.in, vbase);
        else
            t.load_partial(p.in, vbase, p.n);
        acc += t.thread_sum();
    }
    acc = block_reduce<T, NT>(acc, s_red, warp, lane);
    if (threadIdx.x == 0) reinterpret_cast<T*>(p.partials)[part] = acc;
}

// ---- kernel 2: spine -- exclusive scan of the P partials (+ carry) ----------------
// One block of 1024 threads, up to kSpineItems partials per thread, single pass.
constexpr int kSpineThreads = 1024;
constexpr int kSpineItems = 4;
constexpr uint32_t kMaxParts = kSpineThreads * kSpineItems;

template <typename T>
__device__ __forceinline__ ScanParams carry_view(const RtsParams& p)
{
    ScanParams s{};
    s.carry_bits = p.carry_bits;
    s.slots = p.slots;
    s.n_slots = p.n_slots;
    s.n_carry = p.n_carry;
    s.slot_tag = p.slot_tag;
    s.status = p.status;
    return s;
}

template <typename T>
__global__ void __launch_bounds__(kSpineThreads) rts_spine_kernel(const __grid_constant__ RtsParams p)
{
    constexpr int NW = kSpineThreads / 32;
    __shared__ T s_warp[NW];
    __shared__ T s_carry;
    const uint32_t lane = lane_id();
    const uint32_t warp = threadIdx.x >> 5;
    T* partials = reinterpret_cast<T*>(p.partials);

    if (warp == 0) {
        const ScanParams cv = carry_view<T>(p);
        const T c = acquire_carry<T>(cv, lane);
        if (lane == 0) s_carry = c;
    }

    T v[kSpineItems];
    const uint32_t base = threadIdx.x * kSpineItems;
    T tsum = T(0);
#pragma unroll
    for (int i = 0; i < kSpineItems; ++i) {
        v[i] = (base + i < p.num_parts) ? partials[base + i] : T(0);
        tsum += v[i];
    }
    const T inc = warp_inclusive_scan(tsum, lane);
    const T lane_excl = warp_exclusive_of(inc, tsum, lane);
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    T warp_prefix, total;
    block_scan_warp_totals<T, NW>(s_warp, warp, lane, warp_prefix, total);

    T run = (s_carry + warp_prefix) + lane_excl;
#pragma unroll
    for (int i = 0; i < kSpineItems; ++i) {
        if (base + i < p.num_parts) partials[base + i] = run;
        run += v[i];
    }
    if (threadIdx.x == 0) {
        const T grand = s_carry + total;
        if (p.total_out) *reinterpret_cast<T*>(p.total_out) = grand;
        publish_total<T>(p.publish, grand);
    }
}

// Fold of the partials into one total, optionally published to (peer) carry slots with
// st.release.sys -- the second half of gps_reduce / gps_reduce_publish.
template <typename T>
__global__ void __launch_bounds__(kSpineThreads) rts_fold_kernel(const __grid_constant__ RtsParams p)
{
    __shared__ T s_red[kSpineThreads / 32];
    const uint32_t lane = lane_id();
    const uint32_t warp = threadIdx.x >> 5;
    const T* partials = reinterpret_cast<const T*>(p.partials);
    T acc = T(0);
    for (uint32_t i = threadIdx.x; i < p.num_parts; i += kSpineThreads) acc += partials[i];
    acc = block_reduce<T, kSpineThreads>(acc, s_red, warp, lane);
    if (threadIdx.x == 0) {
        if (p.total_out) *reinterpret_cast<T*>(p.total_out) = acc;
        publish_total<T>(p.publish, acc);
    }
}

// ---- kernel 3: propagate ----------------------------------------------------------
template <typename T, int NT, int V, bool INCLUSIVE>
__global__ void __launch_bounds__(NT) rts_propagate_kernel(const __grid_constant__ RtsParams p)
{
    using Tile = RegTile<T, NT, V>;
    constexpr int NW = Tile::NW;
    __shared__ T s_warp[2][NW];
    const uint32_t lane = lane_id();
    const uint32_t warp = threadIdx.x >> 5;
    const uint32_t part = blockIdx.x;
    const uint32_t first = part * p.tiles_per_part;
    const uint32_t last = min(first + p.tiles_per_part, p.num_tiles);

    T running = reinterpret_cast<const T*>(p.partials)[part];
    for (uint32_t tile = first; tile < last; ++tile) {
        const uint64_t vbase = Tile::vec_base(tile, warp, lane);
        const bool full = ((uint64_t)tile + 1u) * Tile::TILE_KEYS <= p.n;
        Tile t;
        if (full)
            t.load_full(p.in, vbase);
        else
            t.load_partial(p.in, vbase, p.n);
        t.scan_warp(lane);
        T* sw = s_warp[tile & 1u];  // double-buffered: one barrier per tile
        if (lane == 0) sw[warp] = t.warp_total;
        __syncthreads();
        T warp_prefix, aggregate;
        block_scan_warp_totals<T, NW>(sw, warp, lane, warp_prefix, aggregate);
        const T prefix = running + warp_prefix;
        if (full)
            t.template store_full<INCLUSIVE>(p.out, vbase, prefix);
        else
            t.template store_partial<INCLUSIVE>(p.out, vbase, p.n, prefix);
        running += aggregate;
    }
}

}  // namespace gps
