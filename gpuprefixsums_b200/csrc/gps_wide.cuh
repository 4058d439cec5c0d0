// gps_wide.cuh -- scans over elements wider than 32 bits: uint64, float64 and the
// four-lane struct scan (vec4<u32>, four independent scans interleaved).
//
// the reference's CUDA port is uint32 only; its WebGPU port
// ships a struct scan (GPUPrefixSumsWebGPUapis/SharedShaders/csdldf_struct.wgsl:61-316,
// validated by validate.wgsl:50-60: all-ones input -> out[i] == (i >> 2) + 1 over the flat
// u32 view).  Here the wide element types run through the Reduce-then-Scan structure of
// gps_rts.cuh (no tile descriptors, so no 128/256-bit descriptor protocol is needed and the
// result is deterministic for float64): one CTA per 24 KB tile, 128-bit coalesced
// accesses, the same lane-strided layout and branch-free PTX scan steps.
#pragma once
#include "gps_common.cuh"

namespace gps {

struct u32x4 {
    uint32_t a, b, c, d;
};
__device__ __forceinline__ u32x4 operator+(u32x4 x, u32x4 y)
{
    return u32x4{x.a + y.a, x.b + y.b, x.c + y.c, x.d + y.d};
}
__device__ __forceinline__ u32x4 operator-(u32x4 x, u32x4 y)
{
    return u32x4{x.a - y.a, x.b - y.b, x.c - y.c, x.d - y.d};
}

template <typename E>
__device__ __forceinline__ E wide_zero()
{
    if constexpr (std::is_same<E, u32x4>::value)
        return u32x4{0u, 0u, 0u, 0u};
    else
        return E(0);
}

// ---- branch-free Kogge-Stone steps for the wide types (see scan_step_up in gps_common.cuh)
__device__ __forceinline__ uint64_t scan_step_up(uint64_t v, int d)
{
    uint64_t r;
    asm volatile(
        "{\n\t.reg .u32 lo, hi, tlo, thi;\n\t.reg .u64 t;\n\t.reg .pred q;\n\t"
        "mov.b64 {lo, hi}, %1;\n\t"
        "shfl.sync.up.b32 tlo|q, lo, %2, 0, 0xffffffff;\n\t"
        "shfl.sync.up.b32 thi, hi, %2, 0, 0xffffffff;\n\t"
        "mov.b64 t, {tlo, thi};\n\t"
        "@q add.u64 t, t, %1;\n\t"
        "mov.u64 %0, t;\n\t}"
        : "=l"(r)
        : "l"(v), "r"(d));
    return r;
}
__device__ __forceinline__ double scan_step_up(double v, int d)
{
    double r;
    asm volatile(
        "{\n\t.reg .u32 lo, hi, tlo, thi;\n\t.reg .f64 t;\n\t.reg .pred q;\n\t"
        "mov.b64 {lo, hi}, %1;\n\t"
        "shfl.sync.up.b32 tlo|q, lo, %2, 0, 0xffffffff;\n\t"
        "shfl.sync.up.b32 thi, hi, %2, 0, 0xffffffff;\n\t"
        "mov.b64 t, {tlo, thi};\n\t"
        "@q add.f64 t, t, %1;\n\t"
        "mov.f64 %0, t;\n\t}"
        : "=d"(r)
        : "d"(v), "r"(d));
    return r;
}
__device__ __forceinline__ u32x4 scan_step_up(u32x4 v, int d)
{
    return u32x4{scan_step_up(v.a, d), scan_step_up(v.b, d), scan_step_up(v.c, d), scan_step_up(v.d, d)};
}

template <typename E>
__device__ __forceinline__ E wide_shfl(E v, int src)
{
    if constexpr (std::is_same<E, u32x4>::value)
        return u32x4{__shfl_sync(kFullMask, v.a, src), __shfl_sync(kFullMask, v.b, src),
                     __shfl_sync(kFullMask, v.c, src), __shfl_sync(kFullMask, v.d, src)};
    else
        return __shfl_sync(kFullMask, v, src);
}
template <typename E>
__device__ __forceinline__ E wide_shfl_up1(E v)
{
    if constexpr (std::is_same<E, u32x4>::value)
        return u32x4{__shfl_up_sync(kFullMask, v.a, 1), __shfl_up_sync(kFullMask, v.b, 1),
                     __shfl_up_sync(kFullMask, v.c, 1), __shfl_up_sync(kFullMask, v.d, 1)};
    else
        return __shfl_up_sync(kFullMask, v, 1);
}
template <typename E>
__device__ __forceinline__ E wide_warp_reduce(E v)
{
    if constexpr (std::is_same<E, u32x4>::value) {
        return u32x4{__reduce_add_sync(kFullMask, v.a), __reduce_add_sync(kFullMask, v.b),
                     __reduce_add_sync(kFullMask, v.c), __reduce_add_sync(kFullMask, v.d)};
    } else {
#pragma unroll
        for (int m = 16; m; m >>= 1) v = v + __shfl_xor_sync(kFullMask, v, m);
        return v;
    }
}
template <typename E>
__device__ __forceinline__ E wide_warp_inclusive_scan(E v)
{
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) v = scan_step_up(v, d);
    return v;
}
template <typename E>
__device__ __forceinline__ E wide_exclusive_of(E inc, E raw, uint32_t lane)
{
    if constexpr (std::is_same<E, double>::value) {
        const E t = wide_shfl_up1(inc);
        return lane ? t : E(0);
    } else {
        return inc - raw;  // exact modular arithmetic
    }
}

// ---- tile: V 128-bit vectors per thread, lane-strided; KPV elements per vector -------------
template <typename E, int NT, int V>
struct WideTile {
    static constexpr int KPV = 16 / (int)sizeof(E);
    static constexpr int NW = NT / 32;
    static constexpr uint32_t TILE_VEC = NT * V;
    static constexpr uint32_t TILE_ELEMS = TILE_VEC * KPV;
    union Vec {
        uint4 u;
        E e[KPV];
    };
    E x[V][KPV];
    E lane_excl[V];
    E warp_total;

    __device__ __forceinline__ static uint64_t vec_base(uint64_t tile, uint32_t warp, uint32_t lane)
    {
        return tile * TILE_VEC + (uint64_t)warp * (32u * V) + lane;
    }
    // elements at index >= n read as zero (n counts ELEMENTS)
    __device__ __forceinline__ void load(const void* in, uint64_t vbase, uint64_t n, bool full)
    {
        const uint4* in4 = reinterpret_cast<const uint4*>(in);
        const E* in1 = reinterpret_cast<const E*>(in);
#pragma unroll
        for (int k = 0; k < V; ++k) {
            const uint64_t v = vbase + 32u * k;
            if (full || (v + 1u) * KPV <= n) {
                Vec t;
                t.u = ldg_stream(in4 + v);
#pragma unroll
                for (int c = 0; c < KPV; ++c) x[k][c] = t.e[c];
            } else {
#pragma unroll
                for (int c = 0; c < KPV; ++c) x[k][c] = (v * KPV + c < n) ? in1[v * KPV + c] : wide_zero<E>();
            }
        }
    }
    __device__ __forceinline__ E row_sum(int k) const
    {
        E s = x[k][0];
#pragma unroll
        for (int c = 1; c < KPV; ++c) s = s + x[k][c];
        return s;
    }
    __device__ __forceinline__ E thread_sum() const
    {
        E s = row_sum(0);
#pragma unroll
        for (int k = 1; k < V; ++k) s = s + row_sum(k);
        return s;
    }
    __device__ __forceinline__ void scan_warp(uint32_t lane)
    {
        E rs[V], inc[V];
#pragma unroll
        for (int k = 0; k < V; ++k) inc[k] = rs[k] = row_sum(k);
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
#pragma unroll
            for (int k = 0; k < V; ++k) inc[k] = scan_step_up(inc[k], d);
        }
        E run = wide_zero<E>();
#pragma unroll
        for (int k = 0; k < V; ++k) {
            lane_excl[k] = run + wide_exclusive_of(inc[k], rs[k], lane);
            run = run + wide_shfl(inc[k], 31);
        }
        warp_total = run;
    }
    template <bool INCLUSIVE>
    __device__ __forceinline__ void store(void* out, uint64_t vbase, uint64_t n, bool full, E prefix) const
    {
        uint4* out4 = reinterpret_cast<uint4*>(out);
        E* out1 = reinterpret_cast<E*>(out);
#pragma unroll
        for (int k = 0; k < V; ++k) {
            Vec t;
            E run = prefix + lane_excl[k];
#pragma unroll
            for (int c = 0; c < KPV; ++c) {
                if (INCLUSIVE) {
                    run = run + x[k][c];
                    t.e[c] = run;
                } else {
                    t.e[c] = run;
                    run = run + x[k][c];
                }
            }
            const uint64_t v = vbase + 32u * k;
            if (full || (v + 1u) * KPV <= n) {
                stg_stream(out4 + v, t.u);
            } else {
#pragma unroll
                for (int c = 0; c < KPV; ++c)
                    if (v * KPV + c < n) out1[v * KPV + c] = t.e[c];
            }
        }
    }
};

template <typename E>
struct WideParams {
    const void* in;
    void* out;
    uint64_t n;          // elements
    void* partials;      // one element per tile
    uint32_t num_tiles;
    E carry;
    void* total_out;     // nullable, one element
};

template <typename E, int NW>
__device__ __forceinline__ void wide_block_scan(const E* s_warp, uint32_t warp, uint32_t lane, E& warp_prefix,
                                                E& aggregate)
{
    const E raw = lane < (uint32_t)NW ? s_warp[lane] : wide_zero<E>();
    const E inc = wide_warp_inclusive_scan(raw);
    const E exc = wide_exclusive_of(inc, raw, lane);
    warp_prefix = wide_shfl(exc, (int)warp);
    aggregate = wide_shfl(inc, NW - 1);
}

// kernel 1: tile sums, back to front (L2 reuse by kernel 3, see gps_rts.cuh)
template <typename E, int NT, int V>
__global__ void __launch_bounds__(NT) wide_reduce_kernel(const __grid_constant__ WideParams<E> p)
{
    using Tile = WideTile<E, NT, V>;
    __shared__ E s_red[NT / 32];
    const uint32_t lane = lane_id(), warp = threadIdx.x >> 5;
    const uint32_t tile = p.num_tiles - 1u - blockIdx.x;
    const bool full = ((uint64_t)tile + 1u) * Tile::TILE_ELEMS <= p.n;
    Tile t;
    t.load(p.in, Tile::vec_base(tile, warp, lane), p.n, full);
    const E ws = wide_warp_reduce(t.thread_sum());
    if (lane == 0) s_red[warp] = ws;
    __syncthreads();
    if (warp == 0) {
        const E tot = wide_warp_reduce(lane < (uint32_t)(NT / 32) ? s_red[lane] : wide_zero<E>());
        if (lane == 0) reinterpret_cast<E*>(p.partials)[tile] = tot;
    }
}

// kernel 2: one-block spine over the tile sums, 4 per thread per coalesced step
template <typename E>
__global__ void __launch_bounds__(1024) wide_spine_kernel(const __grid_constant__ WideParams<E> p)
{
    constexpr int NW = 32, ITEMS = 4;
    constexpr uint32_t CHUNK = 1024 * ITEMS;
    __shared__ E s_warp[2][NW];
    const uint32_t lane = lane_id(), warp = threadIdx.x >> 5;
    E* partials = reinterpret_cast<E*>(p.partials);
    E running = p.carry;
    uint32_t it = 0;
    for (uint32_t base = 0; base < p.num_tiles; base += CHUNK, ++it) {
        const uint32_t i0 = base + threadIdx.x * ITEMS;
        E v[ITEMS];
        E tsum = wide_zero<E>();
#pragma unroll
        for (int i = 0; i < ITEMS; ++i) {
            v[i] = (i0 + i < p.num_tiles) ? partials[i0 + i] : wide_zero<E>();
            tsum = tsum + v[i];
        }
        const E inc = wide_warp_inclusive_scan(tsum);
        const E lane_excl = wide_exclusive_of(inc, tsum, lane);
        E* sw = s_warp[it & 1u];
        if (lane == 31) sw[warp] = inc;
        __syncthreads();
        E warp_prefix, total;
        wide_block_scan<E, NW>(sw, warp, lane, warp_prefix, total);
        E run = (running + warp_prefix) + lane_excl;
#pragma unroll
        for (int i = 0; i < ITEMS; ++i) {
            if (i0 + i < p.num_tiles) partials[i0 + i] = run;
            run = run + v[i];
        }
        running = running + total;
    }
    if (threadIdx.x == 0 && p.total_out) *reinterpret_cast<E*>(p.total_out) = running;
}

// kernel 3: propagate
#ifndef GPS_WIDE_REGS
#define GPS_WIDE_REGS 64
#endif
template <typename E, int NT, int V, bool INCLUSIVE>
__global__ void __launch_bounds__(NT, min_ctas_per_sm(NT, GPS_WIDE_REGS)) wide_propagate_kernel(const __grid_constant__ WideParams<E> p)
{
    using Tile = WideTile<E, NT, V>;
    __shared__ E s_warp[Tile::NW];
    const uint32_t lane = lane_id(), warp = threadIdx.x >> 5;
    const uint32_t tile = blockIdx.x;
    const bool full = ((uint64_t)tile + 1u) * Tile::TILE_ELEMS <= p.n;
    const uint64_t vbase = Tile::vec_base(tile, warp, lane);
    const E tile_prefix = reinterpret_cast<const E*>(p.partials)[tile];
    Tile t;
    t.load(p.in, vbase, p.n, full);
    t.scan_warp(lane);
    if (lane == 0) s_warp[warp] = t.warp_total;
    __syncthreads();
    E warp_prefix, aggregate;
    wide_block_scan<E, Tile::NW>(s_warp, warp, lane, warp_prefix, aggregate);
    t.template store<INCLUSIVE>(p.out, vbase, p.n, full, tile_prefix + warp_prefix);
}

}  // namespace gps
