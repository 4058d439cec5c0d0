// gps_wide.cuh -- scans over elements wider than 32 bits: uint64, float64 and the
// four-lane struct scan (vec4<u32>, four independent scans interleaved).
//
// the reference's CUDA port is uint32 only; its WebGPU port
// ships a struct scan (GPUPrefixSumsWebGPUapis/SharedShaders/csdldf_struct.wgsl:61-316,
// validated by validate.wgsl:50-60: all-ones input -> out[i] == (i >> 2) + 1 over the flat
// u32 view).
//
// Why this is not the uint32 kernel with a wider type: a lane-strided tile needs one warp
// scan per 512-byte row, i.e. 5 shuffles per 32-bit word of the element per row.  At the
// HBM rate an SM has ~22 clocks per row and issues one shuffle per clock: uint32 spends 7,
// uint64 ~14, vec4<u32> ~26 -- the struct scan would be SHUFFLE bound (the first version of
// this file, lane-strided Reduce-then-Scan in registers, ran at 2.3-3.4 TB/s of useful
// bytes).  So the wide kernels use a BLOCKED tile instead:
//   * one TMA bulk load (cp.async.bulk) brings the whole tile into shared memory, as in
//     gps_csdl_smem.cuh;
//   * thread t owns the V consecutive 16-byte vectors [t*V, (t+1)*V) of the tile; V is ODD,
//     so the LDS.128/STS.128 of a quarter-warp fall into 8 distinct 16-byte bank groups
//     (stride V*16 B, V coprime to 8): bank-conflict free without padding or swizzle;
//   * the thread scans its run serially; ONE warp scan of the thread totals per tile (5 steps
//     x W shuffles, W = 32-bit words per element) replaces the V per-row warp scans;
//   * results are written back into the same shared memory and leave with one TMA bulk
//     store per warp (cp.async.bulk.global.shared::cta), so the global writes are full
//     lines although each thread produced a contiguous run.
// Chained scan (default): tile descriptors are 2*W 64-bit words {value32 | epoch32} per
// tile -- W words for the AGGREGATE and W separate words for the INCLUSIVE prefix.  Nothing
// is ever overwritten within a launch, every word validates itself by its epoch, and a set
// counts as published when all its W words carry the launch's epoch: no 128/256-bit atomic
// is needed and a torn update cannot be mistaken for a value.  Reduce-then-Scan
// (GPS_ALGO_RTS) runs the same tile kernel with the prefix taken from the spine; it is
// deterministic for float64.
#pragma once
#include "gps_common.cuh"
#include "gps_csdl_reg.cuh"  // LB_DONE / LB_FALLBACK

namespace gps {

struct __align__(16) u32x4 {  // 16-byte aligned: one LDG.128 / STG.128 per element
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

// ---- 32-bit words of an element (descriptor payload) --------------------------------------
template <typename E>
__device__ __forceinline__ uint32_t wide_word(const E& v, uint32_t w)
{
    if constexpr (std::is_same<E, u32x4>::value) {
        return w == 0 ? v.a : w == 1 ? v.b : w == 2 ? v.c : v.d;
    } else if constexpr (std::is_same<E, uint64_t>::value) {
        return w ? (uint32_t)(v >> 32) : (uint32_t)v;
    } else if constexpr (std::is_same<E, double>::value) {
        const uint64_t b = (uint64_t)__double_as_longlong(v);
        return w ? (uint32_t)(b >> 32) : (uint32_t)b;
    } else {
        return to_bits(v);
    }
}
template <typename E>
__device__ __forceinline__ E wide_from_words(const uint32_t* w)
{
    if constexpr (std::is_same<E, u32x4>::value) {
        return u32x4{w[0], w[1], w[2], w[3]};
    } else if constexpr (std::is_same<E, uint64_t>::value) {
        return ((uint64_t)w[1] << 32) | w[0];
    } else if constexpr (std::is_same<E, double>::value) {
        return __longlong_as_double((long long)(((uint64_t)w[1] << 32) | w[0]));
    } else {
        return from_bits<E>(w[0]);
    }
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
    } else if constexpr (std::is_same<E, uint32_t>::value) {
        return __reduce_add_sync(kFullMask, v);
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
    if constexpr (std::is_floating_point<E>::value) {
        const E t = wide_shfl_up1(inc);
        return lane ? t : E(0);
    } else {
        return inc - raw;  // exact modular arithmetic
    }
}

// ---- parameters ----------------------------------------------------------------------------
template <typename E>
struct WideParams {
    const void* in;
    void* out;
    uint64_t n;             // ELEMENTS
    uint64_t* desc;         // chained: 2*W words per tile {value32 | epoch32}: W aggregate, W inclusive
    uint32_t* counter;      // tile-claim counter shared with the 32-bit kernels, never reset
    uint32_t counter_base;
    uint32_t epoch;         // != 0, unique per launch until the host clears desc
    uint32_t num_tiles;
    void* partials;         // RTS: one element per tile
    uint32_t* status;       // watchdog word
    E carry;
    void* total_out;        // nullable, one element: carry + sum of all elements
};

template <typename E, int NT, int V>
struct WideShape {
    static_assert(V % 2 == 1, "V must be odd: conflict-free LDS.128 at stride V*16 bytes");
    static constexpr int W = (int)sizeof(E) / 4;     // 32-bit words per element
    static constexpr int KPV = 16 / (int)sizeof(E);  // elements per 16-byte vector
    static constexpr int NW = NT / 32;
    static constexpr uint32_t TILE_VEC = NT * V;
    static constexpr uint32_t TILE_ELEMS = TILE_VEC * KPV;
    static constexpr uint32_t TILE_BYTES = TILE_VEC * 16;
    static constexpr uint32_t WARP_VEC = 32 * V;
    union Vec {
        uint4 u;
        E e[KPV];
    };
};

__device__ __forceinline__ void wide_desc_publish(uint64_t* ptr, uint32_t value, uint32_t epoch)
{
    uint64_t v = ((uint64_t)epoch << 32) | value;
    asm volatile("atom.relaxed.gpu.global.exch.b64 %0, [%1], %2;" : "=l"(v) : "l"(ptr), "l"(v) : "memory");
}

// Warp-parallel decoupled look-back over the two-set descriptors: lane l examines tile
// (top - l); a tile is INCLUSIVE when its W inclusive words carry this launch's epoch,
// AGGREGATE when its W aggregate words do, NOT_READY otherwise.  Same window policy as
// lookback_warp (gps_common.cuh): consume the ready run nearest to the tile, stop at the
// nearest INCLUSIVE in it.
template <typename E>
__device__ __forceinline__ E wide_lookback(const WideParams<E>& p, uint32_t tile, uint32_t lane)
{
    constexpr int W = (int)sizeof(E) / 4;
    E prefix = wide_zero<E>();
    int64_t top = (int64_t)tile - 1;
    uint32_t polls = 0;
    while (true) {
        const int64_t idx = top - (int64_t)lane;
        uint32_t state = ST_INCLUSIVE;  // virtual entry below tile 0: INCLUSIVE / 0
        E v = wide_zero<E>();
        if (idx >= 0) {
            const uint64_t* d = p.desc + (uint64_t)idx * (2 * W);
            uint64_t w[2 * W];
            if constexpr (W == 1) {
                w[0] = ld_relaxed_gpu(d);
                w[1] = ld_relaxed_gpu(d + 1);
            } else {
#pragma unroll
                for (int i = 0; i < 2 * W; i += 2)
                    asm volatile("ld.relaxed.gpu.global.v2.u64 {%0,%1}, [%2];"
                                 : "=l"(w[i]), "=l"(w[i + 1])
                                 : "l"(d + i)
                                 : "memory");
            }
            bool agg_ok = true, inc_ok = true;
#pragma unroll
            for (int i = 0; i < W; ++i) {
                agg_ok = agg_ok && (uint32_t)(w[i] >> 32) == p.epoch;
                inc_ok = inc_ok && (uint32_t)(w[W + i] >> 32) == p.epoch;
            }
            state = inc_ok ? ST_INCLUSIVE : agg_ok ? ST_AGGREGATE : ST_NOT_READY;
            uint32_t words[W];
#pragma unroll
            for (int i = 0; i < W; ++i) words[i] = (uint32_t)(inc_ok ? w[W + i] : w[i]);
            v = wide_from_words<E>(words);
        }
        const uint32_t not_ready = __ballot_sync(kFullMask, state == ST_NOT_READY);
        const uint32_t run = not_ready ? (uint32_t)__ffs((int)not_ready) - 1u : 32u;
        const uint32_t in_run = run >= 32u ? kFullMask : ((1u << run) - 1u);
        const uint32_t incl = __ballot_sync(kFullMask, state == ST_INCLUSIVE) & in_run;
        if (incl) {
            const uint32_t nearest = (uint32_t)__ffs((int)incl) - 1u;
            prefix = prefix + wide_warp_reduce(lane <= nearest ? v : wide_zero<E>());
            break;
        }
        if (run == 0u) {
            const bool bail = watchdog_tick(polls, p.status);
            if (__any_sync(kFullMask, bail)) break;  // warp-uniform exit
            continue;
        }
        prefix = prefix + wide_warp_reduce(lane < run ? v : wide_zero<E>());
        top -= (int64_t)run;
    }
    return prefix;
}

// ---- the tile kernel: chained scan (CHAINED) or RTS propagate ------------------------------
template <typename E, int NT, int V, bool INCLUSIVE, bool CHAINED, int REGS = 40>
__global__ void __launch_bounds__(NT, min_ctas_per_sm(NT, REGS))
    wide_scan_kernel(const __grid_constant__ WideParams<E> p)
{
    pdl_entry();
    using S = WideShape<E, NT, V>;
    using Vec = typename S::Vec;
    constexpr int W = S::W, KPV = S::KPV, NW = S::NW;
    __shared__ __align__(128) uint4 s_data[S::TILE_VEC];
    __shared__ __align__(8) uint64_t s_bar;
    __shared__ __align__(16) E s_warp[NW];
    __shared__ __align__(16) E s_prefix;
    __shared__ uint32_t s_tile;

    const uint32_t tid = threadIdx.x;
    const uint32_t lane = tid & 31u;
    const uint32_t warp = tid >> 5;

    if (tid == 0) {
        mbar_init(&s_bar, 1);
        fence_mbar_init();
        const uint32_t t = CHAINED ? atomicAdd(p.counter, 1u) - p.counter_base : blockIdx.x;
        s_tile = t;
        if (((uint64_t)t + 1u) * S::TILE_ELEMS <= p.n) {
            mbar_arrive_expect_tx(&s_bar, S::TILE_BYTES);
            bulk_load(s_data, reinterpret_cast<const unsigned char*>(p.in) + (uint64_t)t * S::TILE_BYTES,
                      S::TILE_BYTES, &s_bar);
        }
    }
    __syncthreads();
    const uint32_t tile = s_tile;
    if (tile >= p.num_tiles) {  // ticket accounting out of step (an earlier launch failed)
        if (tid == 0) atomicOr(p.status, STATUS_WATCHDOG);
        return;
    }
    const bool full = ((uint64_t)tile + 1u) * S::TILE_ELEMS <= p.n;
    const uint32_t v0 = tid * V;  // this thread's first vector within the tile

    if (full) {
        if (!mbar_wait_bounded(&s_bar, 0u)) {  // never expected; do not hang the GPU
            atomicOr(p.status, STATUS_WATCHDOG);
            return;
        }
    } else {
        // partial (last) tile: guarded loads through the LSU, elements at index >= n read as 0;
        // each thread only re-reads what it wrote itself until the warp-level copy-out
        const E* in1 = reinterpret_cast<const E*>(p.in);
#pragma unroll 1
        for (int k = 0; k < V; ++k) {
            const uint64_t e0 = ((uint64_t)tile * S::TILE_VEC + v0 + k) * KPV;
            Vec t;
#pragma unroll
            for (int c = 0; c < KPV; ++c) t.e[c] = e0 + c < p.n ? in1[e0 + c] : wide_zero<E>();
            s_data[v0 + k] = t.u;
        }
    }

    // ---- pass 1: thread total -> one warp scan per tile -> aggregate first
    E acc = wide_zero<E>();
#pragma unroll
    for (int k = 0; k < V; ++k) {
        Vec t;
        t.u = s_data[v0 + k];
#pragma unroll
        for (int c = 0; c < KPV; ++c) acc = acc + t.e[c];
    }
    const E inc = wide_warp_inclusive_scan(acc);
    const E thread_excl = wide_exclusive_of(inc, acc, lane);
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    E warp_prefix = wide_zero<E>(), aggregate = wide_zero<E>();
#pragma unroll
    for (int j = 0; j < NW; ++j) {
        const E w = s_warp[j];
        if ((uint32_t)j < warp) warp_prefix = warp_prefix + w;
        aggregate = aggregate + w;
    }

    E tile_prefix;
    if constexpr (CHAINED) {
        if (warp == 0) {
            uint64_t* d = p.desc + (uint64_t)tile * (2 * W);
            if (tile != 0 && lane < (uint32_t)W) wide_desc_publish(d + lane, wide_word(aggregate, lane), p.epoch);
            const E prefix = tile == 0 ? p.carry : wide_lookback<E>(p, tile, lane);
            const E inclusive = prefix + aggregate;
            if (lane < (uint32_t)W) wide_desc_publish(d + W + lane, wide_word(inclusive, lane), p.epoch);
            if (lane == 0) {
                s_prefix = prefix;
                if (tile == p.num_tiles - 1 && p.total_out) *reinterpret_cast<E*>(p.total_out) = inclusive;
            }
        }
        __syncthreads();
        tile_prefix = s_prefix;
    } else {
        tile_prefix = reinterpret_cast<const E*>(p.partials)[tile];
    }

    // ---- pass 2: serial scan of the thread's run, in place in shared memory
    E run = (tile_prefix + warp_prefix) + thread_excl;
#pragma unroll
    for (int k = 0; k < V; ++k) {
        Vec t;
        t.u = s_data[v0 + k];
#pragma unroll
        for (int c = 0; c < KPV; ++c) {
            const E x = t.e[c];
            if (INCLUSIVE) {
                run = run + x;
                t.e[c] = run;
            } else {
                t.e[c] = run;
                run = run + x;
            }
        }
        s_data[v0 + k] = t.u;
    }

    // ---- the warp's 32*V vectors are contiguous in the tile and in memory
    const uint64_t warp_vec0 = (uint64_t)tile * S::TILE_VEC + (uint64_t)warp * S::WARP_VEC;
    if (full) {
        fence_proxy_async_smem();  // generic-proxy writes -> visible to the TMA store
        __syncwarp();
        if (lane == 0) {
            bulk_store(reinterpret_cast<unsigned char*>(p.out) + warp_vec0 * 16u, s_data + warp * S::WARP_VEC,
                       S::WARP_VEC * 16u);
            bulk_commit();
            bulk_wait_read<0>();  // shared memory must outlive the engine's read
        }
    } else {
        __syncwarp();
        const E* s_e = reinterpret_cast<const E*>(s_data + warp * S::WARP_VEC);
        E* out1 = reinterpret_cast<E*>(p.out);
        const uint64_t e0 = warp_vec0 * KPV;
#pragma unroll 1
        for (uint32_t i = lane; i < S::WARP_VEC * KPV; i += 32u)
            if (e0 + i < p.n) out1[e0 + i] = s_e[i];
    }
}

// ---- the chained tile kernel with Decoupled Fallback (and optional emulated stalls) ----------
// The struct scan of the reference IS a fallback kernel (GPUPrefixSumsWebGPUapis/SharedShaders/
// csdldf_struct.wgsl:61-316: per-member flag words, a spin budget, the whole workgroup re-reducing
// the blocking tile from the input, insertion with an atomic that never overwrites a finished
// word).  Here: warp 0 walks the two-set descriptors with a spin budget; when a window stays
// blocked, the whole CTA re-reduces the nearest not-ready tile from the INPUT -- in exactly the
// association order its owner uses (thread t sums its V consecutive vectors serially, one warp
// scan, warp totals added in order), so the recomputed aggregate is BIT-IDENTICAL to the owner's
// for every element type, float64 included, and words written by the owner and by rescuers can
// mix freely -- and inserts the W aggregate words with atomicCAS from the stale word it observes.
// In place: the owner's descriptor words are performed (fence) before the barrier that precedes
// its first output store, and a rescuer fences between its re-read and its CAS; a rescuer that
// read half-overwritten input therefore finds every word already valid and takes the owner's.
struct WideFallbackParams {
    uint32_t stall_mask;  // tiles with (tile & mask) == 0 stall; 0xFFFFFFFF = none
    uint32_t max_spin;    // unsuccessful polls of a window before falling back
    uint32_t* stats;      // nullable: {spins, fallbacks, insertions, lookback windows}
};

template <typename E>
__device__ __forceinline__ bool wide_word_cas_insert(uint64_t* ptr, uint32_t value, uint32_t epoch, uint32_t& result)
{
    // stale word -> {value | epoch}; a word that already carries this launch's epoch is kept
    const uint64_t mine = ((uint64_t)epoch << 32) | value;
    uint64_t seen = ld_relaxed_gpu(ptr);
    while ((uint32_t)(seen >> 32) != epoch) {
        const uint64_t old = atomicCAS(reinterpret_cast<unsigned long long*>(ptr), (unsigned long long)seen,
                                       (unsigned long long)mine);
        if (old == seen) {
            result = value;
            return true;
        }
        seen = old;
    }
    result = (uint32_t)seen;
    return false;
}

template <typename E, int NT, int V, bool INCLUSIVE, int REGS = 48>
__global__ void __launch_bounds__(NT, min_ctas_per_sm(NT, REGS))
    wide_scan_fallback_kernel(const __grid_constant__ WideParams<E> p, const WideFallbackParams f)
{
    pdl_entry();
    using S = WideShape<E, NT, V>;
    using Vec = typename S::Vec;
    constexpr int W = S::W, KPV = S::KPV, NW = S::NW;
    __shared__ __align__(128) uint4 s_data[S::TILE_VEC];
    __shared__ __align__(8) uint64_t s_bar;
    __shared__ __align__(16) E s_warp[NW];
    __shared__ __align__(16) E s_fb[NW];
    __shared__ __align__(16) E s_prefix;
    __shared__ uint32_t s_tile, s_state, s_fb_tile;
    __shared__ uint32_t s_stats[4];  // {spins, fallbacks, insertions, windows}: only touched when f.stats is set

    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const bool in_place = p.in == p.out;  // only then may a rescuer's re-read race with the owner's stores
    const bool stats = f.stats != nullptr;
    if (tid < 4u) s_stats[tid] = 0u;
    if (tid == 0) {
        mbar_init(&s_bar, 1);
        fence_mbar_init();
        const uint32_t t = atomicAdd(p.counter, 1u) - p.counter_base;
        s_tile = t;
        if (((uint64_t)t + 1u) * S::TILE_ELEMS <= p.n && t < p.num_tiles) {
            mbar_arrive_expect_tx(&s_bar, S::TILE_BYTES);
            bulk_load(s_data, reinterpret_cast<const unsigned char*>(p.in) + (uint64_t)t * S::TILE_BYTES,
                      S::TILE_BYTES, &s_bar);
        }
    }
    __syncthreads();
    const uint32_t tile = s_tile;
    if (tile >= p.num_tiles) {  // ticket accounting out of step (an earlier launch failed)
        if (tid == 0) atomicOr(p.status, STATUS_WATCHDOG);
        return;
    }
    const bool full = ((uint64_t)tile + 1u) * S::TILE_ELEMS <= p.n;
    const uint32_t v0 = tid * V;
    if (full) {
        if (!mbar_wait_bounded(&s_bar, 0u)) {
            atomicOr(p.status, STATUS_WATCHDOG);
            return;
        }
    } else {
        const E* in1 = reinterpret_cast<const E*>(p.in);
#pragma unroll 1
        for (int k = 0; k < V; ++k) {
            const uint64_t e0 = ((uint64_t)tile * S::TILE_VEC + v0 + k) * KPV;
            Vec t;
#pragma unroll
            for (int c = 0; c < KPV; ++c) t.e[c] = e0 + c < p.n ? in1[e0 + c] : wide_zero<E>();
            s_data[v0 + k] = t.u;
        }
    }

    // ---- pass 1 (the association order the fallback below reproduces)
    E acc = wide_zero<E>();
#pragma unroll
    for (int k = 0; k < V; ++k) {
        Vec t;
        t.u = s_data[v0 + k];
#pragma unroll
        for (int c = 0; c < KPV; ++c) acc = acc + t.e[c];
    }
    const E inc = wide_warp_inclusive_scan(acc);
    E local_prefix = wide_exclusive_of(inc, acc, lane);  // within the warp ...
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    E aggregate = wide_zero<E>();
    {
        E warp_prefix = wide_zero<E>();
#pragma unroll
        for (int j = 0; j < NW; ++j) {
            const E w = s_warp[j];
            if ((uint32_t)j < warp) warp_prefix = warp_prefix + w;
            aggregate = aggregate + w;
        }
        local_prefix = warp_prefix + local_prefix;  // ... and within the tile: one value kept across the look-back
    }

    // stall emulation (EmulatedDeadlocking.cuh:63-72 / the WebGPU test variants): the tile refuses
    // to publish until a successor's fallback has published for it; the last tile never stalls
    const bool stalled = f.stall_mask != 0xFFFFFFFFu && (tile & f.stall_mask) == 0u && p.num_tiles > 1u &&
                         tile != p.num_tiles - 1u;
    uint64_t* const my_desc = p.desc + (uint64_t)tile * (2 * W);
    if (warp == 0) {
        if (stalled && lane < (uint32_t)W) {
            uint32_t polls = 0;
            uint64_t* word = my_desc + (tile ? 0 : W) + lane;  // tile 0 is rescued with INCLUSIVE words
            while ((uint32_t)(ld_relaxed_gpu(word) >> 32) != p.epoch)
                if (watchdog_tick(polls, p.status)) break;
        }
        __syncwarp();
        if (tile != 0 && lane < (uint32_t)W) {
            wide_desc_publish(my_desc + lane, wide_word(aggregate, lane), p.epoch);
            if (in_place) __threadfence();  // performed before any thread of this CTA overwrites the input
        }
    }

    // One barrier per round: warp 0 walks (with a spin budget); when the walk is complete it publishes
    // INCLUSIVE at once -- that word is what every successor waits for -- and only then meets the
    // other warps at the barrier; a round that ends in LB_FALLBACK sends the whole CTA to re-reduce.
    E prefix = tile ? wide_zero<E>() : p.carry;
    bool finished = tile == 0;  // tile 0: nothing to look back at
    int64_t top = (int64_t)tile - 1;
    while (true) {
        if (warp == 0) {
            uint32_t state = LB_DONE, spins = 0, polls = 0, fb = 0;
            while (!finished) {
                const int64_t idx = top - (int64_t)lane;
                uint32_t st = ST_INCLUSIVE;  // virtual entry below tile 0
                E v = wide_zero<E>();
                if (idx >= 0) {
                    const uint64_t* d = p.desc + (uint64_t)idx * (2 * W);
                    uint64_t w[2 * W];  // relaxed: every word carries its own epoch tag next to its payload
                    if constexpr (W == 1) {
                        w[0] = ld_relaxed_gpu(d);
                        w[1] = ld_relaxed_gpu(d + 1);
                    } else {
#pragma unroll
                        for (int i = 0; i < 2 * W; i += 2)
                            asm volatile("ld.relaxed.gpu.global.v2.u64 {%0,%1}, [%2];"
                                         : "=l"(w[i]), "=l"(w[i + 1])
                                         : "l"(d + i)
                                         : "memory");
                    }
                    bool agg_ok = true, inc_ok = true;
#pragma unroll
                    for (int i = 0; i < W; ++i) {
                        agg_ok = agg_ok && (uint32_t)(w[i] >> 32) == p.epoch;
                        inc_ok = inc_ok && (uint32_t)(w[W + i] >> 32) == p.epoch;
                    }
                    st = inc_ok ? ST_INCLUSIVE : agg_ok ? ST_AGGREGATE : ST_NOT_READY;
                    uint32_t words[W];
#pragma unroll
                    for (int i = 0; i < W; ++i) words[i] = (uint32_t)(inc_ok ? w[W + i] : w[i]);
                    v = wide_from_words<E>(words);
                }
                if (stats && lane == 0) atomicAdd(&s_stats[3], 1u);
                const uint32_t not_ready = __ballot_sync(kFullMask, st == ST_NOT_READY);
                const uint32_t first_nr = not_ready ? (uint32_t)__ffs((int)not_ready) - 1u : 32u;
                const uint32_t below = first_nr >= 32u ? kFullMask : ((1u << first_nr) - 1u);
                const uint32_t incl = __ballot_sync(kFullMask, st == ST_INCLUSIVE) & below;
                if (incl) {
                    const uint32_t nearest = (uint32_t)__ffs((int)incl) - 1u;
                    prefix = prefix + wide_warp_reduce(lane <= nearest ? v : wide_zero<E>());
                    break;
                }
                if (first_nr == 32u) {  // whole window ready, none inclusive
                    prefix = prefix + wide_warp_reduce(v);
                    top -= 32;
                    spins = 0;
                    continue;
                }
                if (stats && lane == 0) atomicAdd(&s_stats[0], 1u);
                const bool bail = watchdog_tick(polls, p.status);
                if (__any_sync(kFullMask, bail)) break;
                if (++spins < f.max_spin) continue;
                // budget exhausted: consume the ready tiles nearer than the blocker, fall back on it
                prefix = prefix + wide_warp_reduce(lane < first_nr ? v : wide_zero<E>());
                fb = (uint32_t)(top - (int64_t)first_nr);
                top = (int64_t)fb - 1;
                state = LB_FALLBACK;
                break;
            }
            if (state == LB_DONE) {  // warp-uniform
                const E inclusive = prefix + aggregate;
                if (lane < (uint32_t)W) {
                    wide_desc_publish(my_desc + W + lane, wide_word(inclusive, lane), p.epoch);
                    if (in_place) __threadfence();
                }
                if (lane == 0) {
                    s_prefix = prefix;
                    if (tile == p.num_tiles - 1 && p.total_out) *reinterpret_cast<E*>(p.total_out) = inclusive;
                }
            }
            if (lane == 0) {
                s_state = state;
                s_fb_tile = fb;
            }
        }
        __syncthreads();
        if (s_state == LB_DONE) break;

        // ---- Decoupled Fallback: every thread re-reads ITS blocked run of tile fb (a fallback
        // target is never the last tile => always full) and the CTA repeats pass 1 on it
        const uint32_t fb = s_fb_tile;
        if (stats && tid == 0) atomicAdd(&s_stats[1], 1u);  // (atomic: a plain ++ lets the compiler hoist the load above the predicate, which racecheck flags)
        {
            const uint4* in4 = reinterpret_cast<const uint4*>(p.in) + (uint64_t)fb * S::TILE_VEC + v0;
            E facc = wide_zero<E>();
#pragma unroll
            for (int k = 0; k < V; ++k) {
                Vec t;
                t.u = in4[k];  // L1-allocating: the 32-byte sectors of this strided pattern are reused by the next k
#pragma unroll
                for (int c = 0; c < KPV; ++c) facc = facc + t.e[c];
            }
            const E finc = wide_warp_inclusive_scan(facc);
            if (lane == 31) s_fb[warp] = finc;
        }
        __syncthreads();  // also: every warp has read s_state / s_fb_tile before warp 0 rewrites them
        if (warp == 0) {
            E red = wide_zero<E>();
#pragma unroll
            for (int j = 0; j < NW; ++j) red = red + s_fb[j];
            if (fb == 0) red = p.carry + red;  // tile 0 is INCLUSIVE incl. the carry (owner: prefix + aggregate)
            if (in_place) __threadfence();  // my re-read is ordered before my CAS (see the header comment)
            uint32_t got = 0;
            bool inserted = false;
            if (lane < (uint32_t)W)
                inserted = wide_word_cas_insert<E>(p.desc + (uint64_t)fb * (2 * W) + (fb ? 0 : W) + lane,
                                                   wide_word(red, lane), p.epoch, got);
            if (stats && __any_sync(kFullMask, inserted) && lane == 0) atomicAdd(&s_stats[2], 1u);
            // the value in the descriptor (mine or the owner's: identical unless my read raced
            // with an in-place overwrite, in which case every word is the owner's)
            uint32_t words[W];
#pragma unroll
            for (int i = 0; i < W; ++i) words[i] = __shfl_sync(kFullMask, got, i);
            prefix = prefix + wide_from_words<E>(words);
            finished = fb == 0;  // the walk goes on below fb, or ends here
        }
    }
    if (stats) {
        __syncthreads();
        if (tid < 4u && s_stats[tid]) atomicAdd(f.stats + tid, s_stats[tid]);
    }

    // ---- pass 2 + store: as wide_scan_kernel
    E run = s_prefix + local_prefix;
#pragma unroll
    for (int k = 0; k < V; ++k) {
        Vec t;
        t.u = s_data[v0 + k];
#pragma unroll
        for (int c = 0; c < KPV; ++c) {
            const E x = t.e[c];
            if (INCLUSIVE) {
                run = run + x;
                t.e[c] = run;
            } else {
                t.e[c] = run;
                run = run + x;
            }
        }
        s_data[v0 + k] = t.u;
    }
    const uint64_t warp_vec0 = (uint64_t)tile * S::TILE_VEC + (uint64_t)warp * S::WARP_VEC;
    if (full) {
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
            bulk_store(reinterpret_cast<unsigned char*>(p.out) + warp_vec0 * 16u, s_data + warp * S::WARP_VEC,
                       S::WARP_VEC * 16u);
            bulk_commit();
            bulk_wait_read<0>();
        }
    } else {
        __syncwarp();
        const E* s_e = reinterpret_cast<const E*>(s_data + warp * S::WARP_VEC);
        E* out1 = reinterpret_cast<E*>(p.out);
        const uint64_t e0 = warp_vec0 * KPV;
#pragma unroll 1
        for (uint32_t i = lane; i < S::WARP_VEC * KPV; i += 32u)
            if (e0 + i < p.n) out1[e0 + i] = s_e[i];
    }
}

// ---- RTS kernel 1: tile sums (tiles of the same extent as wide_scan_kernel's), visited
// back to front so that the front of the array is what L2 still holds for kernel 3
template <typename E, int NT, int V>
__global__ void __launch_bounds__(NT) wide_reduce_kernel(const __grid_constant__ WideParams<E> p)
{
    pdl_entry();
    using S = WideShape<E, NT, V>;
    using Vec = typename S::Vec;
    constexpr int KPV = S::KPV;
    __shared__ __align__(16) E s_red[S::NW];
    const uint32_t lane = lane_id(), warp = threadIdx.x >> 5;
    const uint32_t tile = p.num_tiles - 1u - blockIdx.x;
    const bool full = ((uint64_t)tile + 1u) * S::TILE_ELEMS <= p.n;
    const uint4* in4 = reinterpret_cast<const uint4*>(p.in);
    const E* in1 = reinterpret_cast<const E*>(p.in);
    const uint64_t vbase = (uint64_t)tile * S::TILE_VEC + (uint64_t)warp * S::WARP_VEC + lane;  // lane-strided
    E acc = wide_zero<E>();
    if (full) {  // hoisted: V independent 128-bit loads back to back, then the sum
        Vec t[V];
#pragma unroll
        for (int k = 0; k < V; ++k) t[k].u = ldg_stream(in4 + vbase + 32u * k);
#pragma unroll
        for (int k = 0; k < V; ++k)
#pragma unroll
            for (int c = 0; c < KPV; ++c) acc = acc + t[k].e[c];
    } else {
#pragma unroll 1
        for (int k = 0; k < V; ++k) {
            const uint64_t v = vbase + 32u * k;
#pragma unroll
            for (int c = 0; c < KPV; ++c)
                if (v * KPV + c < p.n) acc = acc + in1[v * KPV + c];
        }
    }
    const E ws = wide_warp_reduce(acc);
    if (lane == 0) s_red[warp] = ws;
    __syncthreads();
    if (threadIdx.x == 0) {
        E tot = s_red[0];
#pragma unroll
        for (int j = 1; j < S::NW; ++j) tot = tot + s_red[j];
        reinterpret_cast<E*>(p.partials)[tile] = tot;
    }
}

// ---- RTS kernel 2: one-block spine over the tile sums, 4 per thread per coalesced step;
// also the n == 0 path of the chained scan (writes total = carry)
template <typename E>
__global__ void __launch_bounds__(1024) wide_spine_kernel(const __grid_constant__ WideParams<E> p)
{
    pdl_entry();
    constexpr int NW = 32, ITEMS = 4;
    constexpr uint32_t CHUNK = 1024 * ITEMS;
    __shared__ __align__(16) E s_warp[2][NW];
    const uint32_t lane = lane_id(), warp = threadIdx.x >> 5;
    E* partials = reinterpret_cast<E*>(p.partials);
    E running = p.carry;
    uint32_t it = 0;
    auto load_chunk = [&](uint32_t base, E (&dst)[ITEMS]) {
        const uint32_t i0 = base + threadIdx.x * ITEMS;
#pragma unroll
        for (int i = 0; i < ITEMS; ++i) dst[i] = (i0 + i < p.num_tiles) ? partials[i0 + i] : wide_zero<E>();
    };
    E v[ITEMS], nxt[ITEMS];
    load_chunk(0, v);
    for (uint32_t base = 0; base < p.num_tiles; base += CHUNK, ++it) {
        const uint32_t i0 = base + threadIdx.x * ITEMS;
        if (base + CHUNK < p.num_tiles) load_chunk(base + CHUNK, nxt);  // in flight while this chunk is scanned
        E tsum = wide_zero<E>();
#pragma unroll
        for (int i = 0; i < ITEMS; ++i) tsum = tsum + v[i];
        const E inc = wide_warp_inclusive_scan(tsum);
        const E lane_excl = wide_exclusive_of(inc, tsum, lane);
        E* sw = s_warp[it & 1u];
        if (lane == 31) sw[warp] = inc;
        __syncthreads();
        const E raw = sw[lane];
        const E winc = wide_warp_inclusive_scan(raw);
        const E wexc = wide_exclusive_of(winc, raw, lane);
        E warp_prefix, total;
        if constexpr (std::is_same<E, u32x4>::value) {
            warp_prefix = u32x4{__shfl_sync(kFullMask, wexc.a, (int)warp), __shfl_sync(kFullMask, wexc.b, (int)warp),
                                __shfl_sync(kFullMask, wexc.c, (int)warp), __shfl_sync(kFullMask, wexc.d, (int)warp)};
            total = u32x4{__shfl_sync(kFullMask, winc.a, 31), __shfl_sync(kFullMask, winc.b, 31),
                          __shfl_sync(kFullMask, winc.c, 31), __shfl_sync(kFullMask, winc.d, 31)};
        } else {
            warp_prefix = __shfl_sync(kFullMask, wexc, (int)warp);
            total = __shfl_sync(kFullMask, winc, 31);
        }
        E run = (running + warp_prefix) + lane_excl;
#pragma unroll
        for (int i = 0; i < ITEMS; ++i) {
            if (i0 + i < p.num_tiles) partials[i0 + i] = run;
            run = run + v[i];
        }
        running = running + total;
#pragma unroll
        for (int i = 0; i < ITEMS; ++i) v[i] = nxt[i];
    }
    if (threadIdx.x == 0 && p.total_out) *reinterpret_cast<E*>(p.total_out) = running;
}

}  // namespace gps
