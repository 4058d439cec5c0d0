// gps_tile_reg.cuh -- register-resident tile: 128-bit coalesced loads, warp-shuffle
// scan, 128-bit coalesced stores.
//
// Replaces ScanCommon.cuh:12-135 (Scan{Inclusive,Exclusive}{Full,Partial},
// Propagate{Full,Partial}) of the reference.  Same data mapping -- warp w owns 32*V
// consecutive uint4, lane l reads uint4 (l + 32k) of it for k < V, so each LDG.128 /
// STG.128 of a warp is one fully coalesced 512-byte request -- but:
//   * the V row scans are issued interleaved (V independent shuffle chains in flight)
//     instead of a serial scan-rotate-broadcast per row (6 SHFL + 1 broadcast each,
//     ScanCommon.cuh:86-87);
//   * inclusive vs exclusive is decided only at store time from the same registers, so
//     one code path serves both kinds and both dtypes;
//   * the partial tile never touches memory past n (the reference rounds the buffer up
//     to 4 keys and writes the whole last uint4, ScanCommon.cuh:126-134);
//   * 64-bit indexing (n may exceed 2^32 on a 2-GPU shard of 2^33 keys).
#pragma once
#include "gps_common.cuh"

namespace gps {

template <typename T, int NT, int V>
struct RegTile {
    static constexpr int NW = NT / 32;
    static constexpr uint32_t TILE_VEC = NT * V;       // uint4 per tile
    static constexpr uint32_t TILE_KEYS = TILE_VEC * 4;

    T x[V][4];       // raw keys of this thread (row k, component c)
    T lane_excl[V];  // exclusive prefix of (row k, this lane) within the warp's chunk
    T warp_total;

    // first uint4 index (within the whole array) of row 0 for this thread
    __device__ __forceinline__ static uint64_t vec_base(uint64_t tile, uint32_t warp, uint32_t lane)
    {
        return tile * TILE_VEC + (uint64_t)warp * (32u * V) + lane;
    }

    __device__ __forceinline__ void set_row(int k, uint4 r)
    {
        x[k][0] = from_bits<T>(r.x);
        x[k][1] = from_bits<T>(r.y);
        x[k][2] = from_bits<T>(r.z);
        x[k][3] = from_bits<T>(r.w);
    }

    // full tile: V back-to-back LDG.128, nothing else between them
    // `allocate` (experiment switch, warp-uniform): plain L1-allocating accesses
    __device__ __forceinline__ void load_full(const void* in, uint64_t vbase, bool allocate = false)
    {
        const uint4* in4 = reinterpret_cast<const uint4*>(in);
        uint4 r[V];
        if (allocate) {
#pragma unroll
            for (int k = 0; k < V; ++k) r[k] = in4[vbase + 32u * k];
        } else {
#pragma unroll
            for (int k = 0; k < V; ++k) r[k] = ldg_stream(in4 + vbase + 32u * k);
        }
#pragma unroll
        for (int k = 0; k < V; ++k) set_row(k, r[k]);
    }

    // full tile, every load tagged with an L2 eviction policy (createpolicy descriptor)
    __device__ __forceinline__ void load_full_policy(const void* in, uint64_t vbase, uint64_t policy)
    {
        const uint4* in4 = reinterpret_cast<const uint4*>(in);
        uint4 r[V];
#pragma unroll
        for (int k = 0; k < V; ++k) {
            const uint4* a = in4 + vbase + 32u * k;
            asm volatile("ld.global.L1::no_allocate.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;"
                         : "=r"(r[k].x), "=r"(r[k].y), "=r"(r[k].z), "=r"(r[k].w)
                         : "l"(a), "l"(policy)
                         : "memory");
        }
#pragma unroll
        for (int k = 0; k < V; ++k) set_row(k, r[k]);
    }

    // partial (last) tile: keys at index >= n read as 0
    __device__ __forceinline__ void load_partial(const void* in, uint64_t vbase, uint64_t n)
    {
        const uint32_t* in1 = reinterpret_cast<const uint32_t*>(in);
#pragma unroll
        for (int k = 0; k < V; ++k) {
            const uint64_t e = (vbase + 32u * k) * 4u;
            if (e + 4u <= n) {
                set_row(k, ldg_stream(reinterpret_cast<const uint4*>(in1 + e)));
            } else {
#pragma unroll
                for (int c = 0; c < 4; ++c)
                    x[k][c] = (e + c < n) ? from_bits<T>(in1[e + c]) : T(0);
            }
        }
    }

    // Cold path only: park the registers of this tile in (shared) memory and fetch them back, lane-major (conflict free).
    // Between the two calls nothing of the tile is live, so a function CALL there costs the hot path no spill.
    static constexpr uint32_t PARK_WORDS = 5u * V + 1u;  // per lane
    __device__ __forceinline__ void park(T* s, uint32_t lane) const
    {
#pragma unroll
        for (int k = 0; k < V; ++k) {
#pragma unroll
            for (int c = 0; c < 4; ++c) s[(4 * k + c) * 32 + lane] = x[k][c];
            s[(4 * V + k) * 32 + lane] = lane_excl[k];
        }
        s[5 * V * 32 + lane] = warp_total;
    }
    __device__ __forceinline__ void unpark(const T* s, uint32_t lane)
    {
#pragma unroll
        for (int k = 0; k < V; ++k) {
#pragma unroll
            for (int c = 0; c < 4; ++c) x[k][c] = s[(4 * k + c) * 32 + lane];
            lane_excl[k] = s[(4 * V + k) * 32 + lane];
        }
        warp_total = s[5 * V * 32 + lane];
    }

    // warp-level scan of the V rows; leaves lane_excl[] and warp_total
    __device__ __forceinline__ void scan_warp(uint32_t lane)
    {
        T rs[V], inc[V];
#pragma unroll
        for (int k = 0; k < V; ++k) {
            rs[k] = ((x[k][0] + x[k][1]) + x[k][2]) + x[k][3];
            inc[k] = rs[k];
        }
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
#pragma unroll
            for (int k = 0; k < V; ++k) inc[k] = scan_step_up(inc[k], d);
        }
        T run = T(0);
#pragma unroll
        for (int k = 0; k < V; ++k) {
            const T ex = warp_exclusive_of(inc[k], rs[k], lane);
            const T tot = __shfl_sync(kFullMask, inc[k], 31);
            lane_excl[k] = run + ex;
            run += tot;
        }
        warp_total = run;
    }

    // sum only (Reduce pass / fallback): no per-lane prefixes
    __device__ __forceinline__ T thread_sum() const
    {
        T s = T(0);
#pragma unroll
        for (int k = 0; k < V; ++k) s += ((x[k][0] + x[k][1]) + x[k][2]) + x[k][3];
        return s;
    }

    template <bool INCLUSIVE>
    __device__ __forceinline__ uint4 finish_row(int k, T prefix) const
    {
        const T b = prefix + lane_excl[k];
        T o0, o1, o2, o3;
        if (INCLUSIVE) {
            o0 = b + x[k][0];
            o1 = o0 + x[k][1];
            o2 = o1 + x[k][2];
            o3 = o2 + x[k][3];
        } else {
            o0 = b;
            o1 = o0 + x[k][0];
            o2 = o1 + x[k][1];
            o3 = o2 + x[k][2];
        }
        return make_uint4(to_bits(o0), to_bits(o1), to_bits(o2), to_bits(o3));
    }

    // prefix = tile prefix + this warp's prefix within the tile
    template <bool INCLUSIVE>
    __device__ __forceinline__ void store_full(void* out, uint64_t vbase, T prefix, bool allocate = false) const
    {
        uint4* out4 = reinterpret_cast<uint4*>(out);
        if (allocate) {
#pragma unroll
            for (int k = 0; k < V; ++k) out4[vbase + 32u * k] = finish_row<INCLUSIVE>(k, prefix);
        } else {
#pragma unroll
            for (int k = 0; k < V; ++k) stg_stream(out4 + vbase + 32u * k, finish_row<INCLUSIVE>(k, prefix));
        }
    }

    template <bool INCLUSIVE>
    __device__ __forceinline__ void store_partial(void* out, uint64_t vbase, uint64_t n, T prefix) const
    {
        uint32_t* out1 = reinterpret_cast<uint32_t*>(out);
#pragma unroll
        for (int k = 0; k < V; ++k) {
            const uint64_t e = (vbase + 32u * k) * 4u;
            const uint4 r = finish_row<INCLUSIVE>(k, prefix);
            if (e + 4u <= n) {
                stg_stream(reinterpret_cast<uint4*>(out1 + e), r);
            } else {
                if (e + 0 < n) out1[e + 0] = r.x;
                if (e + 1 < n) out1[e + 1] = r.y;
                if (e + 2 < n) out1[e + 2] = r.z;
            }
        }
    }
};

}  // namespace gps
