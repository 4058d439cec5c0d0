// gps_csdl_smem.cuh -- Chained Scan with Decoupled Lookback, shared-memory-resident tile.
//
// Third variant of the CSDL tile pipeline (same algorithm and descriptor protocol as
// gps_csdl_reg.cuh, replacing ChainedScanDecoupledLookback.cuh:79-139 of the reference).
// What the other two taught:
//   * the register path wins because MANY small independent CTAs per SM keep the loads in
//     flight -- but every byte in flight, and every byte parked behind a look-back, holds a
//     register, and registers cap it at 7 CTAs x 24 KB per SM;
//   * the TMA engine delivers a tile sooner than LDG does under load (~2.3 vs ~4.2 us
//     issue -> landed), but a persistent CTA with a ring serialises its own tiles.
// So: ONE tile per CTA as on the register path, but the tile is fetched by ONE
// `cp.async.bulk` (TMA) into shared memory and STAYS there -- no register is held while the
// load is in flight or while warp 0 walks the look-back.  Pass 1 reads the rows only to sum
// them (aggregate first); pass 2 re-reads them row by row, scans, adds the prefix and stores
// with coalesced STG.128.  Rows are lane-strided (lane l owns uint4 l + 32k of the warp's
// chunk), so LDS.128 is bank-conflict free and STG.128 fully coalesced without any
// transposition.  ~40 registers per thread; occupancy is set by shared memory instead.
#pragma once
#include "gps_tile_reg.cuh"

namespace gps {

// G: rows scanned together in pass 2 (ILP vs registers); REGS: register budget per thread
template <typename T, int NT, int V, bool INCLUSIVE, int WIN = 1, int G = 4, int REGS = 40>
__global__ void __launch_bounds__(NT, min_ctas_per_sm(NT, REGS))
    csdl_smem_kernel(const __grid_constant__ ScanParams p)
{
    constexpr int NW = NT / 32;
    constexpr uint32_t TILE_VEC = NT * V;
    constexpr uint32_t TILE_KEYS = TILE_VEC * 4;
    constexpr uint32_t TILE_BYTES = TILE_KEYS * 4;
    __shared__ __align__(128) uint4 s_data[TILE_VEC];
    __shared__ __align__(8) uint64_t s_bar;
    __shared__ T s_warp[NW];
    __shared__ T s_prefix;
    __shared__ uint32_t s_tile;

    const uint32_t tid = threadIdx.x;
    const uint32_t lane = tid & 31u;
    const uint32_t warp = tid >> 5;

    if (tid == 0) {
        mbar_init(&s_bar, 1);
        fence_mbar_init();
        const uint32_t t = (p.flags & PF_STATIC_TILE) ? blockIdx.x : claim_tile(p);
        s_tile = t;
        if (((uint64_t)t + 1u) * TILE_KEYS <= p.n) {
            mbar_arrive_expect_tx(&s_bar, TILE_BYTES);
            bulk_load(s_data, reinterpret_cast<const unsigned char*>(p.in) + (uint64_t)t * TILE_BYTES,
                      TILE_BYTES, &s_bar);
        }
    }
    __syncthreads();
    const uint32_t tile = s_tile;
    const bool full = ((uint64_t)tile + 1u) * TILE_KEYS <= p.n;
    const uint32_t row0 = warp * (32u * V) + lane;  // this thread's uint4 of row 0 (index in tile)
    const uint64_t vbase = (uint64_t)tile * TILE_VEC + row0;

    if (full) {
        if (!mbar_wait_bounded(&s_bar, 0u)) {  // never expected; do not hang the GPU
            atomicOr(p.status, STATUS_WATCHDOG);
            return;
        }
    } else {
        // partial (last) tile: guarded loads through the LSU, keys at index >= n read as 0
        const uint32_t* in1 = reinterpret_cast<const uint32_t*>(p.in);
#pragma unroll 1
        for (int k = 0; k < V; ++k) {
            const uint64_t e = (vbase + 32u * k) * 4u;
            uint4 r;
            r.x = e + 0 < p.n ? in1[e + 0] : 0u;
            r.y = e + 1 < p.n ? in1[e + 1] : 0u;
            r.z = e + 2 < p.n ? in1[e + 2] : 0u;
            r.w = e + 3 < p.n ? in1[e + 3] : 0u;
            s_data[row0 + 32u * k] = r;
        }
        // each thread only ever re-reads what it wrote itself: no barrier needed
    }

    // ---- pass 1: aggregate first
    T acc = T(0);
#pragma unroll
    for (int k = 0; k < V; ++k) {
        const uint4 r = s_data[row0 + 32u * k];
        acc += (from_bits<T>(r.x) + from_bits<T>(r.y)) + (from_bits<T>(r.z) + from_bits<T>(r.w));
    }
    const T wsum = warp_reduce(acc);
    if (lane == 0) s_warp[warp] = wsum;
    __syncthreads();
    const T wraw = lane < (uint32_t)NW ? s_warp[lane] : T(0);
    const T aggregate = warp_reduce(wraw);
    if (warp == 0 && lane == 0 && tile != 0)
        desc_publish(p, p.desc + tile, pack_desc(to_bits(aggregate), p.epoch, ST_AGGREGATE));
    const T winc = warp_inclusive_scan(wraw, lane);
    const T warp_prefix = __shfl_sync(kFullMask, warp_exclusive_of(winc, wraw, lane), (int)warp);

    if (warp == 0) {
        T prefix;
        if (tile == 0)
            prefix = acquire_carry<T>(p, lane);
        else
            prefix = (p.flags & PF_SKIP_LOOKBACK) ? T(0) : lookback_warp<T, WIN>(p, tile, lane);
        if (lane == 0) {
            const T inclusive = prefix + aggregate;
            desc_publish(p, p.desc + tile, pack_desc(to_bits(inclusive), p.epoch, ST_INCLUSIVE));
            s_prefix = prefix;
            if (tile == p.num_tiles - 1) write_total<T>(p, inclusive);
        }
    }
    __syncthreads();

    // ---- pass 2: rows in groups of 4 (bounded register use), scan, add prefix, store
    T run = s_prefix + warp_prefix;
    uint4* out4 = reinterpret_cast<uint4*>(p.out);
    uint32_t* out1 = reinterpret_cast<uint32_t*>(p.out);
#pragma unroll 1
    for (int k0 = 0; k0 < V; k0 += G) {
        T x[G][4], rs[G], inc[G];
#pragma unroll
        for (int g = 0; g < G; ++g) {
            if (k0 + g < V) {
                const uint4 r = s_data[row0 + 32u * (k0 + g)];
                x[g][0] = from_bits<T>(r.x);
                x[g][1] = from_bits<T>(r.y);
                x[g][2] = from_bits<T>(r.z);
                x[g][3] = from_bits<T>(r.w);
            } else {
                x[g][0] = x[g][1] = x[g][2] = x[g][3] = T(0);
            }
            rs[g] = ((x[g][0] + x[g][1]) + x[g][2]) + x[g][3];
            inc[g] = rs[g];
        }
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
#pragma unroll
            for (int g = 0; g < G; ++g) inc[g] = scan_step_up(inc[g], d);
        }
#pragma unroll
        for (int g = 0; g < G; ++g) {
            const T ex = warp_exclusive_of(inc[g], rs[g], lane);
            const T tot = __shfl_sync(kFullMask, inc[g], 31);
            const T b = run + ex;
            run += tot;
            if (k0 + g < V) {
                T o0, o1, o2, o3;
                if (INCLUSIVE) {
                    o0 = b + x[g][0];
                    o1 = o0 + x[g][1];
                    o2 = o1 + x[g][2];
                    o3 = o2 + x[g][3];
                } else {
                    o0 = b;
                    o1 = o0 + x[g][0];
                    o2 = o1 + x[g][1];
                    o3 = o2 + x[g][2];
                }
                const uint4 r = make_uint4(to_bits(o0), to_bits(o1), to_bits(o2), to_bits(o3));
                const uint64_t v = vbase + 32u * (k0 + g);
                if (full || (v + 1u) * 4u <= p.n) {
                    stg_stream(out4 + v, r);
                } else {
                    const uint64_t e = v * 4u;
                    if (e + 0 < p.n) out1[e + 0] = r.x;
                    if (e + 1 < p.n) out1[e + 1] = r.y;
                    if (e + 2 < p.n) out1[e + 2] = r.z;
                }
            }
        }
    }
}

}  // namespace gps
