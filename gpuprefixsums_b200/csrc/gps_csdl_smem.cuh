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
#include "gps_csdl_blocked.cuh"
#include "gps_csdl_reg.cuh"  // FallbackParams, LB_* states
#include "gps_tile_reg.cuh"

namespace gps {

// pass 2 of the smem-resident tile: rows in groups of G (bounded register use), scan, add
// the prefix, store with coalesced 128-bit stores; `run0` = tile prefix + warp prefix
template <typename T, int V, bool INCLUSIVE, int G>
__device__ __forceinline__ void smem_tile_finish(const uint4* s_data, uint32_t row0, uint64_t vbase, bool full,
                                                 uint64_t n, void* out, uint32_t lane, T run0, uint32_t flags = 0u)
{
    T run = run0;
    const bool keep = flags & PF_L2_OUT_EVICT_LAST;  // warp-uniform
    const uint64_t pol = keep ? l2_policy_evict_last() : 0ull;
    uint4* out4 = reinterpret_cast<uint4*>(out);
    uint32_t* out1 = reinterpret_cast<uint32_t*>(out);
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
                if (full || (v + 1u) * 4u <= n) {
                    if (keep)
                        stg_stream_hint(out4 + v, r, pol);
                    else
                        stg_stream(out4 + v, r);
                } else {
                    const uint64_t e = v * 4u;
                    if (e + 0 < n) out1[e + 0] = r.x;
                    if (e + 1 < n) out1[e + 1] = r.y;
                    if (e + 2 < n) out1[e + 2] = r.z;
                }
            }
        }
    }
}

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
    pdl_entry();

    if (tid == 0) {
        mbar_init(&s_bar, 1);
        fence_mbar_init();
        const uint32_t t = (p.flags & PF_STATIC_TILE) ? blockIdx.x : claim_tile(p);
        s_tile = t;
        if (((uint64_t)t + 1u) * TILE_KEYS <= p.n) {
            mbar_arrive_expect_tx(&s_bar, TILE_BYTES);
            tile_bulk_load(p.flags, s_data, reinterpret_cast<const unsigned char*>(p.in) + (uint64_t)t * TILE_BYTES,
                           TILE_BYTES, &s_bar);
        }
    }
    __syncthreads();
    const uint32_t tile = s_tile;
    if (tile >= p.num_tiles) {  // ticket accounting out of step (an earlier launch failed): never index past the descriptors
        if (threadIdx.x == 0) atomicOr(p.status, STATUS_WATCHDOG);
        return;
    }
    const bool full = ((uint64_t)tile + 1u) * TILE_KEYS <= p.n;
    const uint32_t row0 = warp * (32u * V) + lane;  // this thread's uint4 of row 0 (index in tile)
    const uint64_t vbase = (uint64_t)tile * TILE_VEC + row0;

    // ---- eager look-back (PF_EAGER_LOOKBACK, full tiles other than tile 0; block-uniform): the walk over the
    // predecessors' descriptors needs nothing of THIS tile, so warp 0 starts it right after the claim and it
    // runs while the TMA load is in flight; warps 1.. reduce the whole tile (warp 0's rows dealt round-robin
    // among them) and publish the aggregate without waiting for warp 0.  The ~2 L2 round trips of the look-back
    // leave the CTA's critical path: load -> reduce -> (prefix already there) -> scan + store.
    // MEASURED (profiles/r02_ticket_free.txt A): bit-exact, and no gain at any size -- what a mid-size CTA waits for is the
    // slowest predecessor LOAD, not the polls.  Kept for the tuning harness only (no library option sets the flag).
    if constexpr (NW >= 3) {
        if ((p.flags & PF_EAGER_LOOKBACK) && full && tile != 0u) {
            __shared__ T s_part[NW];
            if (warp == 0) {
                const T prefix = lookback_warp<T, WIN, GPS_LOOKBACK_MAXWIN>(p, tile, lane);
                if (lane == 0) s_prefix = prefix;
            } else {
                if (!mbar_wait_bounded(&s_bar, 0u)) atomicOr(p.status, STATUS_WATCHDOG);  // never expected
                T acc = T(0), acc0 = T(0);
#pragma unroll
                for (int k = 0; k < V; ++k) {
                    const uint4 r = s_data[row0 + 32u * k];
                    acc += (from_bits<T>(r.x) + from_bits<T>(r.y)) + (from_bits<T>(r.z) + from_bits<T>(r.w));
                }
#pragma unroll 1
                for (int k = (int)warp - 1; k < V; k += NW - 1) {
                    const uint4 r = s_data[lane + 32u * k];
                    acc0 += (from_bits<T>(r.x) + from_bits<T>(r.y)) + (from_bits<T>(r.z) + from_bits<T>(r.w));
                }
                const T wsum = warp_reduce(acc), w0 = warp_reduce(acc0);
                if (lane == 0) {
                    s_warp[warp] = wsum;
                    s_part[warp] = w0;
                }
                named_bar_sync(1u, NT - 32u);
                if (warp == 1) {
                    const T part = (lane >= 1u && lane < (uint32_t)NW) ? s_part[lane] : T(0);
                    const T w0sum = warp_reduce(part);
                    if (lane == 0) s_warp[0] = w0sum;
                    __syncwarp();
                    const T wr = lane < (uint32_t)NW ? s_warp[lane] : T(0);
                    const T agg = warp_reduce(wr);  // the same operation every warp repeats below: same bits
                    if (lane == 0) desc_publish(p, p.desc + tile, pack_desc(to_bits(agg), p.epoch, ST_AGGREGATE));
                }
            }
            __syncthreads();
            if (warp == 0) (void)mbar_try_wait(&s_bar, 0u);  // phase long complete: orders this warp's reads after the TMA writes
            const T wraw = lane < (uint32_t)NW ? s_warp[lane] : T(0);
            const T aggregate = warp_reduce(wraw);
            const T winc = warp_inclusive_scan(wraw, lane);
            const T warp_prefix = __shfl_sync(kFullMask, warp_exclusive_of(winc, wraw, lane), (int)warp);
            const T prefix = s_prefix;
            if (tid == 0) {
                const T inclusive = prefix + aggregate;
                desc_publish(p, p.desc + tile, pack_desc(to_bits(inclusive), p.epoch, ST_INCLUSIVE));
                if (tile == p.num_tiles - 1) write_total<T>(p, inclusive);
            }
            smem_tile_finish<T, V, INCLUSIVE, G>(s_data, row0, vbase, true, p.n, p.out, lane, prefix + warp_prefix, p.flags);
            return;
        }
    }

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
            prefix = (p.flags & PF_SKIP_LOOKBACK) ? T(0) : lookback_warp<T, WIN, GPS_LOOKBACK_MAXWIN>(p, tile, lane);
        if (lane == 0) {
            const T inclusive = prefix + aggregate;
            desc_publish(p, p.desc + tile, pack_desc(to_bits(inclusive), p.epoch, ST_INCLUSIVE));
            s_prefix = prefix;
            if (tile == p.num_tiles - 1) write_total<T>(p, inclusive);
        }
    }
    __syncthreads();

    smem_tile_finish<T, V, INCLUSIVE, G>(s_data, row0, vbase, full, p.n, p.out, lane, s_prefix + warp_prefix, p.flags);
}

// ---------------------------------------------------------------------------------
// Persistent one-wave variant: the grid is at most ONE wave of CTAs (the host passes gridDim.x <= CTAs the
// device holds at once) and CTA c scans tiles c, c + gridDim.x, c + 2 gridDim.x, ... in that order.  No ticket
// and no dependence on the dispatch order: every tile below the one a CTA waits for belongs to a CTA that is
// resident and works through ITS tiles in increasing order, so by induction on the tile index every wait ends
// (given that all CTAs of the grid do become resident, i.e. that other kernels on the device drain).  Per
// tile it saves the ticket's round trip and the launch of a new CTA; the tile pipeline is csdl_smem_kernel's.
// MEASURED (profiles/r02_ticket_free.txt I): no gain from 2^22 to 2^28 -- a looping CTA serialises load -> look-back ->
// store per slot exactly like a fresh CTA.  Instantiated by the tuning harness only.
// ---------------------------------------------------------------------------------
template <typename T, int NT, int V, bool INCLUSIVE, int WIN = 1, int G = 4, int REGS = 40>
__global__ void __launch_bounds__(NT, min_ctas_per_sm(NT, REGS))
    csdl_smem_persistent_kernel(const __grid_constant__ ScanParams p)
{
    constexpr int NW = NT / 32;
    constexpr uint32_t TILE_VEC = NT * V;
    constexpr uint32_t TILE_KEYS = TILE_VEC * 4;
    constexpr uint32_t TILE_BYTES = TILE_KEYS * 4;
    __shared__ __align__(128) uint4 s_data[TILE_VEC];
    __shared__ __align__(8) uint64_t s_bar;
    __shared__ T s_warp[NW];
    __shared__ T s_prefix;

    const uint32_t tid = threadIdx.x;
    const uint32_t lane = tid & 31u;
    const uint32_t warp = tid >> 5;
    const uint32_t row0 = warp * (32u * V) + lane;
    pdl_entry();
    if (tid == 0) {
        mbar_init(&s_bar, 1);
        fence_mbar_init();
    }
    __syncthreads();

    uint32_t parity = 0;
#pragma unroll 1
    for (uint32_t tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        const bool full = ((uint64_t)tile + 1u) * TILE_KEYS <= p.n;  // block-uniform
        const uint64_t vbase = (uint64_t)tile * TILE_VEC + row0;
        if (full) {
            if (tid == 0) {
                mbar_arrive_expect_tx(&s_bar, TILE_BYTES);
                tile_bulk_load(p.flags, s_data, reinterpret_cast<const unsigned char*>(p.in) + (uint64_t)tile * TILE_BYTES,
                               TILE_BYTES, &s_bar);
            }
            if (!mbar_wait_bounded(&s_bar, parity)) {  // never expected; do not hang the GPU
                atomicOr(p.status, STATUS_WATCHDOG);
                return;
            }
            parity ^= 1u;
        } else {
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
            const T prefix = tile == 0 ? acquire_carry<T>(p, lane) : lookback_warp<T, WIN, GPS_LOOKBACK_MAXWIN>(p, tile, lane);
            if (lane == 0) {
                const T inclusive = prefix + aggregate;
                desc_publish(p, p.desc + tile, pack_desc(to_bits(inclusive), p.epoch, ST_INCLUSIVE));
                s_prefix = prefix;
                if (tile == p.num_tiles - 1) write_total<T>(p, inclusive);
            }
        }
        __syncthreads();
        smem_tile_finish<T, V, INCLUSIVE, G>(s_data, row0, vbase, full, p.n, p.out, lane, s_prefix + warp_prefix, p.flags);
        __syncthreads();  // the tile buffer, s_warp and s_prefix are reused by the next tile
    }
}

// ---------------------------------------------------------------------------------
// The same tile pipeline with Decoupled Fallback (and optional emulated stalls): the
// smem-resident counterpart of csdl_fallback_kernel in gps_csdl_reg.cuh, which documents the
// protocol (EmulatedDeadlocking.cuh:63-151 of the reference: spin budget, whole-CTA
// re-reduction of the blocking tile from the INPUT, insertion with atomicCAS).  The fallback
// target is re-read from global memory with 128-bit loads; this CTA's own tile stays parked
// in shared memory meanwhile.
// ---------------------------------------------------------------------------------
template <typename T, int NT, int V, bool INCLUSIVE, int G = 4, int REGS = 40>
__global__ void __launch_bounds__(NT, min_ctas_per_sm(NT, REGS))
    csdl_smem_fallback_kernel(const __grid_constant__ ScanParams p, const FallbackParams f)
{
    pdl_entry();
    constexpr int NW = NT / 32;
    constexpr uint32_t TILE_VEC = NT * V;
    constexpr uint32_t TILE_KEYS = TILE_VEC * 4;
    constexpr uint32_t TILE_BYTES = TILE_KEYS * 4;
    __shared__ __align__(128) uint4 s_data[TILE_VEC];
    __shared__ __align__(8) uint64_t s_bar;
    __shared__ T s_warp[NW];
    __shared__ T s_red[NW];
    __shared__ T s_prefix;
    __shared__ uint32_t s_tile;
    __shared__ uint32_t s_state;    // LB_DONE / LB_FALLBACK
    __shared__ uint32_t s_fb_tile;  // tile to recompute

    const uint32_t tid = threadIdx.x;
    const uint32_t lane = tid & 31u;
    const uint32_t warp = tid >> 5;

    if (tid == 0) {
        mbar_init(&s_bar, 1);
        fence_mbar_init();
        // Decoupled Fallback needs no ticket: a CTA never depends on another CTA making progress, so tile =
        // blockIdx.x is deadlock-free under ANY dispatch order (PF_STATIC_TILE; plain CSDL with static tiles
        // relies on in-order dispatch, as cub::DeviceScan does)
        const uint32_t t = (p.flags & PF_STATIC_TILE) ? blockIdx.x : claim_tile(p);
        s_tile = t;
        if (((uint64_t)t + 1u) * TILE_KEYS <= p.n) {
            mbar_arrive_expect_tx(&s_bar, TILE_BYTES);
            tile_bulk_load(p.flags, s_data, reinterpret_cast<const unsigned char*>(p.in) + (uint64_t)t * TILE_BYTES,
                           TILE_BYTES, &s_bar);
        }
    }
    __syncthreads();
    const uint32_t tile = s_tile;
    if (tile >= p.num_tiles) {  // ticket accounting out of step (an earlier launch failed): never index past the descriptors
        if (threadIdx.x == 0) atomicOr(p.status, STATUS_WATCHDOG);
        return;
    }
    const bool full = ((uint64_t)tile + 1u) * TILE_KEYS <= p.n;
    const uint32_t row0 = warp * (32u * V) + lane;
    const uint64_t vbase = (uint64_t)tile * TILE_VEC + row0;

    if (full) {
        if (!mbar_wait_bounded(&s_bar, 0u)) {
            atomicOr(p.status, STATUS_WATCHDOG);
            return;
        }
    } else {
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
    const T winc = warp_inclusive_scan(wraw, lane);
    const T warp_prefix = __shfl_sync(kFullMask, warp_exclusive_of(winc, wraw, lane), (int)warp);

    // stall emulation: the tile refuses to publish until a successor's fallback published for
    // it; the last tile never stalls (nobody could rescue it)
    const bool stalled = f.stall_mask != kNoStall && (tile & f.stall_mask) == 0u && p.num_tiles > 1u &&
                         tile != p.num_tiles - 1u;
    T carry = T(0);
    if (warp == 0) {
        if (tile == 0) carry = acquire_carry<T>(p, lane);
        if (lane == 0) {
            if (stalled) {
                uint32_t polls = 0;
                while (!desc_valid(ld_relaxed_gpu(p.desc + tile), p.epoch)) {
                    if (watchdog_tick(polls, p.status)) break;
                }
            }
            const T v = tile ? aggregate : carry + aggregate;
            atomicExch(reinterpret_cast<unsigned long long*>(p.desc + tile),
                       (unsigned long long)pack_desc(to_bits(v), p.epoch, tile ? ST_AGGREGATE : ST_INCLUSIVE));
            // in-place safety of the fallback: the descriptor is performed at L2 before any
            // thread of this CTA overwrites the tile's input (fence + the barrier below), so a
            // successor re-reading that input afterwards fails its CAS and discards the sum.
            // Out of place the input is never overwritten and the descriptor carries its
            // payload in the same word: no ordering against other memory is needed.
            if (p.in == p.out) __threadfence();
        }
    }
    __syncthreads();  // holds the whole block while stalled ("purely for emulation")

    uint32_t n_spins = 0, n_fallbacks = 0, n_inserts = 0, n_windows = 0;
    T prefix = carry;  // tile 0: done
    if (tile != 0) {
        int64_t top = (int64_t)tile - 1;
        prefix = T(0);
        while (true) {
            if (warp == 0) {
                uint32_t state = LB_DONE, spins = 0, polls = 0, fb = 0;
                while (true) {
                    const int64_t idx = top - (int64_t)lane;
                    uint64_t d = pack_desc(0u, p.epoch, ST_INCLUSIVE);
                    bool ready = true;
                    if (idx >= 0) {
                        d = ld_relaxed_gpu(p.desc + idx);  // payload and flag share the word
                        ready = desc_valid(d, p.epoch);
                    }
                    ++n_windows;
                    const uint32_t not_ready = __ballot_sync(kFullMask, !ready);
                    const uint32_t first_nr = not_ready ? (uint32_t)__ffs((int)not_ready) - 1u : 32u;
                    const uint32_t below = first_nr >= 32u ? kFullMask : ((1u << first_nr) - 1u);
                    const uint32_t incl = __ballot_sync(kFullMask, ready && desc_status(d) == ST_INCLUSIVE) & below;
                    const T v = from_bits<T>(desc_value(d));
                    if (incl) {
                        const uint32_t nearest = (uint32_t)__ffs((int)incl) - 1u;
                        prefix += warp_reduce(lane <= nearest ? v : T(0));
                        break;
                    }
                    if (first_nr == 32u) {  // whole window ready, none inclusive
                        prefix += warp_reduce(v);
                        top -= 32;
                        spins = 0;
                        continue;
                    }
                    ++n_spins;
                    const bool bail = watchdog_tick(polls, p.status);
                    if (__any_sync(kFullMask, bail)) break;
                    if (++spins < f.max_spin) {
                        if (p.poll_sleep_ns) __nanosleep(p.poll_sleep_ns);
                        continue;
                    }
                    // budget exhausted: consume the ready tiles nearer than the blocker, fall
                    // back on the blocker itself
                    prefix += warp_reduce(lane < first_nr ? v : T(0));
                    fb = (uint32_t)(top - (int64_t)first_nr);
                    top = (int64_t)fb - 1;
                    state = LB_FALLBACK;
                    break;
                }
                if (lane == 0) {
                    s_state = state;
                    s_fb_tile = fb;
                }
            }
            __syncthreads();
            if (s_state == LB_DONE) break;

            // ---- Decoupled Fallback: every thread helps to re-reduce tile fb from the INPUT
            // (a fallback target is never the last tile => always full)
            const uint32_t fb = s_fb_tile;
            ++n_fallbacks;
            {
                const uint4* in4 = reinterpret_cast<const uint4*>(p.in) + (uint64_t)fb * TILE_VEC + row0;
                T fs = T(0);
#pragma unroll
                for (int k = 0; k < V; ++k) {
                    const uint4 r = ldg_stream(in4 + 32u * k);
                    fs += (from_bits<T>(r.x) + from_bits<T>(r.y)) + (from_bits<T>(r.z) + from_bits<T>(r.w));
                }
                const T ws = warp_reduce(fs);
                if (lane == 0) s_red[warp] = ws;
            }
            __syncthreads();
            if (warp == 0) {
                T red = warp_reduce(lane < (uint32_t)NW ? s_red[lane] : T(0));
                if (fb == 0) red += acquire_carry<T>(p, lane);  // tile 0 is INCLUSIVE incl. carry
                uint32_t done = 0;
                if (lane == 0) {
                    const uint32_t st = fb ? ST_AGGREGATE : ST_INCLUSIVE;
                    const uint64_t mine = pack_desc(to_bits(red), p.epoch, st);
                    uint64_t seen = ld_relaxed_gpu(p.desc + fb);
                    T use = red;
                    uint32_t use_st = st;
                    bool inserted = false;
                    while (!desc_valid(seen, p.epoch)) {  // CAS from the stale word we observe to ours
                        const uint64_t old = atomicCAS(reinterpret_cast<unsigned long long*>(p.desc + fb),
                                                       (unsigned long long)seen, (unsigned long long)mine);
                        if (old == seen) {
                            inserted = true;
                            break;
                        }
                        seen = old;
                    }
                    if (!inserted) {  // the owner (or another rescuer) got there first
                        use = from_bits<T>(desc_value(seen));
                        use_st = desc_status(seen);
                    } else {
                        ++n_inserts;
                    }
                    prefix += use;
                    done = (use_st == ST_INCLUSIVE || fb == 0) ? 1u : 0u;
                }
                prefix = __shfl_sync(kFullMask, prefix, 0);
                done = __shfl_sync(kFullMask, done, 0);
                if (lane == 0) s_state = done ? LB_DONE : LB_FALLBACK;
            }
            __syncthreads();
            if (s_state == LB_DONE) break;
            __syncthreads();  // s_state is rewritten by the next look-back round
        }
    }

    if (warp == 0 && lane == 0) {
        const T inclusive = prefix + aggregate;
        if (tile != 0)
            atomicExch(reinterpret_cast<unsigned long long*>(p.desc + tile),
                       (unsigned long long)pack_desc(to_bits(inclusive), p.epoch, ST_INCLUSIVE));
        s_prefix = prefix;
        if (tile == p.num_tiles - 1) write_total<T>(p, inclusive);
        if (f.stats) {
            if (n_spins) atomicAdd(f.stats + 0, n_spins);
            if (n_inserts) atomicAdd(f.stats + 2, n_inserts);
            if (n_windows) atomicAdd(f.stats + 3, n_windows);
        }
    }
    if (tid == 0 && f.stats && n_fallbacks) atomicAdd(f.stats + 1, n_fallbacks);
    __syncthreads();

    smem_tile_finish<T, V, INCLUSIVE, G>(s_data, row0, vbase, full, p.n, p.out, lane, s_prefix + warp_prefix, p.flags);
}


// ---- and on the blocked tile of gps_csdl_blocked.cuh (same protocol, same fallback re-read)
template <typename T, int NT, int V, bool INCLUSIVE, int REGS = 40>
__global__ void __launch_bounds__(NT, min_ctas_per_sm(NT, REGS))
    csdl_blocked_fallback_kernel(const __grid_constant__ ScanParams p, const FallbackParams f)
{
    pdl_entry();
    using Tile = BlockedTile<T, NT, V>;
    constexpr int NW = Tile::NW;
    constexpr uint32_t TILE_VEC = Tile::TILE_VEC;
    __shared__ __align__(128) uint4 s_data[TILE_VEC];
    __shared__ __align__(8) uint64_t s_bar;
    __shared__ T s_warp[NW];
    __shared__ T s_red[NW];
    __shared__ T s_prefix;
    __shared__ uint32_t s_tile;
    __shared__ uint32_t s_state;    // LB_DONE / LB_FALLBACK
    __shared__ uint32_t s_fb_tile;  // tile to recompute

    const uint32_t tid = threadIdx.x;
    const uint32_t lane = tid & 31u;
    const uint32_t warp = tid >> 5;

    if (tid == 0) {
        mbar_init(&s_bar, 1);
        fence_mbar_init();
        const uint32_t t = (p.flags & PF_STATIC_TILE) ? blockIdx.x : claim_tile(p);  // ticket-free: see csdl_smem_fallback_kernel
        s_tile = t;
        Tile::issue_load(s_data, &s_bar, p.in, t, p.n);
    }
    __syncthreads();
    const uint32_t tile = s_tile;
    if (tile >= p.num_tiles) {  // ticket accounting out of step (an earlier launch failed): never index past the descriptors
        if (threadIdx.x == 0) atomicOr(p.status, STATUS_WATCHDOG);
        return;
    }
    const bool full = ((uint64_t)tile + 1u) * Tile::TILE_KEYS <= p.n;
    const uint32_t v0 = tid * V;                     // this thread's first uint4 within the tile
    const uint32_t row0 = warp * (32u * V) + lane;   // lane-strided start for the fallback re-read
    if (!Tile::arrive(s_data, &s_bar, p.in, tile, p.n, full, v0)) {
        atomicOr(p.status, STATUS_WATCHDOG);
        return;
    }

    // ---- pass 1: thread total -> one warp scan per tile -> aggregate first
    const T acc = Tile::thread_sum(s_data, v0);
    const T tinc = warp_inclusive_scan(acc, lane);
    const T thread_excl = warp_exclusive_of(tinc, acc, lane);
    if (lane == 31) s_warp[warp] = tinc;
    __syncthreads();
    T warp_prefix = T(0), aggregate = T(0);
#pragma unroll
    for (int j = 0; j < NW; ++j) {
        const T w = s_warp[j];
        if ((uint32_t)j < warp) warp_prefix += w;
        aggregate += w;
    }

    // stall emulation: the tile refuses to publish until a successor's fallback published for
    // it; the last tile never stalls (nobody could rescue it)
    const bool stalled = f.stall_mask != kNoStall && (tile & f.stall_mask) == 0u && p.num_tiles > 1u &&
                         tile != p.num_tiles - 1u;
    T carry = T(0);
    if (warp == 0) {
        if (tile == 0) carry = acquire_carry<T>(p, lane);
        if (lane == 0) {
            if (stalled) {
                uint32_t polls = 0;
                while (!desc_valid(ld_relaxed_gpu(p.desc + tile), p.epoch)) {
                    if (watchdog_tick(polls, p.status)) break;
                }
            }
            const T v = tile ? aggregate : carry + aggregate;
            atomicExch(reinterpret_cast<unsigned long long*>(p.desc + tile),
                       (unsigned long long)pack_desc(to_bits(v), p.epoch, tile ? ST_AGGREGATE : ST_INCLUSIVE));
            // in-place safety of the fallback: the descriptor is performed at L2 before any
            // thread of this CTA overwrites the tile's input (fence + the barrier below), so a
            // successor re-reading that input afterwards fails its CAS and discards the sum.
            // Out of place the input is never overwritten and the descriptor carries its
            // payload in the same word: no ordering against other memory is needed.
            if (p.in == p.out) __threadfence();
        }
    }
    __syncthreads();  // holds the whole block while stalled ("purely for emulation")

    uint32_t n_spins = 0, n_fallbacks = 0, n_inserts = 0, n_windows = 0;
    T prefix = carry;  // tile 0: done
    if (tile != 0) {
        int64_t top = (int64_t)tile - 1;
        prefix = T(0);
        while (true) {
            if (warp == 0) {
                uint32_t state = LB_DONE, spins = 0, polls = 0, fb = 0;
                while (true) {
                    const int64_t idx = top - (int64_t)lane;
                    uint64_t d = pack_desc(0u, p.epoch, ST_INCLUSIVE);
                    bool ready = true;
                    if (idx >= 0) {
                        d = ld_relaxed_gpu(p.desc + idx);  // payload and flag share the word
                        ready = desc_valid(d, p.epoch);
                    }
                    ++n_windows;
                    const uint32_t not_ready = __ballot_sync(kFullMask, !ready);
                    const uint32_t first_nr = not_ready ? (uint32_t)__ffs((int)not_ready) - 1u : 32u;
                    const uint32_t below = first_nr >= 32u ? kFullMask : ((1u << first_nr) - 1u);
                    const uint32_t incl = __ballot_sync(kFullMask, ready && desc_status(d) == ST_INCLUSIVE) & below;
                    const T v = from_bits<T>(desc_value(d));
                    if (incl) {
                        const uint32_t nearest = (uint32_t)__ffs((int)incl) - 1u;
                        prefix += warp_reduce(lane <= nearest ? v : T(0));
                        break;
                    }
                    if (first_nr == 32u) {  // whole window ready, none inclusive
                        prefix += warp_reduce(v);
                        top -= 32;
                        spins = 0;
                        continue;
                    }
                    ++n_spins;
                    const bool bail = watchdog_tick(polls, p.status);
                    if (__any_sync(kFullMask, bail)) break;
                    if (++spins < f.max_spin) {
                        if (p.poll_sleep_ns) __nanosleep(p.poll_sleep_ns);
                        continue;
                    }
                    // budget exhausted: consume the ready tiles nearer than the blocker, fall
                    // back on the blocker itself
                    prefix += warp_reduce(lane < first_nr ? v : T(0));
                    fb = (uint32_t)(top - (int64_t)first_nr);
                    top = (int64_t)fb - 1;
                    state = LB_FALLBACK;
                    break;
                }
                if (lane == 0) {
                    s_state = state;
                    s_fb_tile = fb;
                }
            }
            __syncthreads();
            if (s_state == LB_DONE) break;

            // ---- Decoupled Fallback: every thread helps to re-reduce tile fb from the INPUT
            // (a fallback target is never the last tile => always full)
            const uint32_t fb = s_fb_tile;
            ++n_fallbacks;
            {
                const uint4* in4 = reinterpret_cast<const uint4*>(p.in) + (uint64_t)fb * TILE_VEC + row0;
                T fs = T(0);
#pragma unroll
                for (int k = 0; k < V; ++k) {
                    const uint4 r = ldg_stream(in4 + 32u * k);
                    fs += (from_bits<T>(r.x) + from_bits<T>(r.y)) + (from_bits<T>(r.z) + from_bits<T>(r.w));
                }
                const T ws = warp_reduce(fs);
                if (lane == 0) s_red[warp] = ws;
            }
            __syncthreads();
            if (warp == 0) {
                T red = warp_reduce(lane < (uint32_t)NW ? s_red[lane] : T(0));
                if (fb == 0) red += acquire_carry<T>(p, lane);  // tile 0 is INCLUSIVE incl. carry
                uint32_t done = 0;
                if (lane == 0) {
                    const uint32_t st = fb ? ST_AGGREGATE : ST_INCLUSIVE;
                    const uint64_t mine = pack_desc(to_bits(red), p.epoch, st);
                    uint64_t seen = ld_relaxed_gpu(p.desc + fb);
                    T use = red;
                    uint32_t use_st = st;
                    bool inserted = false;
                    while (!desc_valid(seen, p.epoch)) {  // CAS from the stale word we observe to ours
                        const uint64_t old = atomicCAS(reinterpret_cast<unsigned long long*>(p.desc + fb),
                                                       (unsigned long long)seen, (unsigned long long)mine);
                        if (old == seen) {
                            inserted = true;
                            break;
                        }
                        seen = old;
                    }
                    if (!inserted) {  // the owner (or another rescuer) got there first
                        use = from_bits<T>(desc_value(seen));
                        use_st = desc_status(seen);
                    } else {
                        ++n_inserts;
                    }
                    prefix += use;
                    done = (use_st == ST_INCLUSIVE || fb == 0) ? 1u : 0u;
                }
                prefix = __shfl_sync(kFullMask, prefix, 0);
                done = __shfl_sync(kFullMask, done, 0);
                if (lane == 0) s_state = done ? LB_DONE : LB_FALLBACK;
            }
            __syncthreads();
            if (s_state == LB_DONE) break;
            __syncthreads();  // s_state is rewritten by the next look-back round
        }
    }

    if (warp == 0 && lane == 0) {
        const T inclusive = prefix + aggregate;
        if (tile != 0)
            atomicExch(reinterpret_cast<unsigned long long*>(p.desc + tile),
                       (unsigned long long)pack_desc(to_bits(inclusive), p.epoch, ST_INCLUSIVE));
        s_prefix = prefix;
        if (tile == p.num_tiles - 1) write_total<T>(p, inclusive);
        if (f.stats) {
            if (n_spins) atomicAdd(f.stats + 0, n_spins);
            if (n_inserts) atomicAdd(f.stats + 2, n_inserts);
            if (n_windows) atomicAdd(f.stats + 3, n_windows);
        }
    }
    if (tid == 0 && f.stats && n_fallbacks) atomicAdd(f.stats + 1, n_fallbacks);
    __syncthreads();

    Tile::template finish<INCLUSIVE>(s_data, v0, (s_prefix + warp_prefix) + thread_excl, p.out, tile, p.n, full, warp,
                                     lane);
}

}  // namespace gps
