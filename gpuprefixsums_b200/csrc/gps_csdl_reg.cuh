// gps_csdl_reg.cuh -- Chained Scan with Decoupled Lookback, register path, plus the
// Decoupled-Fallback / emulated-deadlock variant.
//
// Replaces ChainedScanDecoupledLookback.cuh:79-153 (CSDL, CSDLInclusive/Exclusive) and
// EmulatedDeadlocking.cuh:26-168 (EmulatedDeadlock) of the reference.  One CTA per
// tile; tile order = order of the atomic claim (CSDL.cuh:90-94), which is what makes
// spinning on predecessors deadlock-free.
#pragma once
#include "gps_tile_reg.cuh"

namespace gps {

// look-back hook of the ticket-free register-tile kernels: park the tile, rescue the late predecessor, fetch the tile back
template <typename T, typename Tile>
struct RegRescue {
    static constexpr bool enabled = true;
    const ScanParams& p;
    Tile& t;
    T* park;
    uint32_t lane;
    bool active;  // ticket-free launch, out of place
    __device__ __forceinline__ void operator()(uint32_t b) const
    {
        if (!active) return;
        t.park(park, lane);
        __syncwarp();
        rescue_tile_warp<T>(p, Tile::TILE_KEYS, b, lane);
        t.unpark(park, lane);
    }
};

// ---------------------------------------------------------------------------------
// CSDL, register path
// ---------------------------------------------------------------------------------
// WIN: look-back windows per round trip; PROF: compile the phase timers in (tuning only)
// REGS: register budget per thread (decides the CTAs per SM the kernel is compiled for)
// RESCUE: compile the warp-level Decoupled Fallback into the look-back's cold path (ticket-free one-wave launches that must
// not depend on anything; costs ~0.1 us per launch in code size, so the other instantiations leave it out)
template <typename T, int NT, int V, bool INCLUSIVE, int WIN = 2, bool PROF = false, int REGS = 24 + 4 * V, bool RESCUE = false>
__global__ void __launch_bounds__(NT, min_ctas_per_sm(NT, REGS)) csdl_reg_kernel(const __grid_constant__ ScanParams p)
{
    using Tile = RegTile<T, NT, V>;
    constexpr int NW = Tile::NW;
    __shared__ T s_warp[NW];
    __shared__ T s_prefix;
    __shared__ uint32_t s_tile;
    __shared__ T s_park[RESCUE ? 32 * Tile::PARK_WORDS : 1];  // warp 0's tile while it rescues a late predecessor (cold path)
    pdl_entry();

    const uint32_t lane = lane_id();
    const uint32_t warp = threadIdx.x >> 5;

    long long tprof[6] = {0, 0, 0, 0, 0, 0};
    LookbackStats lbs;
    const bool profiler = PROF && p.prof != nullptr && threadIdx.x == 0;
    if (profiler) tprof[0] = clock64();
    uint32_t tile;
    if (p.flags & PF_STATIC_TILE) {
        tile = blockIdx.x;
    } else {
        if (threadIdx.x == 0) s_tile = claim_tile(p);
        __syncthreads();
        tile = s_tile;
    }
    if (tile >= p.num_tiles) {  // ticket accounting out of step (an earlier launch failed): never index past the descriptors
        if (threadIdx.x == 0) atomicOr(p.status, STATUS_WATCHDOG);
        return;
    }
    if (p.debug_delay_tile && tile + 1u == p.debug_delay_tile) {  // test hook: a CTA that starts ~40 ms late
#pragma unroll 1
        for (int i = 0; i < 400; ++i) __nanosleep(100000u);
    }
    const bool full = ((uint64_t)tile + 1u) * Tile::TILE_KEYS <= p.n;
    const uint64_t vbase = Tile::vec_base(tile, warp, lane);

    if (profiler) tprof[1] = clock64();
    const bool allocate = p.flags & PF_LD_ALLOCATE;
    Tile t;
    if (full)
        t.load_full(p.in, vbase, allocate);
    else
        t.load_partial(p.in, vbase, p.n);
    // Aggregate first: every tile that is claimed but not yet published is a hole all later
    // tiles spin on, so the sum takes the shortest path -- thread sums, one REDUX per warp,
    // one barrier, one REDUX -- and is published BEFORE the V row scans the outputs need.
    const T wsum = warp_reduce(t.thread_sum());
    if (lane == 0) s_warp[warp] = wsum;
    __syncthreads();

    if (profiler) tprof[2] = clock64();
    const T wraw = lane < (uint32_t)NW ? s_warp[lane] : T(0);
    const T aggregate = warp_reduce(wraw);
    if (warp == 0 && lane == 0 && tile != 0)
        desc_publish(p, p.desc + tile, pack_desc(to_bits(aggregate), p.epoch, ST_AGGREGATE));
    if (profiler) tprof[3] = clock64();

    // scan work that does not need the tile prefix
    const T winc = warp_inclusive_scan(wraw, lane);
    const T warp_prefix = __shfl_sync(kFullMask, warp_exclusive_of(winc, wraw, lane), (int)warp);
    t.scan_warp(lane);

    if (warp == 0) {
        T prefix;
        if (tile == 0)
            prefix = acquire_carry<T>(p, lane);
        else if constexpr (RESCUE)
            // ticket-free launches (one-wave grids) rescue a late predecessor instead of waiting for it
            prefix = lookback_warp<T, WIN, 0, RegRescue<T, Tile>>(
                p, tile, lane, PROF ? &lbs : nullptr,
                RegRescue<T, Tile>{p, t, s_park, lane, (p.flags & PF_STATIC_TILE) && p.in != p.out});
        else
            prefix = (p.flags & PF_SKIP_LOOKBACK) ? T(0) : lookback_warp<T, WIN>(p, tile, lane, PROF ? &lbs : nullptr);
        if (profiler) tprof[4] = clock64();
        if (lane == 0) {
            const T inclusive = prefix + aggregate;
            desc_publish(p, p.desc + tile, pack_desc(to_bits(inclusive), p.epoch, ST_INCLUSIVE));
            s_prefix = prefix;
            if (tile == p.num_tiles - 1) write_total<T>(p, inclusive);
        }
    }
    if (p.flags & PF_FENCE_ALL) __threadfence();
    __syncthreads();

    const T prefix = s_prefix + warp_prefix;
    if (full)
        t.template store_full<INCLUSIVE>(p.out, vbase, prefix, allocate);
    else
        t.template store_partial<INCLUSIVE>(p.out, vbase, p.n, prefix);
    if (profiler) {
        tprof[5] = clock64();
        prof_add(p, tile, tprof, lbs);
    }
}

// ---------------------------------------------------------------------------------
// CSDL + Decoupled Fallback, with optional emulated stalls
// ---------------------------------------------------------------------------------
struct FallbackParams {
    uint32_t stall_mask;   // tiles with (tile & mask) == 0 stall; 0xFFFFFFFF = none
    uint32_t max_spin;     // unsuccessful polls of a window before falling back
    uint32_t* stats;       // nullable: {spins, fallbacks, insertions, lookback windows}
};
constexpr uint32_t kNoStall = 0xFFFFFFFFu;

enum : uint32_t { LB_DONE = 0, LB_FALLBACK = 1 };

// one tile of the chained scan with Decoupled Fallback (the whole pipeline after the claim)
template <typename T, int NT, int V, bool INCLUSIVE>
__device__ __forceinline__ void csdl_fallback_tile(const ScanParams& p, const FallbackParams& f, const uint32_t tile)
{
    using Tile = RegTile<T, NT, V>;
    constexpr int NW = Tile::NW;
    __shared__ T s_warp[NW];
    __shared__ T s_red[NW];
    __shared__ T s_prefix;
    __shared__ uint32_t s_state;     // LB_DONE / LB_FALLBACK
    __shared__ uint32_t s_fb_tile;   // tile to recompute

    const uint32_t lane = lane_id();
    const uint32_t warp = threadIdx.x >> 5;
    const bool full = ((uint64_t)tile + 1u) * Tile::TILE_KEYS <= p.n;
    const uint64_t vbase = Tile::vec_base(tile, warp, lane);

    Tile t;
    if (full)
        t.load_full(p.in, vbase);
    else
        t.load_partial(p.in, vbase, p.n);
    // aggregate first (see csdl_reg_kernel), row scans afterwards
    const T wsum = warp_reduce(t.thread_sum());
    if (lane == 0) s_warp[warp] = wsum;
    __syncthreads();
    const T wraw = lane < (uint32_t)NW ? s_warp[lane] : T(0);
    const T aggregate = warp_reduce(wraw);
    const T winc = warp_inclusive_scan(wraw, lane);
    const T warp_prefix = __shfl_sync(kFullMask, warp_exclusive_of(winc, wraw, lane), (int)warp);

    // Stall emulation (EmulatedDeadlocking.cuh:63-72): the tile refuses to publish until
    // some successor's fallback has published for it.  The last tile never stalls: it
    // has no successor to rescue it (the reference hangs there).
    const bool stalled = f.stall_mask != kNoStall && (tile & f.stall_mask) == 0u &&
                         p.num_tiles > 1u && tile != p.num_tiles - 1u;
    T carry = T(0);
    if (warp == 0) {
        if (tile == 0) carry = acquire_carry<T>(p, lane);
        if (lane == 0) {
            if (stalled) {
                uint32_t polls = 0;
                while (!desc_valid(ld_acquire_gpu(p.desc + tile), p.epoch)) {
                    if (watchdog_tick(polls, p.status)) break;
                }
            }
            // overwrite (same value for uint32; for float32 the recomputed aggregate may
            // differ in the last bits -- both are within the stated tolerance)
            const T v = tile ? aggregate : carry + aggregate;
            atomicExch(reinterpret_cast<unsigned long long*>(p.desc + tile),
                       (unsigned long long)pack_desc(to_bits(v), p.epoch, tile ? ST_AGGREGATE : ST_INCLUSIVE));
            // In-place safety of the fallback: a successor may be re-reading this tile's
            // INPUT right now.  The descriptor must be performed at L2 before any thread of
            // this CTA overwrites the input with outputs, so that successor's CAS then
            // fails and it discards what it read.  (fence + the __syncthreads below.)
            __threadfence();
        }
    }
    if (!stalled) t.scan_warp(lane);  // overlaps thread 0's publish
    __syncthreads();  // holds the whole block while stalled ("purely for emulation")
    if (stalled) t.scan_warp(lane);

    uint32_t n_spins = 0, n_fallbacks = 0, n_inserts = 0, n_windows = 0;
    T prefix = carry;  // tile 0: done
    if (tile != 0) {
        // warp 0 runs the warp-parallel look-back with a spin budget; when the budget is
        // exhausted the whole CTA recomputes the aggregate of the nearest not-ready tile
        int64_t top = (int64_t)tile - 1;
        prefix = T(0);
        while (true) {
            if (warp == 0) {
                uint32_t state = LB_DONE, spins = 0, polls = 0;
                uint32_t fb = 0;
                while (true) {
                    const int64_t idx = top - (int64_t)lane;
                    uint64_t d = pack_desc(0u, p.epoch, ST_INCLUSIVE);
                    bool ready = true;
                    if (idx >= 0) {
                        d = ld_acquire_gpu(p.desc + idx);
                        ready = desc_valid(d, p.epoch);
                    }
                    ++n_windows;
                    const uint32_t not_ready = __ballot_sync(kFullMask, !ready);
                    const uint32_t first_nr = not_ready ? (uint32_t)__ffs((int)not_ready) - 1u : 32u;
                    const uint32_t below = first_nr >= 32u ? kFullMask : ((1u << first_nr) - 1u);
                    const uint32_t incl =
                        __ballot_sync(kFullMask, ready && desc_status(d) == ST_INCLUSIVE) & below;
                    const T v = from_bits<T>(desc_value(d));
                    if (incl) {
                        const uint32_t nearest = (uint32_t)__ffs((int)incl) - 1u;
                        prefix += warp_reduce(lane <= nearest ? v : T(0));
                        state = LB_DONE;
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
                    if (__any_sync(kFullMask, bail)) {
                        state = LB_DONE;
                        break;
                    }
                    if (++spins < f.max_spin) {
                        if (p.poll_sleep_ns) __nanosleep(p.poll_sleep_ns);
                        continue;
                    }
                    // budget exhausted: consume the ready tiles nearer than the blocker and
                    // fall back on the blocker itself
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

            // ---- Decoupled Fallback (EmulatedDeadlocking.cuh:112-151): every thread
            // helps to re-reduce tile fb from the INPUT array
            const uint32_t fb = s_fb_tile;
            ++n_fallbacks;
            {
                Tile ft;  // a fallback target is never the last tile => always full
                ft.load_full(p.in, Tile::vec_base(fb, warp, lane));
                const T ws = warp_reduce(ft.thread_sum());
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
                    // CAS from the stale (earlier-epoch) word we observe to ours
                    uint64_t seen = ld_acquire_gpu(p.desc + fb);
                    T use = red;
                    uint32_t use_st = st;
                    bool inserted = false;
                    while (!desc_valid(seen, p.epoch)) {
                        const uint64_t old =
                            atomicCAS(reinterpret_cast<unsigned long long*>(p.desc + fb),
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
                // lane 0 carries the running prefix; broadcast so warp_reduce adds stay uniform
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
    if (threadIdx.x == 0 && f.stats && n_fallbacks) atomicAdd(f.stats + 1, n_fallbacks);
    __syncthreads();

    const T out_prefix = s_prefix + warp_prefix;
    if (full)
        t.template store_full<INCLUSIVE>(p.out, vbase, out_prefix);
    else
        t.template store_partial<INCLUSIVE>(p.out, vbase, p.n, out_prefix);
}

template <typename T, int NT, int V, bool INCLUSIVE>
__global__ void __launch_bounds__(NT, min_ctas_per_sm(NT, 32 + 4 * V))
    csdl_fallback_kernel(const __grid_constant__ ScanParams p, const FallbackParams f)
{
    pdl_entry();
    __shared__ uint32_t s_tile;
    uint32_t tile;
    if (p.flags & PF_STATIC_TILE) {
        // Decoupled Fallback needs no ticket: no CTA ever depends on another CTA making progress, so
        // tile = blockIdx.x is deadlock-free under any dispatch order
        tile = blockIdx.x;
    } else {
        if (threadIdx.x == 0) s_tile = claim_tile(p);
        __syncthreads();
        tile = s_tile;
    }
    if (tile >= p.num_tiles) {  // ticket accounting out of step (an earlier launch failed): never index past the descriptors
        if (threadIdx.x == 0) atomicOr(p.status, STATUS_WATCHDOG);
        return;
    }
    csdl_fallback_tile<T, NT, V, INCLUSIVE>(p, f, tile);
}

// Occupancy-estimation variant (GPUPrefixSumsWebGPUapis/SharedShaders/TestVariants/csdldf_occ.wgsl:
// 72-82): the same scan, but every CTA keeps claiming tiles until none are left, and a CTA counts
// itself once if its FIRST claim was a real tile.  With one CTA launched per tile, only the CTAs
// resident at the start ever get a tile (those launched later find the counter exhausted), so the
// count estimates how many CTAs of this kernel the device holds at once.  The result of the scan
// stays correct.  Claims made: num_tiles + gridDim.x.
template <typename T, int NT, int V, bool INCLUSIVE>
__global__ void __launch_bounds__(NT, min_ctas_per_sm(NT, 32 + 4 * V))
    csdl_fallback_occ_kernel(const __grid_constant__ ScanParams p, const FallbackParams f, uint32_t* occupancy)
{
    pdl_entry();
    __shared__ uint32_t s_tile;
    for (bool first = true;; first = false) {
        if (threadIdx.x == 0) s_tile = claim_tile(p);
        __syncthreads();
        const uint32_t tile = s_tile;
        if (tile >= p.num_tiles) break;
        if (first && threadIdx.x == 0) atomicAdd(occupancy, 1u);
        csdl_fallback_tile<T, NT, V, INCLUSIVE>(p, f, tile);
        __syncthreads();  // shared state of the tile pipeline is reused by the next tile
    }
}

}  // namespace gps
