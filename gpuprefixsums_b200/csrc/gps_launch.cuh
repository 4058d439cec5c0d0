// gps_launch.cuh -- host-side launch helpers shared by the library (gps_api.cu) and
// the tuning harness (tools/gps_tune.cu): scratch bookkeeping (epoch tags, claim
// counter base), grid arithmetic and the <<< >>> calls, templated on the kernel
// configuration.  Mirrors the private Dispatch* methods of the reference dispatchers
// (ChainedScanDecoupledLookbackDispatcher.cuh:166-195, ReduceThenScanDispatcher.cuh:158-187)
// minus the per-launch cudaMemsets.
#pragma once
#include <algorithm>

#include "gps_csdl_blocked.cuh"
#include "gps_csdl_reg.cuh"
#include "gps_csdl_smem.cuh"
#include "gps_csdl_tma.cuh"
#include "gps_rts.cuh"

namespace gps {

inline uint64_t div_up(uint64_t a, uint64_t b) { return (a + b - 1) / b; }

// <<< >>> or, with `pdl`, cudaLaunchKernelEx with the programmatic-stream-serialization
// attribute (the kernel must begin with pdl_entry(), gps_common.cuh)
template <typename... KArgs, typename... Args>
cudaError_t launch_scan_kernel(void (*kern)(KArgs...), uint32_t grid, uint32_t block, size_t smem, cudaStream_t st,
                               bool pdl, const Args&... args)
{
    if (!pdl) {
        kern<<<grid, block, smem, st>>>(args...);
        return cudaGetLastError();
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(block);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, args...);
}

// device scratch of one context + the host-side counters that replace memsets
struct Scratch {
    uint64_t* desc = nullptr;     // tile descriptors (zero-initialised once)
    uint32_t desc_tiles = 0;
    uint32_t* counter = nullptr;  // tile-claim counter, never reset
    uint32_t* status = nullptr;   // watchdog word
    void* partials = nullptr;     // one key-sized value per RTS tile
    uint32_t partials_capacity = 0;
    uint32_t counter_base = 0;
    uint32_t debug_delay_tile = 0;  // test hook (gps_set_option "debug_delay_tile"): tile + 1 whose CTA starts late
    uint32_t epoch = 0;
    int sm_count = 148;
    uint32_t poll_sleep_ns = 0;
    uint64_t reduce_keep_bytes = 0;  // gps_reduce_publish: front of the slice loaded evict_last (0 = no L2 policies)
    uint32_t slot_poll_limit = 0;  // 0 = kSlotPollsDefault (option "slot_wait_polls")
    uint32_t flags = 0;
    unsigned long long* prof = nullptr;  // optional profiling counters (tuning harness)
    // wide-element chained scan (gps_wide.cuh): 2*W words per tile, own epoch counter
    uint64_t* wdesc = nullptr;
    uint64_t wdesc_words = 0;
    uint32_t wepoch = 0;
};

// everything a scan launch needs besides the buffers
struct CarrySpec {
    uint32_t carry_bits = 0;
    const uint64_t* slots = nullptr;
    uint32_t n_slots = 0, n_carry = 0, tag = 0;
    void* total_out = nullptr;
    PublishList publish{};
};

inline uint32_t next_epoch(Scratch& s, cudaStream_t st)
{
    if (++s.epoch >= (1u << 30)) {  // wrap: epoch values would alias -> clear once
        cudaMemsetAsync(s.desc, 0, (size_t)s.desc_tiles * sizeof(uint64_t), st);
        s.epoch = 1;
    }
    return s.epoch;
}

inline ScanParams make_scan_params(Scratch& s, const void* in, void* out, uint64_t n,
                                   uint32_t tile_keys, const CarrySpec& cs, cudaStream_t st)
{
    ScanParams p{};
    p.in = in;
    p.out = out;
    p.n = n;
    p.desc = s.desc;
    p.counter = s.counter;
    p.counter_base = s.counter_base;
    p.epoch = next_epoch(s, st);
    p.num_tiles = (uint32_t)div_up(n, tile_keys);
    p.carry_bits = cs.carry_bits;
    p.slots = cs.slots;
    p.n_slots = cs.n_slots;
    p.n_carry = cs.n_carry;
    p.slot_tag = cs.tag;
    p.slot_poll_limit = s.slot_poll_limit;
    p.total_out = cs.total_out;
    p.publish = cs.publish;
    p.status = s.status;
    p.poll_sleep_ns = s.poll_sleep_ns;
    p.flags = s.flags;
    p.prof = s.prof;
    p.debug_delay_tile = s.debug_delay_tile;
    return p;
}

// ---- n == 0: only the carry travels ------------------------------------------------
template <typename T>
__global__ void carry_only_kernel(const __grid_constant__ ScanParams p)
{
    pdl_entry();
    const T c = acquire_carry<T>(p, lane_id());
    if (threadIdx.x == 0) write_total<T>(p, c);
}

template <typename T>
cudaError_t launch_carry_only(Scratch& s, const CarrySpec& cs, cudaStream_t st)
{
    ScanParams p = make_scan_params(s, nullptr, nullptr, 0, 4096, cs, st);
    return launch_scan_kernel(carry_only_kernel<T>, 1, 32, 0, st, s.flags & PF_PDL, p);
}

// ---- CSDL register path --------------------------------------------------------------
template <typename T, int NT, int V, bool INCL, int WIN = 2, bool PROF = false, int REGS = 24 + 4 * V, bool RESCUE = false>
cudaError_t launch_csdl_reg(Scratch& s, const void* in, void* out, uint64_t n, const CarrySpec& cs,
                            cudaStream_t st)
{
    using Tile = RegTile<T, NT, V>;
    ScanParams p = make_scan_params(s, in, out, n, Tile::TILE_KEYS, cs, st);
    const cudaError_t e =
        launch_scan_kernel(csdl_reg_kernel<T, NT, V, INCL, WIN, PROF, REGS, RESCUE>, p.num_tiles, NT, 0, st, s.flags & PF_PDL, p);
    // one claim per CTA -- of a launch that actually happened
    if (e == cudaSuccess && !(s.flags & PF_STATIC_TILE)) s.counter_base += p.num_tiles;
    return e;
}

// ---- CSDL shared-memory-resident tile (one TMA bulk load per CTA) ---------------------
// configured: in/out flag, false = set the shared-memory carve-out preference first
template <typename T, int NT, int V, bool INCL, int WIN = 1, int G = 4, int REGS = 40>
cudaError_t launch_csdl_smem(Scratch& s, const void* in, void* out, uint64_t n, const CarrySpec& cs,
                             bool& configured, cudaStream_t st)
{
    auto kern = csdl_smem_kernel<T, NT, V, INCL, WIN, G, REGS>;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout,
                                             (int)cudaSharedmemCarveoutMaxShared);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    ScanParams p = make_scan_params(s, in, out, n, (uint32_t)(NT * V * 4), cs, st);
    const cudaError_t e = launch_scan_kernel(kern, p.num_tiles, NT, 0, st, s.flags & PF_PDL, p);
    // one claim per CTA -- of a launch that actually happened
    if (e == cudaSuccess && !(s.flags & PF_STATIC_TILE)) s.counter_base += p.num_tiles;
    return e;
}

// persistent one-wave variant: `resident` = in/out cache of the CTAs the device holds at once (0 = ask the runtime)
template <typename T, int NT, int V, bool INCL, int WIN = 1, int G = 4, int REGS = 40>
cudaError_t launch_csdl_smem_persistent(Scratch& s, const void* in, void* out, uint64_t n, const CarrySpec& cs,
                                        int& resident, cudaStream_t st)
{
    auto kern = csdl_smem_persistent_kernel<T, NT, V, INCL, WIN, G, REGS>;
    if (resident == 0) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout,
                                             (int)cudaSharedmemCarveoutMaxShared);
        int per_sm = 0;
        if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, NT, 0);
        if (e != cudaSuccess) return e;
        resident = per_sm > 0 ? per_sm * s.sm_count : -1;
    }
    if (resident <= 0) return cudaErrorInvalidConfiguration;
    ScanParams p = make_scan_params(s, in, out, n, (uint32_t)(NT * V * 4), cs, st);
    const uint32_t grid = p.num_tiles < (uint32_t)resident ? p.num_tiles : (uint32_t)resident;
    return launch_scan_kernel(kern, grid, NT, 0, st, s.flags & PF_PDL, p);  // no tickets: counter_base stays
}

// ---- CSDL blocked shared-memory tile (TMA bulk load + per-warp TMA bulk store) --------// ---- CSDL blocked shared-memory tile (TMA bulk load + per-warp TMA bulk store) --------
template <typename T, int NT, int V, bool INCL, int WIN = 1, int REGS = 40>
cudaError_t launch_csdl_blocked(Scratch& s, const void* in, void* out, uint64_t n, const CarrySpec& cs,
                                bool& configured, cudaStream_t st)
{
    auto kern = csdl_blocked_kernel<T, NT, V, INCL, WIN, REGS>;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout,
                                             (int)cudaSharedmemCarveoutMaxShared);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    ScanParams p = make_scan_params(s, in, out, n, (uint32_t)(NT * V * 4), cs, st);
    const cudaError_t e = launch_scan_kernel(kern, p.num_tiles, NT, 0, st, s.flags & PF_PDL, p);
    // one claim per CTA -- of a launch that actually happened
    if (e == cudaSuccess && !(s.flags & PF_STATIC_TILE)) s.counter_base += p.num_tiles;
    return e;
}

// ---- CSDL TMA path -------------------------------------------------------------------
// ctas_per_sm: in/out cache (0 = query occupancy and opt in to the dynamic smem size);
// max_ctas_per_sm: 0 = as many as fit
template <typename T, int NT, int STAGES, bool INCL, int WIN = 2, bool PROF = false>
cudaError_t launch_csdl_tma(Scratch& s, const void* in, void* out, uint64_t n, const CarrySpec& cs,
                            int& ctas_per_sm, int max_ctas_per_sm, cudaStream_t st)
{
    using Cfg = TmaCfg<NT, STAGES>;
    auto kern = csdl_tma_kernel<T, NT, STAGES, INCL, WIN, PROF>;
    if (ctas_per_sm == 0) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)Cfg::RING_BYTES);
        if (e != cudaSuccess) return e;
        int occ = 0;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, NT, Cfg::RING_BYTES);
        if (e != cudaSuccess) return e;
        if (occ < 1) return cudaErrorLaunchOutOfResources;
        ctas_per_sm = occ;
    }
    int occ = ctas_per_sm;
    if (max_ctas_per_sm > 0) occ = std::min(occ, max_ctas_per_sm);
    ScanParams p = make_scan_params(s, in, out, n, Cfg::TILE_KEYS, cs, st);
    const uint32_t grid = std::min<uint32_t>(p.num_tiles, (uint32_t)(occ * s.sm_count));
    kern<<<grid, NT, Cfg::RING_BYTES, st>>>(p);
    const cudaError_t e = cudaGetLastError();
    // every CTA claims STAGES-1 tiles up front and one more per tile it processes
    if (e == cudaSuccess) s.counter_base += grid * (STAGES - 1) + p.num_tiles;
    return e;
}

// ---- CSDL + Decoupled Fallback ---------------------------------------------------------
template <typename T, int NT, int V, bool INCL>
cudaError_t launch_csdl_fallback(Scratch& s, const void* in, void* out, uint64_t n,
                                 const CarrySpec& cs, const FallbackParams& f, cudaStream_t st)
{
    using Tile = RegTile<T, NT, V>;
    ScanParams p = make_scan_params(s, in, out, n, Tile::TILE_KEYS, cs, st);
    const cudaError_t e = launch_scan_kernel(csdl_fallback_kernel<T, NT, V, INCL>, p.num_tiles, NT, 0, st, s.flags & PF_PDL, p, f);
    if (e == cudaSuccess && !(s.flags & PF_STATIC_TILE)) s.counter_base += p.num_tiles;
    return e;
}

// occupancy-estimation variant: one CTA per tile is launched, the resident ones drain the tickets
template <typename T, int NT, int V, bool INCL>
cudaError_t launch_csdl_fallback_occ(Scratch& s, const void* in, void* out, uint64_t n, const CarrySpec& cs,
                                     const FallbackParams& f, uint32_t* d_occupancy, cudaStream_t st)
{
    using Tile = RegTile<T, NT, V>;
    ScanParams p = make_scan_params(s, in, out, n, Tile::TILE_KEYS, cs, st);
    const cudaError_t e = launch_scan_kernel(csdl_fallback_occ_kernel<T, NT, V, INCL>, p.num_tiles, NT, 0, st, false, p, f,
                                             d_occupancy);
    if (e == cudaSuccess) s.counter_base += 2u * p.num_tiles;  // every CTA claims until it draws a ticket past the end
    return e;
}

// the same on the shared-memory-resident tile (gps_csdl_smem.cuh)
template <typename T, int NT, int V, bool INCL>
cudaError_t launch_csdl_smem_fallback(Scratch& s, const void* in, void* out, uint64_t n, const CarrySpec& cs,
                                      const FallbackParams& f, bool& configured, cudaStream_t st)
{
    auto kern = csdl_smem_fallback_kernel<T, NT, V, INCL>;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout,
                                             (int)cudaSharedmemCarveoutMaxShared);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    ScanParams p = make_scan_params(s, in, out, n, (uint32_t)(NT * V * 4), cs, st);
    const cudaError_t e = launch_scan_kernel(kern, p.num_tiles, NT, 0, st, s.flags & PF_PDL, p, f);
    if (e == cudaSuccess && !(s.flags & PF_STATIC_TILE)) s.counter_base += p.num_tiles;
    return e;
}

template <typename T, int NT, int V, bool INCL>
cudaError_t launch_csdl_blocked_fallback(Scratch& s, const void* in, void* out, uint64_t n, const CarrySpec& cs,
                                         const FallbackParams& f, bool& configured, cudaStream_t st)
{
    auto kern = csdl_blocked_fallback_kernel<T, NT, V, INCL>;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout,
                                             (int)cudaSharedmemCarveoutMaxShared);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    ScanParams p = make_scan_params(s, in, out, n, (uint32_t)(NT * V * 4), cs, st);
    const cudaError_t e = launch_scan_kernel(kern, p.num_tiles, NT, 0, st, s.flags & PF_PDL, p, f);
    if (e == cudaSuccess && !(s.flags & PF_STATIC_TILE)) s.counter_base += p.num_tiles;
    return e;
}

// ---- RTS -----------------------------------------------------------------------------
// s.partials must hold one key-sized value per tile (s.partials_capacity entries)
template <typename T, int NT, int V>
cudaError_t rts_params(Scratch& s, uint64_t n, RtsParams& p)
{
    using Tile = RegTile<T, NT, V>;
    p.n = n;
    p.num_tiles = (uint32_t)div_up(n, Tile::TILE_KEYS);
    p.partials = s.partials;
    p.status = s.status;
    return p.num_tiles <= s.partials_capacity ? cudaSuccess : cudaErrorInvalidValue;
}

template <typename T, int NT, int V, bool INCL>
cudaError_t launch_rts(Scratch& s, const void* in, void* out, uint64_t n, const CarrySpec& cs,
                       cudaStream_t st)
{
    RtsParams p{};
    cudaError_t e = rts_params<T, NT, V>(s, n, p);
    if (e != cudaSuccess) return e;
    p.in = in;
    p.out = out;
    p.carry_bits = cs.carry_bits;
    p.slots = cs.slots;
    p.n_slots = cs.n_slots;
    p.n_carry = cs.n_carry;
    p.slot_tag = cs.tag;
    p.slot_poll_limit = s.slot_poll_limit;
    p.total_out = cs.total_out;
    p.publish = cs.publish;
    const bool pdl = s.flags & PF_PDL;
    e = launch_scan_kernel(rts_reduce_kernel<T, NT, V>, p.num_tiles, NT, 0, st, pdl, p);
    if (e == cudaSuccess) e = launch_scan_kernel(rts_spine_kernel<T>, 1, kSpineThreads, 0, st, pdl, p);
    if (e == cudaSuccess) e = launch_scan_kernel(rts_propagate_kernel<T, NT, V, INCL>, p.num_tiles, NT, 0, st, pdl, p);
    return e;
}

// RTS with the propagate kernel on the blocked shared-memory tile
template <typename T, int NT, int V, bool INCL>
cudaError_t launch_rts_blocked(Scratch& s, const void* in, void* out, uint64_t n, const CarrySpec& cs, bool& configured,
                               cudaStream_t st)
{
    auto kern = rts_propagate_blocked_kernel<T, NT, V, INCL>;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout,
                                             (int)cudaSharedmemCarveoutMaxShared);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    RtsParams p{};
    cudaError_t e = rts_params<T, NT, V>(s, n, p);
    if (e != cudaSuccess) return e;
    p.in = in;
    p.out = out;
    p.carry_bits = cs.carry_bits;
    p.slots = cs.slots;
    p.n_slots = cs.n_slots;
    p.n_carry = cs.n_carry;
    p.slot_tag = cs.tag;
    p.slot_poll_limit = s.slot_poll_limit;
    p.total_out = cs.total_out;
    p.publish = cs.publish;
    // plain launches: this variant serves the large inputs, where programmatic dependent launch
    // costs 1.6 % (2^28 keys: 554 / 551 Gkeys/s without, 546 / 542 with -- tools/rts_ab.py), as it
    // does for the sharded reduce + scan pair
    const bool pdl = false;
    e = launch_scan_kernel(rts_reduce_kernel<T, NT, V>, p.num_tiles, NT, 0, st, pdl, p);
    if (e == cudaSuccess) e = launch_scan_kernel(rts_spine_kernel<T>, 1, kSpineThreads, 0, st, pdl, p);
    if (e == cudaSuccess) e = launch_scan_kernel(kern, p.num_tiles, NT, 0, st, pdl, p);
    return e;
}

template <typename T, int NT, int V>
cudaError_t launch_reduce(Scratch& s, const void* in, uint64_t n, void* total_out,
                          const PublishList& pl, cudaStream_t st)
{
    RtsParams p{};
    cudaError_t e = rts_params<T, NT, V>(s, n, p);
    if (e != cudaSuccess) return e;
    p.in = in;
    p.total_out = total_out;
    p.publish = pl;
    p.keep_tiles = s.reduce_keep_bytes ? (uint32_t)std::min<uint64_t>(p.num_tiles, s.reduce_keep_bytes / (RegTile<T, NT, V>::TILE_KEYS * 4)) : 0u;
    const bool pdl = s.flags & PF_PDL;
    if (p.num_tiles) e = launch_scan_kernel(rts_reduce_kernel<T, NT, V>, p.num_tiles, NT, 0, st, pdl, p);
    // the one-block fold may be launched programmatically even where the big kernels are not
    // (PF_PDL_FOLD, the sharded step): its launch latency then hides behind the reduce's last wave
    if (e == cudaSuccess)
        e = launch_scan_kernel(rts_fold_kernel<T>, 1, kSpineThreads, 0, st, pdl || (s.flags & PF_PDL_FOLD), p);
    return e;
}

}  // namespace gps
