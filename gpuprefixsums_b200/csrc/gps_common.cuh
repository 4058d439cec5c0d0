// gps_common.cuh -- device primitives shared by all sm_100a scan kernels.
//
// Replaces the reference's L0 layer (paths relative to /root/reference/GPUPrefixSumsCUDA):
//   Utils.cuh:22-71   lane id, InclusiveWarpScan, InclusiveWarpScanCircularShift,
//                     WarpReduceSum, uint4 helpers
//   ChainedScanDecoupledLookback.cuh:19-23,50-77   flag encoding + LookbackWarp
// Differences that matter on B200:
//   * tile descriptor is ONE 64-bit word {value:32 | epoch:30 | status:2}, so the aggregate is
//     a full 32-bit value (the reference packs 30 bits next to the flag and silently
//     truncates totals >= 2^30) and float32 bit patterns travel unchanged.  Because payload
//     and flag share the word, publish = atom.exch.relaxed.gpu and poll = ld.relaxed.gpu by
//     default (measured fastest, see desc_publish below); st.release.gpu / ld.acquire.gpu
//     stay selectable;
//   * the epoch tag makes descriptors of earlier launches read as NOT_READY, so no
//     cudaMemset per launch (reference: ...Dispatcher.cuh:172-175);
//   * polls are gpu-scope loads, not `volatile` (.SYS) loads, and no block-wide
//     __threadfence() is needed (the reference fences in every thread).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <type_traits>

namespace gps {

// ---------------------------------------------------------------------------------
// kernel parameter block (passed by value as a __grid_constant__)
// ---------------------------------------------------------------------------------
// Destinations of a tagged total {value:32 | tag:32}: carry slots of downstream
// consumers (peer GPUs over NVLink, or the next chunk of the host pipeline).
constexpr int kMaxPublish = 16;
struct PublishList {
    uint64_t* dst[kMaxPublish];
    uint32_t n;
    uint32_t tag;
};

struct ScanParams {
    const void* in;
    void* out;
    uint64_t n;               // keys
    uint64_t* desc;           // one 64-bit descriptor per tile
    uint32_t* counter;        // monotonically increasing tile-claim counter
    uint32_t counter_base;    // value of *counter when this launch starts
    uint32_t epoch;           // 1 .. 2^30-1, unique per launch within a ctx
    uint32_t num_tiles;
    uint32_t carry_bits;      // immediate carry (uint32 or float bits)
    const uint64_t* slots;    // optional carry slots {value:32 | tag:32}
    uint32_t n_slots;         // slots polled until their tag == slot_tag
    uint32_t n_carry;         // the first n_carry of them are summed into the carry
    uint32_t slot_tag;
    uint32_t slot_poll_limit; // polls a carry slot may take (0 = kSlotPollsDefault)
    void* total_out;          // optional: carry + sum of all keys
    PublishList publish;      // optional: tagged total for downstream consumers
    uint32_t* status;         // watchdog / error word (bit 0 = poll budget exceeded)
    uint32_t poll_sleep_ns;
    uint32_t flags;           // PF_* bits
    unsigned long long* prof; // optional (tuning): kProfSlots x kProfFields counters
    uint32_t debug_delay_tile; // test hook: the CTA of tile (value - 1) sleeps ~40 ms before it starts (0 = off)
};
// profiling counters (per tile, accumulated): SM cycles spent in each phase + look-back effort
enum : int { PROF_TILES = 0, PROF_CLAIM, PROF_LOAD, PROF_REDUCE, PROF_LOOKBACK, PROF_FINISH,
             PROF_WINDOWS, PROF_POLLS, kProfFields };
constexpr int kProfSlots = 64;  // spread over 64 rows to keep the atomics uncontended
// Descriptor protocol (measured on B200 at 2^28 keys, register path <256,4>):
//   publish st.release.gpu + poll ld.acquire.gpu : 281 Gkeys/s  (MEMBAR.ALL.GPU before every
//            store; CCTL.IVALL after every poll; the store itself lingers behind the SM's
//            streaming data stores before it becomes visible)
//   publish st.relaxed.gpu + poll ld.relaxed.gpu : 309 Gkeys/s
//   publish atom.exch.gpu  + poll ld.relaxed.gpu : 426 Gkeys/s  <- default
// The descriptor carries its payload in the same 64-bit word as the flag, so plain CSDL
// needs no ordering against other memory; the atomic is used because it is performed at L2
// immediately instead of waiting in the store path.  The fallback kernel adds the one
// fence it needs explicitly.
enum : uint32_t {
    PF_PUBLISH_ST_RELEASE = 1u,  // publish with st.release.gpu (else atom.exch.relaxed.gpu)
    PF_SKIP_LOOKBACK = 2u,       // experiment: publish but never look back (WRONG results)
    PF_FENCE_ALL = 4u,           // experiment: every thread fences after the publish (reference style)
    PF_LD_ALLOCATE = 8u,         // experiment: plain (L1-allocating) loads/stores
    PF_STATIC_TILE = 16u,        // tile = blockIdx.x (relies on in-order CTA dispatch, like CUB)
    PF_POLL_VOLATILE = 32u,      // poll with ld.volatile (SASS LDG.STRONG.SYS, reference style)
    PF_PUBLISH_ST_RELAXED = 64u, // publish with st.relaxed.gpu
    PF_DELAY_FIRST = 128u,       // experiment: nanosleep(poll_sleep_ns) once before the first poll
    PF_POLL_ACQUIRE = 256u,      // poll with ld.acquire.gpu (else ld.relaxed.gpu)
    PF_LOOKBACK_WAIT_ALL = 512u, // experiment: consume a window only when all 32 are ready
    PF_LOOKBACK_SINGLE = 1024u,  // one-thread look-back (reference LookbackSingle, CSDL.cuh:28-47)
    PF_EAGER_LOOKBACK = 2048u,   // smem tile: warp 0 walks the look-back WHILE the tile is in flight (the other warps reduce it)
    PF_PDL = 4096u,              // host: launch with programmatic stream serialization (see pdl_entry)
    PF_PDL_FOLD = 32768u,        // host: the one-block fold of gps_reduce_publish is launched programmatically
    PF_L2_OUT_EVICT_FIRST = 65536u, // tile stores carry an L2 evict_first policy (the sharded scan: keep the input the
                                    // reduce pass left in L2 instead of the output nobody re-reads)
    PF_L2_IN_EVICT_FIRST = 8192u,   // tile loads carry an L2 evict_first policy (input is read once)
    PF_L2_OUT_EVICT_LAST = 16384u,  // tile stores carry an L2 evict_last policy (output stays for the consumer)
};

enum : uint32_t { ST_NOT_READY = 0, ST_AGGREGATE = 1, ST_INCLUSIVE = 2, ST_MASK = 3 };
// STATUS_RESCUED is informational (a one-wave ticket-free scan rescued a tile whose CTA was late, see rescue_tile_warp)
enum : uint32_t { STATUS_WATCHDOG = 1u, STATUS_SLOT_TIMEOUT = 2u, STATUS_RESCUED = 0x100u };
constexpr uint32_t kRescuePolls = 1u << 14;  // unsuccessful polls of the nearest tile (~10 ms) before a rescue
constexpr uint32_t kFullMask = 0xFFFFFFFFu;
// CTAs per SM the scan kernels are compiled for: as many as fit at `regs` registers/thread
constexpr int min_ctas_per_sm(int nt, int regs = 40) { return 65536 / (nt * regs) > 0 ? 65536 / (nt * regs) : 1; }
#ifndef GPS_LOOKBACK_MAXWIN
#define GPS_LOOKBACK_MAXWIN 1  // windows per round trip a far-behind look-back may grow to (1 = never grow);
                               // measured: growth LOSES at every size (2^24 keys: 27.1 / 27.6 / 30.3 / 34.6 us for 1 / 2 / 4 / 8)
#endif
constexpr uint32_t kWatchdogPolls = 1u << 22;  // * ~0.5 us per poll => seconds, then bail out

// ---------------------------------------------------------------------------------
// bit casts
// ---------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ uint32_t to_bits(T v)
{
    if constexpr (std::is_same<T, float>::value)
        return __float_as_uint(v);
    else
        return (uint32_t)v;
}
template <typename T>
__device__ __forceinline__ T from_bits(uint32_t b)
{
    if constexpr (std::is_same<T, float>::value)
        return __uint_as_float(b);
    else
        return (T)b;
}

__device__ __forceinline__ uint64_t pack_desc(uint32_t value_bits, uint32_t epoch, uint32_t status)
{
    return ((uint64_t)((epoch << 2) | status) << 32) | (uint64_t)value_bits;
}
__device__ __forceinline__ uint32_t desc_value(uint64_t d) { return (uint32_t)d; }
__device__ __forceinline__ uint32_t desc_status(uint64_t d) { return (uint32_t)(d >> 32) & ST_MASK; }
__device__ __forceinline__ bool desc_valid(uint64_t d, uint32_t epoch)
{
    return ((uint32_t)(d >> 34)) == epoch;
}

// ---------------------------------------------------------------------------------
// PTX: scoped loads / stores for descriptors and carry slots
// ---------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t ld_acquire_gpu(const uint64_t* p)
{
    uint64_t v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ uint64_t ld_relaxed_gpu(const uint64_t* p)
{
    uint64_t v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_gpu(uint64_t* p, uint64_t v)
{
    asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void st_relaxed_gpu(uint64_t* p, uint64_t v)
{
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
// system-scope relaxed: the carry slots.  A slot word {value | tag} is self-contained -- nothing
// else has to become visible with it -- so neither side needs release / acquire ordering; what
// matters is that the access is STRONG at system scope (performed at the point of coherence, never
// served from a stale line).  st.release.sys costs a system-wide fence per store, and a rank
// publishes to every peer one after the other: with relaxed stores the G stores are in flight
// together.
__device__ __forceinline__ uint64_t ld_relaxed_sys(const uint64_t* p)
{
    uint64_t v;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_sys(uint64_t* p, uint64_t v)
{
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_relaxed_gpu_u32(const uint32_t* p)
{
    uint32_t v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// Descriptor publish / poll (protocol selected by warp-uniform flags, see PF_* above).
__device__ __forceinline__ void desc_publish(const ScanParams& p, uint64_t* ptr, uint64_t v)
{
    if (p.flags & PF_PUBLISH_ST_RELEASE)
        st_release_gpu(ptr, v);
    else if (p.flags & PF_PUBLISH_ST_RELAXED)
        st_relaxed_gpu(ptr, v);
    else
        asm volatile("atom.relaxed.gpu.global.exch.b64 %0, [%1], %2;" : "=l"(v) : "l"(ptr), "l"(v) : "memory");
}
__device__ __forceinline__ uint64_t desc_poll(const ScanParams& p, const uint64_t* ptr)
{
    if (p.flags & PF_POLL_VOLATILE) return *reinterpret_cast<const volatile uint64_t*>(ptr);
    return (p.flags & PF_POLL_ACQUIRE) ? ld_acquire_gpu(ptr) : ld_relaxed_gpu(ptr);
}

// ---------------------------------------------------------------------------------
// PTX: streaming 128-bit global access (data is touched exactly once: keep it out of
// L1).  Not .nc: the scan may run in place.
// ---------------------------------------------------------------------------------
__device__ __forceinline__ uint4 ldg_stream(const uint4* p)
{
    uint4 r;
    asm volatile("ld.global.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p)
                 : "memory");
    return r;
}
__device__ __forceinline__ void stg_stream(uint4* p, const uint4& v)
{
    asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x),
                 "r"(v.y), "r"(v.z), "r"(v.w)
                 : "memory");
}

// ---------------------------------------------------------------------------------
// PTX: mbarrier + TMA bulk copies (cp.async.bulk -> SASS UBLKCP)
// ---------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_addr(const void* p)
{
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count)
                 : "memory");
}
__device__ __forceinline__ void fence_mbar_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_addr(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_addr(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
    while (!mbar_try_wait(bar, parity)) {
    }
}
// bounded variant: false if the phase did not complete within ~2^22 try_waits (each of
// which suspends the thread for a HW-defined interval) -- a TMA that never lands must not
// hang the GPU
__device__ __forceinline__ bool mbar_wait_bounded(uint64_t* bar, uint32_t parity)
{
    for (uint32_t i = 0; i < (1u << 22); ++i)
        if (mbar_try_wait(bar, parity)) return true;
    return false;
}
// named barrier over a subset of the CTA's warps (id 1..15; `threads` a multiple of 32)
__device__ __forceinline__ void named_bar_sync(uint32_t id, uint32_t threads)
{
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}
// global -> shared::cta, completion signalled on an mbarrier (bytes % 16 == 0)
__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gmem_src, uint32_t bytes,
                                          uint64_t* bar)
{
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
            "r"(smem_addr(smem_dst)),
        "l"(gmem_src), "r"(bytes), "r"(smem_addr(bar))
        : "memory");
}
// shared::cta -> global, tracked by the per-thread bulk async-group
__device__ __forceinline__ void bulk_store(void* gmem_dst, const void* smem_src, uint32_t bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem_dst),
                 "r"(smem_addr(smem_src)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read()
{
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
template <int N>
__device__ __forceinline__ void bulk_wait()
{
    asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}
// ---- L2 cache policies (createpolicy) for inputs whose output fits the 126 MB L2 ------------
// A scan reads its input exactly once and produces an output somebody reads next.  For mid-size
// inputs (output <= ~L2) the loads are tagged evict_first and the stores evict_last: the input
// streams through without pushing the freshly written output out of L2, so the output is never
// written to DRAM while it is still wanted -- by the consumer, or by the next scan that overwrites
// it.  Measured: DESIGN.md (mid-size regime).
__device__ __forceinline__ uint64_t l2_policy_evict_first()
{
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ uint64_t l2_policy_evict_last()
{
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ void bulk_load_hint(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar,
                                               uint64_t policy)
{
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::
            "r"(smem_addr(smem_dst)),
        "l"(gmem_src), "r"(bytes), "r"(smem_addr(bar)), "l"(policy)
        : "memory");
}
__device__ __forceinline__ void bulk_store_hint(void* gmem_dst, const void* smem_src, uint32_t bytes, uint64_t policy)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(gmem_dst),
                 "r"(smem_addr(smem_src)), "r"(bytes), "l"(policy)
                 : "memory");
}
__device__ __forceinline__ void stg_stream_hint(uint4* p, const uint4& v, uint64_t policy)
{
    asm volatile("st.global.L1::no_allocate.L2::cache_hint.v4.u32 [%0], {%1,%2,%3,%4}, %5;" ::"l"(p), "r"(v.x),
                 "r"(v.y), "r"(v.z), "r"(v.w), "l"(policy)
                 : "memory");
}
// tile load / store that honour the PF_L2_* flags (warp-uniform)
__device__ __forceinline__ void tile_bulk_load(uint32_t flags, void* smem_dst, const void* gmem_src, uint32_t bytes,
                                               uint64_t* bar)
{
    if (flags & PF_L2_IN_EVICT_FIRST)
        bulk_load_hint(smem_dst, gmem_src, bytes, bar, l2_policy_evict_first());
    else
        bulk_load(smem_dst, gmem_src, bytes, bar);
}
__device__ __forceinline__ void tile_bulk_store(uint32_t flags, void* gmem_dst, const void* smem_src, uint32_t bytes)
{
    if (flags & PF_L2_OUT_EVICT_LAST)
        bulk_store_hint(gmem_dst, smem_src, bytes, l2_policy_evict_last());
    else if (flags & PF_L2_OUT_EVICT_FIRST)
        bulk_store_hint(gmem_dst, smem_src, bytes, l2_policy_evict_first());
    else
        bulk_store(gmem_dst, smem_src, bytes);
}

// make generic-proxy shared-memory writes visible to the async proxy (TMA store)
__device__ __forceinline__ void fence_proxy_async_smem()
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// ---------------------------------------------------------------------------------
// Programmatic dependent launch: a kernel launched with the programmatic-stream-serialization
// attribute may have its CTAs made resident while the previous kernel of the stream is still
// draining.  Every scan kernel starts with pdl_entry(): it lets ITS successor do the same as
// soon as all of this grid's CTAs have started, then waits until the predecessor grid has
// completed and its memory is visible -- so nothing a scan reads or claims (input, ticket
// counter, descriptors) is touched early; what overlaps is the launch latency and the CTA
// ramp-up.  Both instructions are no-ops for a normally launched kernel.
// ---------------------------------------------------------------------------------
__device__ __forceinline__ void pdl_entry()
{
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
}

// ---------------------------------------------------------------------------------
// warp primitives (Utils.cuh:22-59), generic over uint32_t / float
// ---------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t lane_id()
{
    uint32_t l;
    asm("mov.u32 %0, %%laneid;" : "=r"(l));
    return l;
}

// One Kogge-Stone step, branch-free: v += (value of lane - d) for lanes >= d.  The add is
// predicated on the shuffle's own in-range predicate in PTX.  (Written as
// `t = __shfl_up_sync(..); if (lane >= d) v += t;` nvcc 12.9 turns the condition into a
// BRANCH around the shuffles and pays WARPSYNC.COLLECTIVE / BRA.DIV sequences for every
// one of them: 536 SASS instructions per warp on the tile path instead of ~200.)
__device__ __forceinline__ uint32_t scan_step_up(uint32_t v, int d)
{
    uint32_t r;
    asm volatile(
        "{\n\t.reg .u32 t;\n\t.reg .pred q;\n\t"
        "shfl.sync.up.b32 t|q, %1, %2, 0, 0xffffffff;\n\t"
        "@q add.u32 t, t, %1;\n\t"
        "mov.u32 %0, t;\n\t}"
        : "=r"(r)
        : "r"(v), "r"(d));
    return r;
}
__device__ __forceinline__ float scan_step_up(float v, int d)
{
    float r;
    asm volatile(
        "{\n\t.reg .f32 t;\n\t.reg .pred q;\n\t"
        "shfl.sync.up.b32 t|q, %1, %2, 0, 0xffffffff;\n\t"
        "@q add.f32 t, t, %1;\n\t"
        "mov.f32 %0, t;\n\t}"
        : "=f"(r)
        : "f"(v), "r"(d));
    return r;
}

template <typename T>
__device__ __forceinline__ T warp_inclusive_scan(T v, uint32_t lane)
{
    (void)lane;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) v = scan_step_up(v, d);
    return v;
}

// exclusive prefix of a lane given the inclusive scan `inc` of `raw`
template <typename T>
__device__ __forceinline__ T warp_exclusive_of(T inc, T raw, uint32_t lane)
{
    if constexpr (std::is_same<T, float>::value) {
        const T t = __shfl_up_sync(kFullMask, inc, 1);
        return lane ? t : T(0);
    } else {
        (void)lane;
        return inc - raw;  // exact mod 2^32
    }
}

// warp sum, result in every lane.  uint32: one REDUX.SUM (sm_80+); float: xor butterfly
template <typename T>
__device__ __forceinline__ T warp_reduce(T v)
{
    if constexpr (std::is_same<T, uint32_t>::value) {
        return __reduce_add_sync(kFullMask, v);
    } else {
#pragma unroll
        for (int m = 16; m; m >>= 1) v += __shfl_xor_sync(kFullMask, v, m);
        return v;
    }
}

// Block-level exclusive scan of the per-warp totals held in s_warp[NW] (NW <= 32):
// every warp redundantly scans the NW totals with shuffles.  Must be called by all
// threads of the warp after the __syncthreads that follows the s_warp writes.
template <typename T, int NW>
__device__ __forceinline__ void block_scan_warp_totals(const T* s_warp, uint32_t warp,
                                                       uint32_t lane, T& warp_prefix, T& aggregate)
{
    static_assert(NW <= 32, "at most 32 warps per block");
    const T raw = lane < (uint32_t)NW ? s_warp[lane] : T(0);
    const T inc = warp_inclusive_scan(raw, lane);
    const T exc = warp_exclusive_of(inc, raw, lane);
    warp_prefix = __shfl_sync(kFullMask, exc, (int)warp);
    aggregate = __shfl_sync(kFullMask, inc, NW - 1);
}

// ---------------------------------------------------------------------------------
// watchdog: a spin that exceeds kWatchdogPolls sets bit 0 of *status and gives up
// (the result is then wrong, gps_check() reports GPS_ERR_WATCHDOG) instead of
// hanging the GPU.  Pollers also bail out early once another tile tripped it.
// ---------------------------------------------------------------------------------
__device__ __forceinline__ bool watchdog_tick(uint32_t& polls, uint32_t* status)
{
    ++polls;
    if ((polls & 1023u) == 0u) {
        if (polls >= kWatchdogPolls) {
            atomicOr(status, STATUS_WATCHDOG);
            return true;
        }
        if (ld_relaxed_gpu_u32(status) & STATUS_WATCHDOG) return true;
    }
    return false;
}

// ---------------------------------------------------------------------------------
// carry acquisition (tile 0 / spine): immediate carry + sum of the tagged carry slots.
// Called by a full warp.  Slots may live in this GPU's memory and be written by peer
// GPUs over NVLink (st.relaxed.sys), hence system scope.
// ---------------------------------------------------------------------------------
// The wait for a slot has its OWN poll budget, far larger than the look-back watchdog: the writer
// is another GPU / process / pipeline stage that may legitimately be seconds late (lazy module
// load on its first call, a slower copy-in, a paused process).  After the first 1024 polls the
// back-off grows to 1 us, so the default budget of 2^26 polls is more than a minute.  A slot that
// never arrives sets STATUS_SLOT_TIMEOUT (reported by gps_check / gps_sharded_* as
// GPS_ERR_WATCHDOG) and contributes NOTHING -- never the mismatched word.
constexpr uint32_t kSlotPollsDefault = 1u << 26;
template <typename T>
__device__ __forceinline__ T acquire_carry(const ScanParams& p, uint32_t lane)
{
    T carry = from_bits<T>(p.carry_bits);
    const uint32_t budget = p.slot_poll_limit ? p.slot_poll_limit : kSlotPollsDefault;
    for (uint32_t base = 0; base < p.n_slots; base += 32) {
        const uint32_t i = base + lane;
        T v = T(0);
        if (i < p.n_slots) {
            uint32_t polls = 0;
            uint64_t s = ld_relaxed_sys(p.slots + i);
            bool ok = true;
            while ((uint32_t)(s >> 32) != p.slot_tag) {
                if (++polls >= budget) {
                    // its own bit only: the look-back pollers bail out early on STATUS_WATCHDOG (watchdog_tick), but a
                    // slot that timed out is about to be published with a DEFINED value (it contributes nothing) --
                    // they must keep waiting for it, or the reported-but-defined result would depend on timing
                    atomicOr(p.status, STATUS_SLOT_TIMEOUT);
                    ok = false;
                    break;
                }
                __nanosleep(polls < 1024u ? 64u : 1000u);
                s = ld_relaxed_sys(p.slots + i);
            }
            if (ok && i < p.n_carry) v = from_bits<T>((uint32_t)s);
        }
        carry += warp_reduce(v);
    }
    return carry;
}

// tagged total -> every destination slot (called by ONE thread)
template <typename T>
__device__ __forceinline__ void publish_total(const PublishList& pl, T total)
{
    const uint64_t word = ((uint64_t)pl.tag << 32) | (uint64_t)to_bits(total);
#pragma unroll 1
    for (uint32_t i = 0; i < pl.n; ++i) st_relaxed_sys(pl.dst[i], word);  // all in flight together
}

// grand total of a launch: plain value + tagged copies for downstream consumers
template <typename T>
__device__ __forceinline__ void write_total(const ScanParams& p, T total)
{
    if (p.total_out) *reinterpret_cast<T*>(p.total_out) = total;
    publish_total<T>(p.publish, total);
}

// claim the next tile index; the counter is never reset (host advances counter_base)
__device__ __forceinline__ uint32_t claim_tile(const ScanParams& p)
{
    return atomicAdd(p.counter, 1u) - p.counter_base;
}

// ---------------------------------------------------------------------------------
// warp-parallel decoupled look-back (ChainedScanDecoupledLookback.cuh:50-77).
// Lane l polls the descriptor of tile (top - l); lanes that fall below tile 0 see a
// virtual INCLUSIVE/0 entry (they are never the nearest INCLUSIVE because tile 0
// itself always publishes INCLUSIVE).  Returns the exclusive prefix of `tile`.
// ---------------------------------------------------------------------------------
// Warp-level Decoupled Fallback for the plain kernels that run ticket-free (tile = blockIdx.x, one-wave grids): when the
// nearest predecessor has not published after kRescuePolls polls -- its CTA is not resident yet, which in-order dispatch
// never lets happen -- the polling warp re-reduces that tile from the INPUT (full by construction: a blocker is never the
// last tile) and inserts its aggregate with CAS, exactly what the Decoupled Fallback kernels do block-wide
// (EmulatedDeadlocking.cuh:103-151 of the reference).  The walk then continues on what it finds there -- its own word or
// the owner's.  Cold path only: a separate function, called through the look-back's OnStuck hook after the caller has parked
// its tile's registers in shared memory (RegTile::park), so that the hot path keeps its register allocation; out of place only
// (in place the owner's stores could race with the re-read; those launches keep the give-up watchdog).  With it the
// one-wave ticket-free scans need neither in-order dispatch nor anything of the kernels they share the device with.
template <typename T>
__device__ __noinline__ void rescue_tile_warp(const ScanParams& p, uint32_t tile_keys, uint32_t b, uint32_t lane)
{
    const uint4* in4 = reinterpret_cast<const uint4*>(p.in) + (uint64_t)b * (tile_keys / 4u);
    T acc = T(0);
#pragma unroll 1
    for (uint32_t i = lane; i < tile_keys / 4u; i += 32u) {
        const uint4 r = ldg_stream(in4 + i);
        acc += (from_bits<T>(r.x) + from_bits<T>(r.y)) + (from_bits<T>(r.z) + from_bits<T>(r.w));
    }
    T red = warp_reduce(acc);
    if (b == 0u) red += acquire_carry<T>(p, lane);  // tile 0 is INCLUSIVE incl. the carry
    if (lane == 0u) {
        const uint64_t mine = pack_desc(to_bits(red), p.epoch, b ? ST_AGGREGATE : ST_INCLUSIVE);
        uint64_t seen = ld_relaxed_gpu(p.desc + b);
        while (!desc_valid(seen, p.epoch)) {  // CAS from the stale word we observe to ours; the owner may get there first
            const uint64_t old = atomicCAS(reinterpret_cast<unsigned long long*>(p.desc + b), (unsigned long long)seen,
                                           (unsigned long long)mine);
            if (old == seen) {
                atomicOr(p.status, STATUS_RESCUED);
                break;
            }
            seen = old;
        }
    }
    __syncwarp();
}

struct LookbackStats {
    uint32_t windows = 0;  // descriptor windows consumed
    uint32_t polls = 0;    // polls that found the nearest tile not ready
};

// WIN = descriptor windows (of 32 tiles) fetched per round trip: the WIN loads of a lane
// are issued back to back, so one L2 round trip covers 32*WIN predecessors.
// MAXWIN = how many windows per round trip a far-behind walk may grow to (see below); the kernels
// whose tile lives in registers keep MAXWIN = WIN (no spare registers for the extra descriptors)
// OnStuck: hook called (warp-uniformly, with the nearest unpublished tile) every kRescuePolls unsuccessful polls; the
// ticket-free register-tile kernels rescue that tile there (rescue_tile_warp), everybody else passes nothing
struct NoRescue {
    static constexpr bool enabled = false;
    __device__ __forceinline__ void operator()(uint32_t) const {}
};
template <typename T, int WIN = 2, int MAXWIN_ = 0, typename OnStuck = NoRescue>
__device__ __forceinline__ T lookback_warp(const ScanParams& p, uint32_t tile, uint32_t lane,
                                           LookbackStats* stats = nullptr, OnStuck on_stuck = OnStuck())
{
    T prefix = T(0);
    int64_t top = (int64_t)tile - 1;
    uint32_t polls = 0, windows = 0;
    if ((p.flags & PF_DELAY_FIRST) && p.poll_sleep_ns) __nanosleep(p.poll_sleep_ns);
    if (p.flags & PF_LOOKBACK_SINGLE) {
        // the reference's other look-back (ChainedScanDecoupledLookback.cuh:28-47, defined
        // there but not called): lane 0 walks one descriptor at a time; kept selectable to
        // measure it (gps_set_option "lookback" = 1).  The reference notes "negligible
        // performance difference" on its hardware; on B200 see DESIGN.md section 5.
        if (lane == 0) {
            while (top >= 0) {
                const uint64_t d = desc_poll(p, p.desc + top);
                if (!desc_valid(d, p.epoch)) {
                    if (watchdog_tick(polls, p.status)) break;
                    continue;
                }
                prefix += from_bits<T>(desc_value(d));
                ++windows;
                if (desc_status(d) == ST_INCLUSIVE) break;
                --top;
            }
        }
        prefix = __shfl_sync(kFullMask, prefix, 0);
        if (stats) {
            stats->windows = windows;
            stats->polls = polls;
        }
        return prefix;
    }
    const bool wait_all = p.flags & PF_LOOKBACK_WAIT_ALL;
    // Windows fetched per L2 round trip.  One window is the measured optimum (the nearest tiles
    // decide; more polls only add traffic).  A walk that consumed a whole round of AGGREGATEs without
    // meeting an INCLUSIVE MAY double its reach every round, up to MAXWIN windows (a round that
    // stops at a tile that is not ready falls back to WIN) -- an experiment aimed at the burst of
    // ~1800 tiles that start together at launch.  It lost at every size (profiles/
    // r02_midsize.txt), so MAXWIN defaults to WIN; the mid-size gap is the tail of the last wave,
    // not the depth of the walks.
    constexpr int MAXWIN = MAXWIN_ > WIN ? MAXWIN_ : WIN;
    int nwin = WIN;
    while (true) {
        uint64_t d[MAXWIN];
        bool ready[MAXWIN];
#pragma unroll
        for (int w = 0; w < MAXWIN; ++w) {
            const int64_t idx = top - 32 * w - (int64_t)lane;
            d[w] = pack_desc(0u, p.epoch, ST_INCLUSIVE);  // virtual entry below tile 0
            ready[w] = true;
            if (w < nwin && idx >= 0) {
                d[w] = desc_poll(p, p.desc + idx);
                ready[w] = desc_valid(d[w], p.epoch);
            }
        }
        // consume the run of ready descriptors nearest to this tile; an INCLUSIVE one in
        // that run ends the walk no matter what farther tiles are doing
        uint32_t consumed = 0;
        bool done = false, stopped = false;
#pragma unroll
        for (int w = 0; w < MAXWIN; ++w) {
            if (w >= nwin) break;
            const uint32_t not_ready = __ballot_sync(kFullMask, !ready[w]);
            const uint32_t run = not_ready ? (uint32_t)__ffs((int)not_ready) - 1u : 32u;
            const uint32_t in_run = run >= 32u ? kFullMask : ((1u << run) - 1u);
            const uint32_t incl =
                __ballot_sync(kFullMask, ready[w] && desc_status(d[w]) == ST_INCLUSIVE) & in_run;
            const T v = from_bits<T>(desc_value(d[w]));
            if (incl) {
                const uint32_t nearest = (uint32_t)__ffs((int)incl) - 1u;
                prefix += warp_reduce(lane <= nearest ? v : T(0));
                done = true;
                break;
            }
            if (run < 32u && wait_all) {  // reference behaviour: all 32 or nothing
                stopped = true;
                break;
            }
            if (run) {
                prefix += warp_reduce(lane < run ? v : T(0));
                consumed += run;
            }
            if (run < 32u) {
                stopped = true;
                break;
            }
        }
        windows += (consumed + 31u) / 32u + (done ? 1u : 0u);
        if (done) break;
        nwin = stopped ? WIN : (2 * nwin < MAXWIN ? 2 * nwin : MAXWIN);
        if (consumed == 0u) {
            const bool bail = watchdog_tick(polls, p.status);
            if (__any_sync(kFullMask, bail)) break;  // warp-uniform exit
            // (polls is warp-uniform: every lane ticks once per round)
            if (OnStuck::enabled && (polls & (kRescuePolls - 1u)) == 0u) on_stuck((uint32_t)top);
            if (p.poll_sleep_ns) __nanosleep(p.poll_sleep_ns);
            continue;
        }
        top -= (int64_t)consumed;
    }
    if (stats) {
        stats->windows = windows;
        stats->polls = polls;
    }
    return prefix;
}

// accumulate one tile's profile (called by ONE thread; p.prof != nullptr)
__device__ __forceinline__ void prof_add(const ScanParams& p, uint32_t tile, const long long (&t)[6],
                                         const LookbackStats& lb)
{
    unsigned long long* row = p.prof + (size_t)(tile % kProfSlots) * kProfFields;
    atomicAdd(row + PROF_TILES, 1ull);
    atomicAdd(row + PROF_CLAIM, (unsigned long long)(t[1] - t[0]));
    atomicAdd(row + PROF_LOAD, (unsigned long long)(t[2] - t[1]));
    atomicAdd(row + PROF_REDUCE, (unsigned long long)(t[3] - t[2]));
    atomicAdd(row + PROF_LOOKBACK, (unsigned long long)(t[4] - t[3]));
    atomicAdd(row + PROF_FINISH, (unsigned long long)(t[5] - t[4]));
    atomicAdd(row + PROF_WINDOWS, (unsigned long long)lb.windows);
    atomicAdd(row + PROF_POLLS, (unsigned long long)lb.polls);
}

}  // namespace gps
