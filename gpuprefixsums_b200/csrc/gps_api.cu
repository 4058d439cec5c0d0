// gps_api.cu -- the C ABI of libgps_b200.so (include/gps_b200.h): context, dispatch,
// harness helpers, host-buffer pipeline, IPC, single-process sharded scan.
//
// Host-side mirror of the reference's private dispatch methods
// (ChainedScanDecoupledLookbackDispatcher.cuh:166-211, ReduceThenScanDispatcher.cuh:158-204,
// EmulatedDeadlockingDispatcher.cuh:140-174): grid arithmetic, scratch management,
// kernel launch, validation -- but over caller-owned buffers, with checked errors and
// without per-launch scratch clears.  There is no CPU path in this file.
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <vector>

#include "../../include/gps_b200.h"
#include "gps_config.h"
#include "gps_launch.cuh"
#include "gps_util_kernels.cuh"
#include "gps_sort.cuh"
#include "gps_wide.cuh"

using namespace gps;

// ---------------------------------------------------------------------------------
// error plumbing
// ---------------------------------------------------------------------------------
namespace {
thread_local char g_err[512] = "";

int fail(int code, const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

#define GPS_CUDA(call)                                                                   \
    do {                                                                                 \
        cudaError_t e_ = (call);                                                         \
        if (e_ != cudaSuccess)                                                           \
            return fail(GPS_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), \
                        __FILE__, __LINE__);                                             \
    } while (0)

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

constexpr uint32_t kMinTileKeys = 4096;  // smallest tile of any compiled CSDL variant
constexpr uint32_t kHostChunkKeysDefault = 1u << 24;  // 64 MiB per pipeline chunk
constexpr int kHostBufs = 3;
static_assert(GPS_REG_NT * GPS_REG_V * 4 >= (int)kMinTileKeys, "descriptor array sized for 4096-key tiles");
static_assert(GPS_TMA_NT * 16 >= (int)kMinTileKeys, "descriptor array sized for 4096-key tiles");
static_assert(GPS_BLOCKED_NT * GPS_BLOCKED_V * 4 >= (int)kMinTileKeys, "descriptor array sized for 4096-key tiles");
static_assert(GPS_SMEM_NT * GPS_SMEM_V * 4 >= (int)kMinTileKeys, "descriptor array sized for 4096-key tiles");
static_assert(GPS_RTS_NT * GPS_RTS_V * 4 >= (int)kMinTileKeys, "partials array sized for 4096-key tiles");
}  // namespace

struct gps_ctx {
    int device = 0;
    uint64_t max_n = 0;
    Scratch s;                      // descriptors, claim counter, status word, partials
    void* scratch_base = nullptr;   // one allocation behind all of s.* (the ctx's or the caller's)
    bool owns_scratch = true;
    // options
    int csdl_variant = 0;           // 0 auto, 1 reg, 2 tma
    uint64_t host_chunk_keys = kHostChunkKeysDefault;
    // cached launch geometry
    int tma_ctas_per_sm[2][2] = {{0, 0}, {0, 0}};  // [dtype][kind]
    bool smem_configured[2][2] = {{false, false}, {false, false}};
    bool blocked_configured[2][2] = {{false, false}, {false, false}};
    bool smem_fb_configured[2][2] = {{false, false}, {false, false}};
    int reg_resident[2][2][3] = {};  // [dtype][kind][small / tiny / mid shape]: CTAs of the register-tile kernels the device holds at once (0 = not asked yet)
    bool blocked_fb_configured[2][2] = {{false, false}, {false, false}};
    bool rts_blocked_configured[2][2] = {{false, false}, {false, false}};
    int fallback_variant = 0;       // 0 auto, 1 register tile, 2 smem-resident tile, 3 blocked tile
    int ticket_free = 1;            // GPS_ALGO_CSDL: ticket-free Decoupled Fallback where it is faster (gps_config.h)
    int rescued = 0;                // gps_check saw STATUS_RESCUED since the last gps_get_option("rescued")
    int rts_variant = 0;            // 0 auto, 1 register-tile propagate, 2 blocked-tile propagate
    int sharded_pdl = GPS_SHARDED_PDL_DEFAULT;  // see NoPdlScope
    int sharded_l2 = GPS_SHARDED_L2_DEFAULT;    // option "sharded_l2"
    int sharded_l2_keep_mb = GPS_SHARDED_L2_KEEP_MB;
    int l2_hints = -1;              // -1 auto, else PF_L2_* bits (bit 0 loads evict_first, bit 1 stores evict_last)
    uint64_t l2_bytes = 0;          // L2 size of the device
    bool wide_configured[5][2][3] = {};             // [dtype][kind][chained | rts | fallback]
    // host pipeline (lazy)
    cudaStream_t s_h2d = nullptr, s_cmp = nullptr, s_d2h = nullptr;
    void* stage[kHostBufs] = {nullptr, nullptr, nullptr};
    cudaEvent_t ev_h2d[kHostBufs] = {}, ev_cmp[kHostBufs] = {}, ev_d2h[kHostBufs] = {};
    uint64_t* chain_slots = nullptr;  // 2 tagged slots chaining chunk totals
    uint32_t chain_tag = 0;
    void* chain_total = nullptr;
    cudaEvent_t ev_total = nullptr;
    bool ev_total_pending = false;
    // shadow scratch for launches that are being captured into a CUDA graph (see CaptureScope)
    uint64_t* g_desc = nullptr;
    uint64_t* g_wdesc = nullptr;
    uint64_t* g_sort_desc = nullptr;
    uint32_t* g_counter = nullptr;
    unsigned long long cap_id = 0;  // capture sequence the cap_* counters belong to
    bool cap_valid = false;
    int cap_cleared = 0;            // CaptureScope::Kind bits whose shadow array this capture already clears
    uint32_t cap_epoch = 0, cap_wepoch = 0, cap_sort_epoch = 0, cap_counter_base = 0;
    cudaStream_t last_stream = nullptr;  // stream of the previous launch that used the scratch
    bool has_last_stream = false;
    cudaEvent_t ev_order = nullptr;
    uint64_t host_chunk_seq = 0;    // chunks issued so far (staging buffer rotation across calls)
    // radix sort (lazy, grows on demand): 256 descriptors per tile shared by all passes,
    // 4 x 256 histogram + 4 x 256 digit bases
    uint64_t* sort_desc = nullptr;
    uint64_t sort_desc_tiles = 0;
    uint32_t* sort_hist = nullptr;
    uint32_t sort_epoch = 0;
    bool sort_configured[16] = {};  // [pairs + 2 * key-map variant + 4 * Decoupled Fallback + 8 * 64-bit keys]
    bool sort_hist_configured[2] = {false, false};  // [64-bit keys]
};

namespace {

struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev)
    {
        cudaGetDevice(&prev);
        if (prev != dev) cudaSetDevice(dev);
        else prev = -1;
    }
    ~DeviceGuard()
    {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

// The scratch of a ctx (ticket counter, descriptors, partials) is consumed in LAUNCH order: the
// host-side counter_base / epoch bookkeeping assumes the kernels of a ctx also EXECUTE in that
// order.  Launches on one stream do; when a launch arrives on a different stream than the
// previous one (e.g. gps_scan on the caller's stream, then the host pipeline on its own
// streams), it is ordered behind the previous launch with an event.
int enter_stream(gps_ctx* c, cudaStream_t st, bool* capturing = nullptr);

// The sharded path (reduce + publish, scan with carry slots) measures 1.5 % SLOWER with
// programmatic dependent launch (2 GPUs, 2^28 keys each: 0.4825 vs 0.4752 ms per step), so its
// launches are plain.
struct NoPdlScope {
    uint32_t& flags;
    const uint32_t saved;
    // mode ("sharded_pdl"): 0 = every launch of the pair plain, 1 = only the one-block fold
    // programmatic, 2 = fold and scan programmatic
    explicit NoPdlScope(uint32_t& f, int mode = 0) : flags(f), saved(f)
    {
        if (mode < 2) flags &= ~PF_PDL;
        if (mode >= 1 && (saved & PF_PDL)) flags |= PF_PDL_FOLD;
    }
    ~NoPdlScope() { flags = saved; }
};

template <typename T>
constexpr int dt_index() { return std::is_same<T, float>::value ? 1 : 0; }

// Decoupled Fallback, by default: from 2^27 keys the blocked shared-memory tile (836 Gkeys/s
// without stalls at 2^28 -- plain CSDL 844 -- and 632 with mask 7 / spin 4); below that the
// lane-strided smem tile when no stalls are emulated (822 vs 752 for the register tile) and
// the register tile otherwise (its slower polls spend the spin budget later: 527 vs 458);
// "fallback_variant" pins one
template <typename T, bool INCL>
cudaError_t launch_fallback(gps_ctx* c, const void* in, void* out, uint64_t n, const CarrySpec& cs,
                            const FallbackParams& f, cudaStream_t st)
{
    const bool smem = c->fallback_variant == 2 ||
                      (c->fallback_variant == 0 && n > GPS_REG_SMALL_MAX_N && f.stall_mask == kNoStall);
    const bool blocked = c->fallback_variant == 3 || (c->fallback_variant == 0 && n >= GPS_BLOCKED_MIN_N);
    if (blocked)
        return launch_csdl_blocked_fallback<T, GPS_BLOCKED_NT, GPS_BLOCKED_V, INCL>(
            c->s, in, out, n, cs, f, c->blocked_fb_configured[dt_index<T>()][INCL ? 0 : 1], st);
    if (smem)
        return launch_csdl_smem_fallback<T, GPS_SMEM_NT, GPS_SMEM_V, INCL>(
            c->s, in, out, n, cs, f, c->smem_fb_configured[dt_index<T>()][INCL ? 0 : 1], st);
    return launch_csdl_fallback<T, GPS_REG_NT, GPS_REG_V, INCL>(c->s, in, out, n, cs, f, st);
}

template <typename T, bool INCL>
int dispatch_scan_tk(gps_ctx* c, int algo, const void* in, void* out, uint64_t n,
                     const CarrySpec& cs, cudaStream_t st)
{
    cudaError_t e;
    if (n == 0) {
        e = launch_carry_only<T>(c->s, cs, st);
    } else {
        if (algo == GPS_ALGO_CSDL && c->csdl_variant == 0 && c->ticket_free && !(c->s.flags & PF_STATIC_TILE) &&
            n <= GPS_REG_SMALL_MAX_N) {
            // small sizes: when the whole grid is ONE wave (every CTA becomes resident without any other CTA of the
            // grid having to finish first), plain CSDL with tile = blockIdx.x cannot deadlock either -- whatever the
            // dispatch order, the tile a CTA waits for belongs to a CTA that is resident or will be as soon as
            // OTHER grids drain (a programmatically launched successor only starts once all of this grid's CTAs
            // have, pdl_entry).  The one assumption: kernels sharing the device terminate independently of this
            // scan (under in-order dispatch, cub::DeviceScan's assumption, not even that); "ticket_free" = 0 for
            // callers who can guarantee neither.  No ticket: 2^10 2.38 -> 2.10 us, 2^17 3.46 -> 3.04, 2^20 5.98 -> 5.14,
            // 2^21 7.27 -> 6.41
            const bool tiny = n <= GPS_REG_TINY_MAX_N && div_up(n, GPS_REG_TINY_NT * GPS_REG_TINY_V * 4) <= c->s.desc_tiles;
            const uint64_t tiles = tiny ? div_up(n, GPS_REG_TINY_NT * GPS_REG_TINY_V * 4) : div_up(n, GPS_REG_SMALL_NT * GPS_REG_SMALL_V * 4);
            int& resident = c->reg_resident[dt_index<T>()][INCL ? 0 : 1][tiny ? 1 : 0];
            if (!resident) {
                // (the smaller of the lean kernel's and the rescue-carrying kernel's occupancy: either may be launched)
                int per_sm = 0, per_sm_r = 0;
                cudaError_t q =
                    tiny ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(
                               &per_sm, csdl_reg_kernel<T, GPS_REG_TINY_NT, GPS_REG_TINY_V, INCL, GPS_LOOKBACK_WIN, false>, GPS_REG_TINY_NT, 0)
                         : cudaOccupancyMaxActiveBlocksPerMultiprocessor(
                               &per_sm, csdl_reg_kernel<T, GPS_REG_SMALL_NT, GPS_REG_SMALL_V, INCL, GPS_LOOKBACK_WIN, false>, GPS_REG_SMALL_NT, 0);
                if (q == cudaSuccess)
                    q = tiny ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(
                                   &per_sm_r, csdl_reg_kernel<T, GPS_REG_TINY_NT, GPS_REG_TINY_V, INCL, GPS_LOOKBACK_WIN, false, 24 + 4 * GPS_REG_TINY_V, true>,
                                   GPS_REG_TINY_NT, 0)
                             : cudaOccupancyMaxActiveBlocksPerMultiprocessor(
                                   &per_sm_r, csdl_reg_kernel<T, GPS_REG_SMALL_NT, GPS_REG_SMALL_V, INCL, GPS_LOOKBACK_WIN, false, 24 + 4 * GPS_REG_SMALL_V, true>,
                                   GPS_REG_SMALL_NT, 0);
                per_sm = std::min(per_sm, per_sm_r);
                resident = q == cudaSuccess && per_sm > 0 ? per_sm * c->s.sm_count : -1;
                if (q != cudaSuccess) cudaGetLastError();
            }
            if (resident > 0 && tiles <= (uint64_t)resident) {
                c->s.flags |= PF_STATIC_TILE;
                constexpr int RT = 24 + 4 * GPS_REG_TINY_V, RS = 24 + 4 * GPS_REG_SMALL_V;
                if (c->ticket_free == 1)  // with the warp-level rescue: needs nothing of the dispatch order or of other kernels
                    e = tiny ? launch_csdl_reg<T, GPS_REG_TINY_NT, GPS_REG_TINY_V, INCL, GPS_LOOKBACK_WIN, false, RT, true>(c->s, in, out, n, cs, st)
                             : launch_csdl_reg<T, GPS_REG_SMALL_NT, GPS_REG_SMALL_V, INCL, GPS_LOOKBACK_WIN, false, RS, true>(c->s, in, out, n, cs, st);
                else  // "ticket_free" = 2: the lean kernels
                    e = tiny ? launch_csdl_reg<T, GPS_REG_TINY_NT, GPS_REG_TINY_V, INCL, GPS_LOOKBACK_WIN>(c->s, in, out, n, cs, st)
                             : launch_csdl_reg<T, GPS_REG_SMALL_NT, GPS_REG_SMALL_V, INCL, GPS_LOOKBACK_WIN>(c->s, in, out, n, cs, st);
                c->s.flags &= ~PF_STATIC_TILE;
                if (e != cudaSuccess) return fail(GPS_ERR_CUDA, "scan launch failed: %s", cudaGetErrorString(e));
                return GPS_OK;
            }
        }
        if (algo == GPS_ALGO_CSDL && c->csdl_variant == 0 && c->ticket_free && !(c->s.flags & PF_STATIC_TILE) &&
            n <= GPS_TICKET_FREE_SMEM_MAX_N) {
            // ... and the sizes just above one wave of the <256,4> tile fit one wave of the <128,8> register tile
            // (9 CTAs per SM): the same argument, the same blockIdx order (gps_config.h GPS_REG_MID_*)
            const uint64_t tiles = div_up(n, GPS_REG_MID_NT * GPS_REG_MID_V * 4);
            int& resident = c->reg_resident[dt_index<T>()][INCL ? 0 : 1][2];
            if (!resident) {
                int per_sm = 0, per_sm_r = 0;
                cudaError_t q = cudaOccupancyMaxActiveBlocksPerMultiprocessor(
                    &per_sm, csdl_reg_kernel<T, GPS_REG_MID_NT, GPS_REG_MID_V, INCL, GPS_LOOKBACK_WIN, false>, GPS_REG_MID_NT, 0);
                if (q == cudaSuccess)
                    q = cudaOccupancyMaxActiveBlocksPerMultiprocessor(
                        &per_sm_r, csdl_reg_kernel<T, GPS_REG_MID_NT, GPS_REG_MID_V, INCL, GPS_LOOKBACK_WIN, false, 24 + 4 * GPS_REG_MID_V, true>,
                        GPS_REG_MID_NT, 0);
                per_sm = std::min(per_sm, per_sm_r);
                resident = q == cudaSuccess && per_sm > 0 ? per_sm * c->s.sm_count : -1;
                if (q != cudaSuccess) cudaGetLastError();
            }
            if (resident > 0 && tiles <= (uint64_t)resident && tiles <= c->s.desc_tiles) {
                c->s.flags |= PF_STATIC_TILE;
                constexpr int RM = 24 + 4 * GPS_REG_MID_V;
                e = c->ticket_free == 1
                        ? launch_csdl_reg<T, GPS_REG_MID_NT, GPS_REG_MID_V, INCL, GPS_LOOKBACK_WIN, false, RM, true>(c->s, in, out, n, cs, st)
                        : launch_csdl_reg<T, GPS_REG_MID_NT, GPS_REG_MID_V, INCL, GPS_LOOKBACK_WIN>(c->s, in, out, n, cs, st);
                c->s.flags &= ~PF_STATIC_TILE;
                if (e != cudaSuccess) return fail(GPS_ERR_CUDA, "scan launch failed: %s", cudaGetErrorString(e));
                return GPS_OK;
            }
        }
        if (algo == GPS_ALGO_CSDL && c->csdl_variant == 0 && c->ticket_free && !(c->s.flags & PF_STATIC_TILE)) {
            // mid sizes: Decoupled Fallback with tile = blockIdx.x (no ticket; deadlock-free under any dispatch
            // order because no CTA of that kernel waits for another one) -- gps_config.h has the measurements
            const bool reg_fb = n >= GPS_TICKET_FREE_REG_MIN_N && n <= GPS_TICKET_FREE_REG_MAX_N;
            const bool smem_fb = n > GPS_REG_SMALL_MAX_N && n <= GPS_TICKET_FREE_SMEM_MAX_N;
            if (reg_fb || smem_fb) {
                const FallbackParams f{kNoStall, GPS_FALLBACK_DEFAULT_SPIN, nullptr};
                c->s.flags |= PF_STATIC_TILE;
                e = reg_fb ? launch_csdl_fallback<T, GPS_REG_SMALL_NT, GPS_REG_SMALL_V, INCL>(c->s, in, out, n, cs, f, st)
                           : launch_csdl_smem_fallback<T, GPS_SMEM_NT, GPS_SMEM_V, INCL>(
                                 c->s, in, out, n, cs, f, c->smem_fb_configured[dt_index<T>()][INCL ? 0 : 1], st);
                c->s.flags &= ~PF_STATIC_TILE;
                if (e != cudaSuccess) return fail(GPS_ERR_CUDA, "scan launch failed: %s", cudaGetErrorString(e));
                return GPS_OK;
            }
        }
        if (algo == GPS_ALGO_CSDL) {
            const int v = c->csdl_variant ? c->csdl_variant : GPS_CSDL_DEFAULT_VARIANT;
            algo = v == 2 ? GPS_ALGO_CSDL_TMA : v == 3 ? GPS_ALGO_CSDL_SMEM : v == 4 ? GPS_ALGO_CSDL_BLOCKED : GPS_ALGO_CSDL_REG;
            // mid-size inputs: the lane-strided smem kernel's 16 KB tiles fill the machine sooner
            if (algo == GPS_ALGO_CSDL_BLOCKED && n < GPS_BLOCKED_MIN_N) algo = GPS_ALGO_CSDL_SMEM;
            // small inputs are launch/latency bound: the register path with 16 KB tiles
            if (algo == GPS_ALGO_CSDL_SMEM && n <= GPS_REG_SMALL_MAX_N) algo = GPS_ALGO_CSDL_REG;
        }
        switch (algo) {
            case GPS_ALGO_CSDL_REG:
                // small inputs: 16 KB tiles spread over more SMs (size sweep)
                // (the descriptor array is sized for 4096-key tiles: the 2048-key variant only
                // when its tiles fit, i.e. for every ctx of max_n >= 2^18)
                if (n <= GPS_REG_TINY_MAX_N && div_up(n, GPS_REG_TINY_NT * GPS_REG_TINY_V * 4) <= c->s.desc_tiles)
                    e = launch_csdl_reg<T, GPS_REG_TINY_NT, GPS_REG_TINY_V, INCL, GPS_LOOKBACK_WIN>(c->s, in, out, n, cs, st);
                else if (n <= GPS_REG_SMALL_MAX_N)
                    e = launch_csdl_reg<T, GPS_REG_SMALL_NT, GPS_REG_SMALL_V, INCL, GPS_LOOKBACK_WIN>(c->s, in, out, n, cs, st);
                else
                    e = launch_csdl_reg<T, GPS_REG_NT, GPS_REG_V, INCL, GPS_LOOKBACK_WIN>(c->s, in, out, n, cs, st);
                break;
            case GPS_ALGO_CSDL_BLOCKED:
                e = launch_csdl_blocked<T, GPS_BLOCKED_NT, GPS_BLOCKED_V, INCL, GPS_LOOKBACK_WIN>(
                    c->s, in, out, n, cs, c->blocked_configured[dt_index<T>()][INCL ? 0 : 1], st);
                break;
            case GPS_ALGO_CSDL_SMEM:
                e = launch_csdl_smem<T, GPS_SMEM_NT, GPS_SMEM_V, INCL, GPS_LOOKBACK_WIN>(
                    c->s, in, out, n, cs, c->smem_configured[dt_index<T>()][INCL ? 0 : 1], st);
                break;
            case GPS_ALGO_CSDL_TMA:
                e = launch_csdl_tma<T, GPS_TMA_NT, GPS_TMA_STAGES, INCL, GPS_LOOKBACK_WIN>(
                    c->s, in, out, n, cs, c->tma_ctas_per_sm[dt_index<T>()][INCL ? 0 : 1],
                    GPS_TMA_MAX_CTAS_PER_SM, st);
                break;
            case GPS_ALGO_CSDL_FALLBACK: {
                FallbackParams f{kNoStall, GPS_FALLBACK_DEFAULT_SPIN, nullptr};
                e = launch_fallback<T, INCL>(c, in, out, n, cs, f, st);
                break;
            }
            case GPS_ALGO_RTS:
                if (c->rts_variant == 2 || (c->rts_variant == 0 && n >= GPS_BLOCKED_MIN_N))
                    e = launch_rts_blocked<T, GPS_BLOCKED_NT, GPS_BLOCKED_V, INCL>(
                        c->s, in, out, n, cs, c->rts_blocked_configured[dt_index<T>()][INCL ? 0 : 1], st);
                else
                    e = launch_rts<T, GPS_RTS_NT, GPS_RTS_V, INCL>(c->s, in, out, n, cs, st);
                break;
            default:
                return fail(GPS_ERR_INVALID, "unknown algo %d", algo);
        }
    }
    if (e != cudaSuccess) return fail(GPS_ERR_CUDA, "scan launch failed: %s", cudaGetErrorString(e));
    return GPS_OK;
}

int check_scan_args(gps_ctx* c, int kind, int dtype, const void* in, void* out, uint64_t n)
{
    if (!c) return fail(GPS_ERR_INVALID, "null ctx");
    if (kind != GPS_INCLUSIVE && kind != GPS_EXCLUSIVE) return fail(GPS_ERR_INVALID, "bad kind %d", kind);
    if (dtype != GPS_U32 && dtype != GPS_F32) return fail(GPS_ERR_INVALID, "bad dtype %d", dtype);
    if (n > c->max_n) return fail(GPS_ERR_INVALID, "n=%llu exceeds ctx capacity %llu",
                                  (unsigned long long)n, (unsigned long long)c->max_n);
    if (n && (!in || !out)) return fail(GPS_ERR_INVALID, "null data pointer");
    if (!aligned16(in) || !aligned16(out)) return fail(GPS_ERR_ALIGNMENT, "data pointers must be 16-byte aligned");
    if (n && in != out) {
        const char* a = static_cast<const char*>(in);
        const char* b = static_cast<const char*>(out);
        if (a < b + n * 4 && b < a + n * 4)
            return fail(GPS_ERR_INVALID, "d_in and d_out partially overlap");
    }
    return GPS_OK;
}

int enter_stream(gps_ctx* c, cudaStream_t st, bool* capturing)
{
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(st, &cap) != cudaSuccess) cudaGetLastError();
    if (capturing) *capturing = cap != cudaStreamCaptureStatusNone;
    if (cap != cudaStreamCaptureStatusNone) return GPS_OK;  // ordering inside a capture is the caller's graph structure
    if (c->has_last_stream && c->last_stream != st) {
        if (!c->ev_order) GPS_CUDA(cudaEventCreateWithFlags(&c->ev_order, cudaEventDisableTiming));
        // (a caller who destroyed the previous stream has necessarily drained it: ignore the error)
        if (cudaEventRecord(c->ev_order, c->last_stream) == cudaSuccess)
            GPS_CUDA(cudaStreamWaitEvent(st, c->ev_order, 0));
        else
            cudaGetLastError();
    }
    c->last_stream = st;
    c->has_last_stream = true;
    return GPS_OK;
}

// CUDA graph capture.  The chained kernels (and the sort) normally take a launch epoch and a
// ticket base from host-side counters baked into their parameters -- a REPLAYED graph would see
// the previous replay's descriptors as valid and hand out stale tile numbers.  A captured launch
// therefore runs on a SHADOW scratch (own descriptor arrays and ticket counter, allocated with
// the ctx).  The FIRST captured launch of a capture sequence (per kind: scan / wide / sort)
// records a clear of that shadow array (and of the ticket counter) into the graph; the launches
// that follow in the same capture take consecutive epochs and ticket bases tracked per capture.
// Every replay thus starts from the same cleared state and no two launches of one replay share
// an epoch -- one memset per graph, not per scan (the reference clears per launch,
// ChainedScanDecoupledLookbackDispatcher.cuh:172-175).  The host-side counters of the ctx are
// untouched, so plain launches continue unaffected.
struct CaptureScope {
    gps_ctx* c = nullptr;
    Scratch saved;
    uint64_t* saved_sort_desc = nullptr;
    uint32_t saved_sort_epoch = 0;
    bool active = false;
    enum Kind { SCAN = 1, WIDE = 2, SORT = 4 };
    int begin(gps_ctx* ctx, cudaStream_t st, int kind)
    {
        c = ctx;
        cudaStreamCaptureStatus status = cudaStreamCaptureStatusNone;
        unsigned long long id = 0;
        GPS_CUDA(cudaStreamGetCaptureInfo(st, &status, &id));
        if (kind == SORT && !c->g_sort_desc)
            return fail(GPS_ERR_UNSUPPORTED, "capture needs one gps_sort call of at least this size outside the capture first");
        if (id != c->cap_id || !c->cap_valid) {  // a new capture sequence
            c->cap_id = id;
            c->cap_valid = true;
            c->cap_cleared = 0;
            c->cap_epoch = c->cap_wepoch = c->cap_sort_epoch = 0;
            c->cap_counter_base = 0;
            GPS_CUDA(cudaMemsetAsync(c->g_counter, 0, 64, st));
        }
        if (!(c->cap_cleared & kind)) {
            if (kind == SCAN) GPS_CUDA(cudaMemsetAsync(c->g_desc, 0, (size_t)c->s.desc_tiles * sizeof(uint64_t), st));
            if (kind == WIDE) GPS_CUDA(cudaMemsetAsync(c->g_wdesc, 0, (size_t)c->s.wdesc_words * sizeof(uint64_t), st));
            if (kind == SORT)
                GPS_CUDA(cudaMemsetAsync(c->g_sort_desc, 0, (size_t)c->sort_desc_tiles * kRadix * sizeof(uint64_t), st));
            c->cap_cleared |= kind;
        }
        saved = c->s;
        saved_sort_desc = c->sort_desc;
        saved_sort_epoch = c->sort_epoch;
        active = true;
        c->s.desc = c->g_desc;
        c->s.wdesc = c->g_wdesc;
        c->s.counter = c->g_counter;
        c->s.counter_base = c->cap_counter_base;
        c->s.epoch = c->cap_epoch;
        c->s.wepoch = c->cap_wepoch;
        c->sort_desc = c->g_sort_desc;
        c->sort_epoch = c->cap_sort_epoch;
        return GPS_OK;
    }
    ~CaptureScope()
    {
        if (!active) return;
        c->cap_counter_base = c->s.counter_base;  // what this launch consumed, for the next one
        c->cap_epoch = c->s.epoch;
        c->cap_wepoch = c->s.wepoch;
        c->cap_sort_epoch = c->sort_epoch;
        c->s = saved;
        c->sort_desc = saved_sort_desc;
        c->sort_epoch = saved_sort_epoch;
    }
};

int dispatch_scan(gps_ctx* c, int algo, int kind, int dtype, const void* in, void* out, uint64_t n,
                  const CarrySpec& cs, cudaStream_t st)
{
    int rc = check_scan_args(c, kind, dtype, in, out, n);
    if (rc) return rc;
    DeviceGuard g(c->device);
    bool capturing = false;
    rc = enter_stream(c, st, &capturing);
    if (rc) return rc;
    CaptureScope cap;
    if (capturing && algo != GPS_ALGO_RTS) {
        rc = cap.begin(c, st, CaptureScope::SCAN);
        if (rc) return rc;
    }
    // L2 policies (gps_common.cuh): auto = evict_first loads while the output is 38..76 % of the L2 (gps_config.h)
    struct FlagScope {
        uint32_t& f;
        const uint32_t saved;
        ~FlagScope() { f = saved; }
    } l2scope{c->s.flags, c->s.flags};
    {
        uint32_t bits = 0;
        if (c->l2_hints >= 0) bits = (uint32_t)c->l2_hints & 0xFFu;  // (bit 2: stores evict_first, internal)
        else if (n * 4 <= c->l2_bytes * GPS_L2_HINT_MAX_OUT_PCT / 100 && n * 4 >= c->l2_bytes * GPS_L2_HINT_MIN_OUT_PCT / 100)
            bits = GPS_L2_HINT_AUTO_BITS;
        c->s.flags = (c->s.flags & ~(PF_L2_IN_EVICT_FIRST | PF_L2_OUT_EVICT_LAST | PF_L2_OUT_EVICT_FIRST)) |
                     ((bits & 1u) ? PF_L2_IN_EVICT_FIRST : 0u) | ((bits & 2u) ? PF_L2_OUT_EVICT_LAST : 0u) |
                     ((bits & 4u) ? PF_L2_OUT_EVICT_FIRST : 0u);
    }
    if (dtype == GPS_U32)
        return kind == GPS_INCLUSIVE ? dispatch_scan_tk<uint32_t, true>(c, algo, in, out, n, cs, st)
                                     : dispatch_scan_tk<uint32_t, false>(c, algo, in, out, n, cs, st);
    return kind == GPS_INCLUSIVE ? dispatch_scan_tk<float, true>(c, algo, in, out, n, cs, st)
                                 : dispatch_scan_tk<float, false>(c, algo, in, out, n, cs, st);
}

template <typename E>
E wide_zero_host()
{
    E z;
    memset(&z, 0, sizeof(E));
    return z;
}

// bytes of wide-element data one ctx can scan = 4 * max_n; descriptor words for the smallest
// wide tile (GPS_WIDE_NT x GPS_WIDE_V vectors of 16 bytes) at W = 4 words per element
constexpr uint64_t kWideTileBytes = (uint64_t)GPS_WIDE_NT * GPS_WIDE_V * 16;
inline uint64_t wide_desc_words(uint64_t max_n) { return std::max<uint64_t>(1, div_up(max_n * 4, kWideTileBytes)) * 8; }

template <typename E, bool INCL>
int launch_wide(gps_ctx* c, int algo, const void* in, void* out, uint64_t n, const void* h_carry, void* total_out,
                bool& configured, cudaStream_t st, const WideFallbackParams* fbp = nullptr)
{
    using S = WideShape<E, GPS_WIDE_NT, GPS_WIDE_V>;
    Scratch& s = c->s;
    WideParams<E> p{};
    p.in = in;
    p.out = out;
    p.n = n;
    p.desc = s.wdesc;
    p.counter = s.counter;
    p.counter_base = s.counter_base;
    p.partials = s.partials;
    p.status = s.status;
    p.num_tiles = (uint32_t)div_up(n, S::TILE_ELEMS);
    p.carry = wide_zero_host<E>();
    if (h_carry) memcpy(&p.carry, h_carry, sizeof(E));
    p.total_out = total_out;
    if ((uint64_t)p.num_tiles * 2 * S::W > s.wdesc_words || p.num_tiles > s.partials_capacity)
        return fail(GPS_ERR_INVALID, "n exceeds ctx capacity");
    const bool pdl = s.flags & PF_PDL;
    if (p.num_tiles == 0) {  // only the carry travels
        GPS_CUDA(launch_scan_kernel(wide_spine_kernel<E>, 1, 1024, 0, st, pdl, p));
        return GPS_OK;
    }
    if (algo == GPS_ALGO_RTS) {
        auto kern = wide_scan_kernel<E, GPS_WIDE_NT, GPS_WIDE_V, INCL, false, GPS_WIDE_REGS>;
        if (!configured) {
            GPS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout,
                                          (int)cudaSharedmemCarveoutMaxShared));
            configured = true;
        }
        GPS_CUDA(launch_scan_kernel(wide_reduce_kernel<E, GPS_WIDE_NT, GPS_WIDE_V>, p.num_tiles, GPS_WIDE_NT, 0, st, pdl, p));
        GPS_CUDA(launch_scan_kernel(wide_spine_kernel<E>, 1, 1024, 0, st, pdl, p));
        GPS_CUDA(launch_scan_kernel(kern, p.num_tiles, GPS_WIDE_NT, 0, st, pdl, p));
    } else if (algo == GPS_ALGO_CSDL_FALLBACK) {
        auto kern = wide_scan_fallback_kernel<E, GPS_WIDE_NT, GPS_WIDE_V, INCL, GPS_WIDE_FB_REGS>;
        if (!configured) {
            GPS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout,
                                          (int)cudaSharedmemCarveoutMaxShared));
            configured = true;
        }
        if (++s.wepoch == 0) {
            GPS_CUDA(cudaMemsetAsync(s.wdesc, 0, (size_t)s.wdesc_words * sizeof(uint64_t), st));
            s.wepoch = 1;
        }
        p.epoch = s.wepoch;
        const WideFallbackParams f = fbp ? *fbp : WideFallbackParams{0xFFFFFFFFu, GPS_FALLBACK_DEFAULT_SPIN, nullptr};
        const cudaError_t e = launch_scan_kernel(kern, p.num_tiles, GPS_WIDE_NT, 0, st, pdl, p, f);
        if (e == cudaSuccess) s.counter_base += p.num_tiles;  // one claim per CTA
        GPS_CUDA(e);
    } else {
        auto kern = wide_scan_kernel<E, GPS_WIDE_NT, GPS_WIDE_V, INCL, true, GPS_WIDE_REGS>;
        if (!configured) {
            GPS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout,
                                          (int)cudaSharedmemCarveoutMaxShared));
            configured = true;
        }
        if (++s.wepoch == 0) {  // wrap: epoch values would alias -> clear once
            GPS_CUDA(cudaMemsetAsync(s.wdesc, 0, (size_t)s.wdesc_words * sizeof(uint64_t), st));
            s.wepoch = 1;
        }
        p.epoch = s.wepoch;
        const cudaError_t e = launch_scan_kernel(kern, p.num_tiles, GPS_WIDE_NT, 0, st, pdl, p);
        if (e == cudaSuccess) s.counter_base += p.num_tiles;  // one claim per CTA
        GPS_CUDA(e);
    }
    GPS_CUDA(cudaGetLastError());
    return GPS_OK;
}

int fill_publish(PublishList& pl, uint64_t* const* dst, uint32_t n_dst, uint32_t tag)
{
    if (n_dst > (uint32_t)kMaxPublish)
        return fail(GPS_ERR_INVALID, "at most %d publish destinations", kMaxPublish);
    pl.n = n_dst;
    pl.tag = tag;
    for (uint32_t i = 0; i < n_dst; ++i) pl.dst[i] = dst[i];
    return GPS_OK;
}

}  // namespace

// ---------------------------------------------------------------------------------
// library / device
// ---------------------------------------------------------------------------------
namespace {
// tile shapes of the digit pass per key width: 32-bit keys <NT, KPT> / pairs, 64-bit keys / pairs
template <typename K, bool PAIRS>
struct SortShape {
    static constexpr bool W64 = sizeof(K) == 8;
    static constexpr int NT = W64 ? GPS_SORT64_NT : (PAIRS ? GPS_SORT_PAIRS_NT : GPS_SORT_NT);
    static constexpr int KPT = W64 ? (PAIRS ? GPS_SORT64_PAIRS_KPT : GPS_SORT64_KPT) : (PAIRS ? GPS_SORT_PAIRS_KPT : GPS_SORT_KPT);
    static constexpr int REGS = W64 ? GPS_SORT64_REGS : (PAIRS ? GPS_SORT_PAIRS_REGS : GPS_SORT_REGS);
    static constexpr uint32_t TILE = (uint32_t)NT * KPT;
    static constexpr size_t SMEM = ((size_t)(sizeof(K) / 4 + (PAIRS ? 1 : 0)) * TILE + (size_t)(NT / 32) * kRadix) * sizeof(uint32_t);
};
constexpr uint32_t cmin(uint32_t a, uint32_t b) { return a < b ? a : b; }
constexpr uint32_t kSortMinTile = cmin(cmin(SortShape<uint32_t, false>::TILE, SortShape<uint32_t, true>::TILE),
                                       cmin(SortShape<uint64_t, false>::TILE, SortShape<uint64_t, true>::TILE));

template <typename K, bool PAIRS, bool XFORM, bool FB = false>
cudaError_t launch_sort_pass(const SortParams& p, bool& configured, bool pdl, cudaStream_t st,
                             const SortFallbackParams& f = SortFallbackParams{0xFFFFFFFFu, 1u, nullptr})
{
    using S = SortShape<K, PAIRS>;
    auto kern = sort_pass_kernel<K, S::NT, S::KPT, PAIRS, S::REGS, GPS_SORT_LB, XFORM, FB>;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S::SMEM);
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    return launch_scan_kernel(kern, p.num_tiles, S::NT, S::SMEM, st, pdl, p, f);
}
template <typename K>
int sort_impl(gps_ctx* c, int key_type, const void* d_keys_in, void* d_keys_out, void* d_keys_tmp,
                     const void* d_vals_in, void* d_vals_out, void* d_vals_tmp, uint64_t n, int end_bit,
                     void* stream, const SortFallbackParams* fbp)
{
    constexpr int kBits = (int)sizeof(K) * 8;
    constexpr bool W64 = sizeof(K) == 8;
    if (!c) return fail(GPS_ERR_INVALID, "null ctx");
    const int base_type = key_type & ~GPS_KEY_DESCENDING;
    if (base_type != GPS_KEY_U32 && base_type != GPS_KEY_I32 && base_type != GPS_KEY_F32)
        return fail(GPS_ERR_INVALID, "bad key type %d", key_type);
    if (key_type != GPS_KEY_U32 && end_bit != kBits)
        return fail(GPS_ERR_INVALID, "signed / floating-point / descending sorts use all %d bits (end_bit %d)", kBits, kBits);
    if (end_bit < 1 || end_bit > kBits) return fail(GPS_ERR_INVALID, "end_bit %d outside 1..%d", end_bit, kBits);
    if (n >= (1ull << 32)) return fail(GPS_ERR_INVALID, "the sort takes fewer than 2^32 keys");
    const bool pairs = d_vals_in || d_vals_out || d_vals_tmp;
    if (n == 0) return GPS_OK;
    if (!d_keys_in || !d_keys_out || !d_keys_tmp) return fail(GPS_ERR_INVALID, "null key pointer");
    if (pairs && (!d_vals_in || !d_vals_out || !d_vals_tmp)) return fail(GPS_ERR_INVALID, "null value pointer");
    if (!aligned16(d_keys_in) || !aligned16(d_keys_out) || !aligned16(d_keys_tmp))
        return fail(GPS_ERR_ALIGNMENT, "key pointers must be 16-byte aligned");
    if (pairs && (!aligned16(d_vals_in) || !aligned16(d_vals_out) || !aligned16(d_vals_tmp)))
        return fail(GPS_ERR_ALIGNMENT, "value pointers must be 16-byte aligned");
    if (d_keys_tmp == d_keys_in || d_keys_tmp == d_keys_out || (pairs && (d_vals_tmp == d_vals_in || d_vals_tmp == d_vals_out)))
        return fail(GPS_ERR_INVALID, "the tmp buffer must be distinct from in and out");
    const int passes = (end_bit + 7) / 8;
    // bits of the last digit that take part: a sort on end_bit bits orders by key & (2^end_bit - 1)
    // and nothing else (cub's end_bit), so the last pass masks its digit when end_bit % 8 != 0
    const uint32_t last_mask = (end_bit & 7) ? ((1u << (end_bit & 7)) - 1u) : 0xFFu;
    if ((passes & 1) && (d_keys_in == d_keys_out || (pairs && d_vals_in == d_vals_out)))
        return fail(GPS_ERR_INVALID, "in == out needs an even number of digit passes (ceil(end_bit / 8) even)");
    DeviceGuard g(c->device);
    cudaStream_t st = (cudaStream_t)stream;
    bool capturing = false;
    if (int rc = enter_stream(c, st, &capturing)) return rc;
    const uint32_t tile = pairs ? SortShape<K, true>::TILE : SortShape<K, false>::TILE;
    const uint32_t tiles = (uint32_t)div_up(n, tile);
    if (capturing && tiles > c->sort_desc_tiles)
        return fail(GPS_ERR_UNSUPPORTED, "capture needs one gps_sort call of at least this size outside the capture first");
    CaptureScope cap;
    if (capturing)
        if (int rc = cap.begin(c, st, CaptureScope::SORT)) return rc;
    if (tiles > c->sort_desc_tiles) {  // grow (synchronous, rare)
        GPS_CUDA(cudaDeviceSynchronize());
        if (c->sort_desc) cudaFree(c->sort_desc);
        c->sort_desc = nullptr;
        c->sort_desc_tiles = 0;
        GPS_CUDA(cudaMalloc(&c->sort_desc, (size_t)tiles * kRadix * sizeof(uint64_t)));
        GPS_CUDA(cudaMemset(c->sort_desc, 0, (size_t)tiles * kRadix * sizeof(uint64_t)));
        if (c->g_sort_desc) cudaFree(c->g_sort_desc);  // shadow for captured sorts (CaptureScope)
        c->g_sort_desc = nullptr;
        GPS_CUDA(cudaMalloc(&c->g_sort_desc, (size_t)tiles * kRadix * sizeof(uint64_t)));
        c->sort_desc_tiles = tiles;
        c->sort_epoch = 0;
    }
    if (!c->sort_hist) GPS_CUDA(cudaMalloc(&c->sort_hist, 2 * kMaxSortPasses * kRadix * sizeof(uint32_t)));
    uint32_t* hist = c->sort_hist;
    uint32_t* digit_base = c->sort_hist + kMaxSortPasses * kRadix;
    GPS_CUDA(cudaMemsetAsync(hist, 0, kMaxSortPasses * kRadix * sizeof(uint32_t), st));
    {
        bool& cfg = c->sort_hist_configured[W64 ? 1 : 0];
        if (!cfg) {
            GPS_CUDA(cudaFuncSetAttribute(sort_hist_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kHistSmemBytes));
            cfg = true;
        }
        const uint32_t grid = (uint32_t)std::min<uint64_t>((uint64_t)c->s.sm_count, div_up(n, (16 / sizeof(K)) * kHistThreads));
        for (int pass0 = 0; pass0 < passes; pass0 += 4)  // the shared-memory table holds 4 passes: 64-bit keys take two sweeps
            sort_hist_kernel<K><<<grid, kHistThreads, kHistSmemBytes, st>>>((const K*)d_keys_in, (uint32_t)n, hist, pass0,
                                                                        std::min(4, passes - pass0), passes,
                                                                        (uint32_t)key_type, last_mask);
        sort_digit_base_kernel<<<passes, kRadix, 0, st>>>(hist, digit_base);
    }
    // ping-pong so that the last pass lands in `out`
    const void* src_k = d_keys_in;
    const uint32_t* src_v = (const uint32_t*)d_vals_in;
    const int cfg0 = W64 ? 8 : 0;
    for (int pass = 0; pass < passes; ++pass) {
        const bool to_out = ((passes - 1 - pass) & 1) == 0;
        SortParams p{};
        p.keys_in = src_k;
        p.keys_out = to_out ? d_keys_out : d_keys_tmp;
        p.vals_in = src_v;
        p.vals_out = (uint32_t*)(to_out ? d_vals_out : d_vals_tmp);
        p.n = (uint32_t)n;
        p.shift = 8u * pass;
        p.desc = c->sort_desc;
        p.digit_base = digit_base + pass * kRadix;
        p.counter = c->s.counter;
        p.counter_base = c->s.counter_base;
        if (++c->sort_epoch >= (1u << 30)) {  // wrap: epoch values would alias -> clear once
            GPS_CUDA(cudaMemsetAsync(c->sort_desc, 0, (size_t)c->sort_desc_tiles * kRadix * sizeof(uint64_t), st));
            c->sort_epoch = 1;
        }
        p.epoch = c->sort_epoch;
        p.num_tiles = tiles;
        p.status = c->s.status;
        p.xform_in = pass == 0 ? (uint32_t)key_type : 0u;
        p.xform_out = pass == passes - 1 ? (uint32_t)key_type : 0u;
        p.digit_mask = pass == passes - 1 ? last_mask : 0xFFu;
        const bool xf = p.xform_in || p.xform_out || p.digit_mask != 0xFFu;
        const bool pdl = c->s.flags & PF_PDL;
        cudaError_t e;
        if (fbp)
            e = pairs ? (xf ? launch_sort_pass<K, true, true, true>(p, c->sort_configured[cfg0 + 7], pdl, st, *fbp)
                            : launch_sort_pass<K, true, false, true>(p, c->sort_configured[cfg0 + 5], pdl, st, *fbp))
                      : (xf ? launch_sort_pass<K, false, true, true>(p, c->sort_configured[cfg0 + 6], pdl, st, *fbp)
                            : launch_sort_pass<K, false, false, true>(p, c->sort_configured[cfg0 + 4], pdl, st, *fbp));
        else
            e = pairs ? (xf ? launch_sort_pass<K, true, true>(p, c->sort_configured[cfg0 + 3], pdl, st)
                            : launch_sort_pass<K, true, false>(p, c->sort_configured[cfg0 + 1], pdl, st))
                      : (xf ? launch_sort_pass<K, false, true>(p, c->sort_configured[cfg0 + 2], pdl, st)
                            : launch_sort_pass<K, false, false>(p, c->sort_configured[cfg0 + 0], pdl, st));
        if (e == cudaSuccess) c->s.counter_base += tiles;  // one claim per CTA of a launch that happened
        GPS_CUDA(e);
        src_k = p.keys_out;
        src_v = p.vals_out;
    }
    GPS_CUDA(cudaGetLastError());
    return GPS_OK;
}

}  // namespace

extern "C" {

int gps_abi_version(void) { return GPS_ABI_VERSION; }
const char* gps_last_error(void) { return g_err; }

int gps_device_count(int* count)
{
    if (!count) return fail(GPS_ERR_INVALID, "null count");
    *count = 0;
    GPS_CUDA(cudaGetDeviceCount(count));
    return GPS_OK;
}

int gps_device_info(int device, char* name, int* sm_count, uint64_t* l2_bytes, uint64_t* mem_bytes)
{
    cudaDeviceProp prop;
    GPS_CUDA(cudaGetDeviceProperties(&prop, device));
    if (name) {
        strncpy(name, prop.name, 255);
        name[255] = 0;
    }
    if (sm_count) *sm_count = prop.multiProcessorCount;
    if (l2_bytes) *l2_bytes = (uint64_t)prop.l2CacheSize;
    if (mem_bytes) *mem_bytes = (uint64_t)prop.totalGlobalMem;
    return GPS_OK;
}

int gps_set_device(int device)
{
    GPS_CUDA(cudaSetDevice(device));
    return GPS_OK;
}

// ---------------------------------------------------------------------------------
// context
// ---------------------------------------------------------------------------------
namespace {
// one device allocation (the ctx's own, or the caller's), carved into 256-byte aligned segments
struct ScratchLayout {
    uint64_t desc_tiles, partials_capacity, wdesc_words;
    uint64_t off_desc, off_partials, off_wdesc, off_counter, off_gdesc, off_gwdesc, off_gcounter, total;
};
inline uint64_t align256(uint64_t v) { return (v + 255u) & ~(uint64_t)255u; }
ScratchLayout scratch_layout(uint64_t max_n)
{
    ScratchLayout L{};
    L.desc_tiles = std::max<uint64_t>(1, div_up(max_n, kMinTileKeys));
    L.wdesc_words = wide_desc_words(max_n);
    L.partials_capacity = std::max<uint64_t>(L.desc_tiles, L.wdesc_words / 8);  // wide tiles may be smaller
    L.off_desc = 0;
    L.off_partials = align256(L.off_desc + L.desc_tiles * sizeof(uint64_t));
    L.off_wdesc = align256(L.off_partials + L.partials_capacity * 16);  // up to 16-byte elements
    L.off_counter = align256(L.off_wdesc + L.wdesc_words * sizeof(uint64_t));
    // shadow descriptors / ticket counter for launches captured into a CUDA graph (CaptureScope)
    L.off_gdesc = L.off_counter + 256;
    L.off_gwdesc = align256(L.off_gdesc + L.desc_tiles * sizeof(uint64_t));
    L.off_gcounter = align256(L.off_gwdesc + L.wdesc_words * sizeof(uint64_t));
    L.total = L.off_gcounter + 256;
    return L;
}
}  // namespace

int gps_create_with_scratch(gps_ctx** out, int device, uint64_t max_n, void* d_scratch, uint64_t scratch_bytes)
{
    if (!out) return fail(GPS_ERR_INVALID, "null ctx out");
    *out = nullptr;
    int count = 0;
    GPS_CUDA(cudaGetDeviceCount(&count));
    if (device < 0 || device >= count) return fail(GPS_ERR_INVALID, "device %d of %d", device, count);
    if (div_up(max_n, kMinTileKeys) >= (1ull << 31))
        return fail(GPS_ERR_INVALID, "max_n too large");
    const ScratchLayout L = scratch_layout(max_n);
    if (d_scratch) {
        if (reinterpret_cast<uintptr_t>(d_scratch) & 255u)
            return fail(GPS_ERR_ALIGNMENT, "caller scratch must be 256-byte aligned");
        if (scratch_bytes < L.total)
            return fail(GPS_ERR_INVALID, "caller scratch of %llu bytes is smaller than gps_scratch_bytes() = %llu",
                        (unsigned long long)scratch_bytes, (unsigned long long)L.total);
    }
    DeviceGuard g(device);
    cudaDeviceProp prop;
    GPS_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
        return fail(GPS_ERR_UNSUPPORTED, "device %d is sm_%d%d; this library is built for sm_100a only",
                    device, prop.major, prop.minor);
    gps_ctx* c = new gps_ctx();
    c->device = device;
    c->max_n = max_n;
    c->s.sm_count = prop.multiProcessorCount;
    c->l2_bytes = (uint64_t)prop.l2CacheSize;
    c->s.flags |= PF_PDL;  // programmatic dependent launch by default (option "pdl")
    c->owns_scratch = d_scratch == nullptr;
    cudaError_t e = cudaSuccess;
    if (c->owns_scratch) e = cudaMalloc(&c->scratch_base, (size_t)L.total);
    else c->scratch_base = d_scratch;
    // descriptors and counters start at zero ONCE; epoch tags replace per-launch clears
    if (e == cudaSuccess) e = cudaMemset(c->scratch_base, 0, (size_t)L.total);
    if (e != cudaSuccess) {
        gps_destroy(c);
        return fail(GPS_ERR_CUDA, "scratch allocation failed: %s", cudaGetErrorString(e));
    }
    unsigned char* base = static_cast<unsigned char*>(c->scratch_base);
    c->s.desc = reinterpret_cast<uint64_t*>(base + L.off_desc);
    c->s.desc_tiles = (uint32_t)L.desc_tiles;
    c->s.partials = base + L.off_partials;
    c->s.partials_capacity = (uint32_t)L.partials_capacity;
    c->s.wdesc = reinterpret_cast<uint64_t*>(base + L.off_wdesc);
    c->s.wdesc_words = L.wdesc_words;
    c->s.counter = reinterpret_cast<uint32_t*>(base + L.off_counter);
    c->s.status = c->s.counter + 1;
    c->g_desc = reinterpret_cast<uint64_t*>(base + L.off_gdesc);
    c->g_wdesc = reinterpret_cast<uint64_t*>(base + L.off_gwdesc);
    c->g_counter = reinterpret_cast<uint32_t*>(base + L.off_gcounter);
    *out = c;
    return GPS_OK;
}

int gps_create(gps_ctx** out, int device, uint64_t max_n) { return gps_create_with_scratch(out, device, max_n, nullptr, 0); }

uint64_t gps_scratch_bytes(uint64_t max_n) { return scratch_layout(max_n).total; }

void gps_destroy(gps_ctx* c)
{
    if (!c) return;
    DeviceGuard g(c->device);
    if (c->owns_scratch && c->scratch_base) cudaFree(c->scratch_base);
    for (int i = 0; i < kHostBufs; ++i) {
        if (c->stage[i]) cudaFree(c->stage[i]);
        if (c->ev_h2d[i]) cudaEventDestroy(c->ev_h2d[i]);
        if (c->ev_cmp[i]) cudaEventDestroy(c->ev_cmp[i]);
        if (c->ev_d2h[i]) cudaEventDestroy(c->ev_d2h[i]);
    }
    if (c->chain_slots) cudaFree(c->chain_slots);
    if (c->sort_desc) cudaFree(c->sort_desc);
    if (c->g_sort_desc) cudaFree(c->g_sort_desc);
    if (c->sort_hist) cudaFree(c->sort_hist);
    if (c->chain_total) cudaFree(c->chain_total);
    if (c->ev_total) cudaEventDestroy(c->ev_total);
    if (c->ev_order) cudaEventDestroy(c->ev_order);
    if (c->s_h2d) cudaStreamDestroy(c->s_h2d);
    if (c->s_cmp) cudaStreamDestroy(c->s_cmp);
    if (c->s_d2h) cudaStreamDestroy(c->s_d2h);
    delete c;
}

int gps_check(gps_ctx* c, void* stream)
{
    if (!c) return fail(GPS_ERR_INVALID, "null ctx");
    DeviceGuard g(c->device);
    GPS_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    uint32_t st = 0;
    GPS_CUDA(cudaMemcpy(&st, c->s.status, sizeof(st), cudaMemcpyDeviceToHost));
    if (st & STATUS_RESCUED) {  // informational: a ticket-free one-wave scan rescued a late tile (results are valid)
        c->rescued = 1;
        st &= ~STATUS_RESCUED;
        if (!st) GPS_CUDA(cudaMemset(c->s.status, 0, sizeof(uint32_t)));
    }
    if (st) {
        // CTAs that bailed out may have skipped ticket claims: bring the host's ticket base back
        // in step with the device counter once everything has drained
        GPS_CUDA(cudaDeviceSynchronize());
        uint32_t counter = 0;
        GPS_CUDA(cudaMemcpy(&counter, c->s.counter, sizeof(counter), cudaMemcpyDeviceToHost));
        c->s.counter_base = counter;
        GPS_CUDA(cudaMemset(c->s.status, 0, sizeof(uint32_t)));
        return fail(GPS_ERR_WATCHDOG, "%s (status 0x%x): results invalid",
                    (st & STATUS_SLOT_TIMEOUT) ? "a carry slot never arrived" : "look-back poll budget exceeded", st);
    }
    return GPS_OK;
}

int gps_get_option(gps_ctx* c, const char* key, int64_t* value)
{
    if (!c || !key || !value) return fail(GPS_ERR_INVALID, "null argument");
    if (!strcmp(key, "rescued")) {  // 1 if a gps_check since the last query saw a rescued tile; reading clears it
        *value = c->rescued;
        c->rescued = 0;
    } else if (!strcmp(key, "ticket_free")) {
        *value = c->ticket_free;
    } else if (!strcmp(key, "pdl")) {
        *value = (c->s.flags & PF_PDL) ? 1 : 0;
    } else if (!strcmp(key, "tile_order")) {
        *value = (c->s.flags & PF_STATIC_TILE) ? 1 : 0;
    } else {
        return fail(GPS_ERR_INVALID, "unknown option '%s'", key);
    }
    return GPS_OK;
}

int gps_set_option(gps_ctx* c, const char* key, int64_t value)
{
    if (!c || !key) return fail(GPS_ERR_INVALID, "null argument");
    if (!strcmp(key, "csdl_variant")) {
        if (value < 0 || value > 4) return fail(GPS_ERR_INVALID, "csdl_variant must be 0..4");
        c->csdl_variant = (int)value;
    } else if (!strcmp(key, "ticket_free")) {
        if (value < 0 || value > 2) return fail(GPS_ERR_INVALID, "ticket_free must be 0, 1 or 2");
        c->ticket_free = (int)value;
    } else if (!strcmp(key, "fallback_variant")) {
        if (value < 0 || value > 3) return fail(GPS_ERR_INVALID, "fallback_variant must be 0..3");
        c->fallback_variant = (int)value;
    } else if (!strcmp(key, "pdl")) {
        // programmatic dependent launch for the CSDL kernels (register / smem / blocked tiles)
        if (value != 0 && value != 1) return fail(GPS_ERR_INVALID, "pdl must be 0 or 1");
        c->s.flags = (c->s.flags & ~PF_PDL) | (value ? PF_PDL : 0u);
    } else if (!strcmp(key, "rts_variant")) {
        if (value < 0 || value > 2) return fail(GPS_ERR_INVALID, "rts_variant must be 0..2");
        c->rts_variant = (int)value;
    } else if (!strcmp(key, "debug_epoch")) {
        // test hook: jump the launch epoch (e.g. just below the 2^30 wrap)
        if (value < 0 || value >= (1ll << 30)) return fail(GPS_ERR_INVALID, "epoch out of range");
        c->s.epoch = (uint32_t)value;
    } else if (!strcmp(key, "debug_delay_tile")) {
        // test hook: the CTA of tile (value - 1) of the register-tile kernels starts ~40 ms late (0 = off)
        if (value < 0 || value > 0xFFFFFFFFll) return fail(GPS_ERR_INVALID, "debug_delay_tile out of range");
        c->s.debug_delay_tile = (uint32_t)value;
    } else if (!strcmp(key, "debug_counter")) {
        // test hook: jump the never-reset ticket counter (e.g. just below the 2^32 wrap)
        DeviceGuard g(c->device);
        const uint32_t v = (uint32_t)value;
        GPS_CUDA(cudaDeviceSynchronize());
        GPS_CUDA(cudaMemcpy(c->s.counter, &v, sizeof(v), cudaMemcpyHostToDevice));
        c->s.counter_base = v;
    } else if (!strcmp(key, "tile_order")) {
        // 0 = atomic ticket (default: provably deadlock-free), 1 = blockIdx.x (what
        // cub::DeviceScan relies on: CTAs are dispatched in index order); register / smem paths
        if (value != 0 && value != 1) return fail(GPS_ERR_INVALID, "tile_order must be 0 or 1");
        c->s.flags = (c->s.flags & ~PF_STATIC_TILE) | (value ? PF_STATIC_TILE : 0u);
    } else if (!strcmp(key, "lookback")) {
        // 0 = warp-parallel look-back (default), 1 = single-thread look-back
        if (value != 0 && value != 1) return fail(GPS_ERR_INVALID, "lookback must be 0 or 1");
        c->s.flags = (c->s.flags & ~PF_LOOKBACK_SINGLE) | (value ? PF_LOOKBACK_SINGLE : 0u);
    } else if (!strcmp(key, "desc_protocol")) {
        // 0 = atom.exch + ld.relaxed (default), 1 = st.release.gpu + ld.acquire.gpu,
        // 2 = st.relaxed + ld.relaxed  (see gps_common.cuh for the measured cost of each)
        if (value < 0 || value > 2) return fail(GPS_ERR_INVALID, "desc_protocol must be 0..2");
        const uint32_t mask = PF_PUBLISH_ST_RELEASE | PF_PUBLISH_ST_RELAXED | PF_POLL_ACQUIRE;
        const uint32_t bits = value == 1 ? (PF_PUBLISH_ST_RELEASE | PF_POLL_ACQUIRE)
                            : value == 2 ? PF_PUBLISH_ST_RELAXED : 0u;
        c->s.flags = (c->s.flags & ~mask) | bits;
    } else if (!strcmp(key, "poll_sleep_ns")) {
        if (value < 0 || value > 100000) return fail(GPS_ERR_INVALID, "poll_sleep_ns out of range");
        c->s.poll_sleep_ns = (uint32_t)value;
    } else if (!strcmp(key, "sharded_l2")) {
        // L2 policies of the sharded step: 0 = none, bit 0 = the reduce pass keeps the front of the slice
        // (evict_last) and streams the rest (evict_first), bit 1 = the scan's stores are evict_first,
        // bit 2 = the scan's loads are evict_first
        if (value < 0 || value > 7) return fail(GPS_ERR_INVALID, "sharded_l2 must be 0..7");
        c->sharded_l2 = (int)value;
    } else if (!strcmp(key, "sharded_l2_keep_mb")) {
        if (value < 1 || value > 4096) return fail(GPS_ERR_INVALID, "sharded_l2_keep_mb out of range");
        c->sharded_l2_keep_mb = (int)value;
    } else if (!strcmp(key, "sharded_pdl")) {
        if (value < 0 || value > 2) return fail(GPS_ERR_INVALID, "sharded_pdl must be 0..2");
        c->sharded_pdl = (int)value;
    } else if (!strcmp(key, "l2_hints")) {
        // L2 cache policies on the tile loads / stores of the smem-resident kernels: -1 = auto (default:
        // both for inputs whose output fits the L2, see dispatch), 0 = off, bit 0 = loads evict_first,
        // bit 1 = stores evict_last
        if (value < -1 || value > 3) return fail(GPS_ERR_INVALID, "l2_hints must be -1..3");
        c->l2_hints = (int)value;
    } else if (!strcmp(key, "slot_wait_polls")) {
        // polls tile 0 / the spine may spend on one carry slot before giving up (0 = default 2^26,
        // more than a minute); tests use a small value to exercise the timeout path
        if (value < 0 || value > 0xFFFFFFFFll) return fail(GPS_ERR_INVALID, "slot_wait_polls out of range");
        c->s.slot_poll_limit = (uint32_t)value;
    } else if (!strcmp(key, "host_chunk_keys")) {
        if (value < 4096 || (uint64_t)value > (1ull << 30)) return fail(GPS_ERR_INVALID, "host_chunk_keys out of range");
        if (c->stage[0]) return fail(GPS_ERR_INVALID, "host pipeline already initialised");
        c->host_chunk_keys = (uint64_t)value & ~3ull;
    } else {
        return fail(GPS_ERR_INVALID, "unknown option '%s'", key);
    }
    return GPS_OK;
}

// ---------------------------------------------------------------------------------
// scans
// ---------------------------------------------------------------------------------
int gps_scan(gps_ctx* c, int algo, int kind, int dtype, const void* d_in, void* d_out, uint64_t n,
             uint32_t carry_in_bits, void* d_total_out, void* stream)
{
    CarrySpec cs;
    cs.carry_bits = carry_in_bits;
    cs.total_out = d_total_out;
    return dispatch_scan(c, algo, kind, dtype, d_in, d_out, n, cs, (cudaStream_t)stream);
}

static int scan_wide_impl(gps_ctx* c, int algo, int kind, int dtype, const void* d_in, void* d_out, uint64_t n,
                          const void* h_carry, void* d_total_out, void* stream, const WideFallbackParams* fbp)
{
    if (!c) return fail(GPS_ERR_INVALID, "null ctx");
    if (kind != GPS_INCLUSIVE && kind != GPS_EXCLUSIVE) return fail(GPS_ERR_INVALID, "bad kind %d", kind);
    if (algo != GPS_ALGO_CSDL && algo != GPS_ALGO_RTS && algo != GPS_ALGO_CSDL_FALLBACK)
        return fail(GPS_ERR_INVALID, "gps_scan_wide takes GPS_ALGO_CSDL, GPS_ALGO_CSDL_FALLBACK or GPS_ALGO_RTS");
    const uint64_t es = dtype == GPS_U64 || dtype == GPS_F64 ? 8 : dtype == GPS_U32X4 ? 16
                        : dtype == GPS_U32 || dtype == GPS_F32 ? 4 : 0;
    if (!es) return fail(GPS_ERR_INVALID, "bad dtype %d", dtype);
    if (n > c->max_n * 4 / es) return fail(GPS_ERR_INVALID, "n * element size exceeds ctx capacity");
    if (n && (!d_in || !d_out)) return fail(GPS_ERR_INVALID, "null data pointer");
    if (!aligned16(d_in) || !aligned16(d_out)) return fail(GPS_ERR_ALIGNMENT, "data pointers must be 16-byte aligned");
    DeviceGuard g(c->device);
    cudaStream_t st = (cudaStream_t)stream;
    bool capturing = false;
    if (int rc = enter_stream(c, st, &capturing)) return rc;
    CaptureScope cap;
    if (capturing && algo != GPS_ALGO_RTS)
        if (int rc = cap.begin(c, st, CaptureScope::WIDE)) return rc;
    const bool incl = kind == GPS_INCLUSIVE;
    bool& cfg = c->wide_configured[dtype][incl ? 0 : 1][algo == GPS_ALGO_RTS ? 1 : algo == GPS_ALGO_CSDL_FALLBACK ? 2 : 0];
    switch (dtype) {
        case GPS_U32:  // the blocked-tile kernel at W = 1, kept to compare it with gps_scan's kernels
            return incl ? launch_wide<uint32_t, true>(c, algo, d_in, d_out, n, h_carry, d_total_out, cfg, st, fbp)
                        : launch_wide<uint32_t, false>(c, algo, d_in, d_out, n, h_carry, d_total_out, cfg, st, fbp);
        case GPS_F32:
            return incl ? launch_wide<float, true>(c, algo, d_in, d_out, n, h_carry, d_total_out, cfg, st, fbp)
                        : launch_wide<float, false>(c, algo, d_in, d_out, n, h_carry, d_total_out, cfg, st, fbp);
        case GPS_U64:
            return incl ? launch_wide<uint64_t, true>(c, algo, d_in, d_out, n, h_carry, d_total_out, cfg, st, fbp)
                        : launch_wide<uint64_t, false>(c, algo, d_in, d_out, n, h_carry, d_total_out, cfg, st, fbp);
        case GPS_F64:
            return incl ? launch_wide<double, true>(c, algo, d_in, d_out, n, h_carry, d_total_out, cfg, st, fbp)
                        : launch_wide<double, false>(c, algo, d_in, d_out, n, h_carry, d_total_out, cfg, st, fbp);
        default:
            return incl ? launch_wide<u32x4, true>(c, algo, d_in, d_out, n, h_carry, d_total_out, cfg, st, fbp)
                        : launch_wide<u32x4, false>(c, algo, d_in, d_out, n, h_carry, d_total_out, cfg, st, fbp);
    }
}

int gps_scan_wide(gps_ctx* c, int algo, int kind, int dtype, const void* d_in, void* d_out, uint64_t n,
                  const void* h_carry, void* d_total_out, void* stream)
{
    return scan_wide_impl(c, algo, kind, dtype, d_in, d_out, n, h_carry, d_total_out, stream, nullptr);
}

int gps_scan_wide_emulated_deadlock(gps_ctx* c, int kind, int dtype, const void* d_in, void* d_out, uint64_t n,
                                    const void* h_carry, uint32_t stall_mask, uint32_t max_spin, uint32_t* d_stats,
                                    void* stream)
{
    if (stall_mask == 0) return fail(GPS_ERR_INVALID, "stall_mask 0 would stall every tile: nobody left to fall back");
    if (max_spin == 0) return fail(GPS_ERR_INVALID, "max_spin must be >= 1");
    const WideFallbackParams f{stall_mask, max_spin, d_stats};
    return scan_wide_impl(c, GPS_ALGO_CSDL_FALLBACK, kind, dtype, d_in, d_out, n, h_carry, nullptr, stream, &f);
}

// ---------------------------------------------------------------------------------
// radix sort: the scan's consumer (gps_sort.cuh)
// ---------------------------------------------------------------------------------
uint64_t gps_sort_scratch_bytes(uint64_t n)
{
    return std::max<uint64_t>(1, div_up(n, kSortMinTile)) * kRadix * sizeof(uint64_t) +
           2 * kMaxSortPasses * kRadix * sizeof(uint32_t);
}

int gps_sort_u32(gps_ctx* c, const void* d_keys_in, void* d_keys_out, void* d_keys_tmp, const void* d_vals_in,
                 void* d_vals_out, void* d_vals_tmp, uint64_t n, int end_bit, void* stream)
{
    return gps_sort_keys32(c, GPS_KEY_U32, d_keys_in, d_keys_out, d_keys_tmp, d_vals_in, d_vals_out, d_vals_tmp, n,
                           end_bit, stream);
}

int gps_sort_keys32(gps_ctx* c, int key_type, const void* d_keys_in, void* d_keys_out, void* d_keys_tmp,
                    const void* d_vals_in, void* d_vals_out, void* d_vals_tmp, uint64_t n, int end_bit, void* stream)
{
    return sort_impl<uint32_t>(c, key_type, d_keys_in, d_keys_out, d_keys_tmp, d_vals_in, d_vals_out, d_vals_tmp, n, end_bit,
                               stream, nullptr);
}

int gps_sort_keys32_fallback(gps_ctx* c, int key_type, const void* d_keys_in, void* d_keys_out, void* d_keys_tmp,
                             const void* d_vals_in, void* d_vals_out, void* d_vals_tmp, uint64_t n, int end_bit,
                             uint32_t stall_mask, uint32_t max_spin, uint32_t* d_stats, void* stream)
{
    if (stall_mask == 0) return fail(GPS_ERR_INVALID, "stall_mask 0 would stall every tile: nobody left to fall back");
    if (max_spin == 0) return fail(GPS_ERR_INVALID, "max_spin must be >= 1");
    const SortFallbackParams f{stall_mask, max_spin, d_stats};
    return sort_impl<uint32_t>(c, key_type, d_keys_in, d_keys_out, d_keys_tmp, d_vals_in, d_vals_out, d_vals_tmp, n, end_bit,
                               stream, &f);
}

int gps_sort_keys64(gps_ctx* c, int key_type, const void* d_keys_in, void* d_keys_out, void* d_keys_tmp,
                    const void* d_vals_in, void* d_vals_out, void* d_vals_tmp, uint64_t n, int end_bit, void* stream)
{
    return sort_impl<uint64_t>(c, key_type, d_keys_in, d_keys_out, d_keys_tmp, d_vals_in, d_vals_out, d_vals_tmp, n, end_bit,
                               stream, nullptr);
}

int gps_sort_keys64_fallback(gps_ctx* c, int key_type, const void* d_keys_in, void* d_keys_out, void* d_keys_tmp,
                             const void* d_vals_in, void* d_vals_out, void* d_vals_tmp, uint64_t n, int end_bit,
                             uint32_t stall_mask, uint32_t max_spin, uint32_t* d_stats, void* stream)
{
    if (stall_mask == 0) return fail(GPS_ERR_INVALID, "stall_mask 0 would stall every tile: nobody left to fall back");
    if (max_spin == 0) return fail(GPS_ERR_INVALID, "max_spin must be >= 1");
    const SortFallbackParams f{stall_mask, max_spin, d_stats};
    return sort_impl<uint64_t>(c, key_type, d_keys_in, d_keys_out, d_keys_tmp, d_vals_in, d_vals_out, d_vals_tmp, n, end_bit,
                               stream, &f);
}

int gps_scan_carry_slots(gps_ctx* c, int algo, int kind, int dtype, const void* d_in, void* d_out,
                         uint64_t n, const uint64_t* d_slots, uint32_t n_slots, uint32_t n_carry,
                         uint32_t tag, void* d_total_out, void* stream)
{
    if (n_slots && !d_slots) return fail(GPS_ERR_INVALID, "null slots");
    if (n_carry > n_slots) return fail(GPS_ERR_INVALID, "n_carry > n_slots");
    CarrySpec cs;
    cs.slots = d_slots;
    cs.n_slots = n_slots;
    cs.n_carry = n_carry;
    cs.tag = tag;
    cs.total_out = d_total_out;
    if (!c) return fail(GPS_ERR_INVALID, "null ctx");
    NoPdlScope plain(c->s.flags, c->sharded_pdl);
    // (dispatch_scan re-derives the PF_L2_* bits from l2_hints: pass the sharded step's through it)
    const int saved_hints = c->l2_hints;
    if (c->sharded_l2 & 6) c->l2_hints = 0x100 | ((c->sharded_l2 & 2) ? 4 : 0) | ((c->sharded_l2 & 4) ? 1 : 0);
    const int rc = dispatch_scan(c, algo, kind, dtype, d_in, d_out, n, cs, (cudaStream_t)stream);
    c->l2_hints = saved_hints;
    return rc;
}

int gps_reduce_publish(gps_ctx* c, int dtype, const void* d_in, uint64_t n,
                       uint64_t* const* h_dst_slots, uint32_t n_dst, uint32_t tag,
                       void* d_total_out, void* stream)
{
    if (!c) return fail(GPS_ERR_INVALID, "null ctx");
    if (dtype != GPS_U32 && dtype != GPS_F32) return fail(GPS_ERR_INVALID, "bad dtype %d", dtype);
    if (n > c->max_n) return fail(GPS_ERR_INVALID, "n exceeds ctx capacity");
    if (n && !d_in) return fail(GPS_ERR_INVALID, "null data pointer");
    if (!aligned16(d_in)) return fail(GPS_ERR_ALIGNMENT, "data pointer must be 16-byte aligned");
    if (n_dst && !h_dst_slots) return fail(GPS_ERR_INVALID, "null destination list");
    PublishList pl{};
    int rc = fill_publish(pl, h_dst_slots, n_dst, tag);
    if (rc) return rc;
    DeviceGuard g(c->device);
    cudaStream_t st = (cudaStream_t)stream;
    rc = enter_stream(c, st);  // no epoch / ticket state: replay-safe as it is
    if (rc) return rc;
    NoPdlScope plain(c->s.flags, c->sharded_pdl);
    c->s.reduce_keep_bytes = (n_dst && (c->sharded_l2 & 1)) ? (uint64_t)c->sharded_l2_keep_mb << 20 : 0;
    const cudaError_t e =
        dtype == GPS_U32
            ? launch_reduce<uint32_t, GPS_RTS_NT, GPS_RTS_V>(c->s, d_in, n, d_total_out, pl, st)
            : launch_reduce<float, GPS_RTS_NT, GPS_RTS_V>(c->s, d_in, n, d_total_out, pl, st);
    if (e != cudaSuccess) return fail(GPS_ERR_CUDA, "reduce launch failed: %s", cudaGetErrorString(e));
    return GPS_OK;
}

int gps_reduce(gps_ctx* c, int dtype, const void* d_in, uint64_t n, void* d_total_out, void* stream)
{
    if (!d_total_out) return fail(GPS_ERR_INVALID, "null total_out");
    return gps_reduce_publish(c, dtype, d_in, n, nullptr, 0, 0, d_total_out, stream);
}

int gps_scan_emulated_deadlock(gps_ctx* c, int kind, int dtype, const void* d_in, void* d_out,
                               uint64_t n, uint32_t stall_mask, uint32_t max_spin,
                               uint32_t* d_stats, void* stream)
{
    int rc = check_scan_args(c, kind, dtype, d_in, d_out, n);
    if (rc) return rc;
    if (stall_mask == 0)
        return fail(GPS_ERR_INVALID, "stall_mask 0 would stall every tile: nobody left to fall back");
    if (max_spin == 0) return fail(GPS_ERR_INVALID, "max_spin must be >= 1");
    if (n == 0) return GPS_OK;
    DeviceGuard g(c->device);
    CarrySpec cs;
    FallbackParams f{stall_mask, max_spin, d_stats};
    cudaStream_t st = (cudaStream_t)stream;
    bool capturing = false;
    rc = enter_stream(c, st, &capturing);
    if (rc) return rc;
    CaptureScope cap;
    if (capturing) {
        rc = cap.begin(c, st, CaptureScope::SCAN);
        if (rc) return rc;
    }
    cudaError_t e;
    if (dtype == GPS_U32)
        e = kind == GPS_INCLUSIVE
                ? launch_fallback<uint32_t, true>(c, d_in, d_out, n, cs, f, st)
                : launch_fallback<uint32_t, false>(c, d_in, d_out, n, cs, f, st);
    else
        e = kind == GPS_INCLUSIVE
                ? launch_fallback<float, true>(c, d_in, d_out, n, cs, f, st)
                : launch_fallback<float, false>(c, d_in, d_out, n, cs, f, st);
    if (e != cudaSuccess) return fail(GPS_ERR_CUDA, "fallback scan launch failed: %s", cudaGetErrorString(e));
    return GPS_OK;
}

int gps_scan_estimate_occupancy(gps_ctx* c, int kind, int dtype, const void* d_in, void* d_out, uint64_t n,
                                uint32_t* h_occupancy, void* stream)
{
    int rc = check_scan_args(c, kind, dtype, d_in, d_out, n);
    if (rc) return rc;
    if (!h_occupancy) return fail(GPS_ERR_INVALID, "null occupancy out");
    *h_occupancy = 0;
    if (n == 0) return GPS_OK;
    DeviceGuard g(c->device);
    cudaStream_t st = (cudaStream_t)stream;
    bool capturing = false;
    rc = enter_stream(c, st, &capturing);
    if (rc) return rc;
    if (capturing) return fail(GPS_ERR_UNSUPPORTED, "gps_scan_estimate_occupancy is synchronous: not capturable");
    uint32_t* d_occ = nullptr;
    GPS_CUDA(cudaMalloc(&d_occ, sizeof(uint32_t)));
    cudaError_t e = cudaMemsetAsync(d_occ, 0, sizeof(uint32_t), st);
    CarrySpec cs;
    const FallbackParams f{kNoStall, GPS_FALLBACK_DEFAULT_SPIN, nullptr};
    if (e == cudaSuccess) {
        if (dtype == GPS_U32)
            e = kind == GPS_INCLUSIVE
                    ? launch_csdl_fallback_occ<uint32_t, GPS_REG_NT, GPS_REG_V, true>(c->s, d_in, d_out, n, cs, f, d_occ, st)
                    : launch_csdl_fallback_occ<uint32_t, GPS_REG_NT, GPS_REG_V, false>(c->s, d_in, d_out, n, cs, f, d_occ, st);
        else
            e = kind == GPS_INCLUSIVE
                    ? launch_csdl_fallback_occ<float, GPS_REG_NT, GPS_REG_V, true>(c->s, d_in, d_out, n, cs, f, d_occ, st)
                    : launch_csdl_fallback_occ<float, GPS_REG_NT, GPS_REG_V, false>(c->s, d_in, d_out, n, cs, f, d_occ, st);
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(h_occupancy, d_occ, sizeof(uint32_t), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    cudaFree(d_occ);
    if (e != cudaSuccess) return fail(GPS_ERR_CUDA, "occupancy scan failed: %s", cudaGetErrorString(e));
    return GPS_OK;
}

// ---------------------------------------------------------------------------------
// host-buffer pipeline: H2D(c+1) || scan(c) || D2H(c-1), chunk totals chained on device
// ---------------------------------------------------------------------------------
static int host_pipeline_init(gps_ctx* c)
{
    if (c->stage[0]) return GPS_OK;
    GPS_CUDA(cudaStreamCreateWithFlags(&c->s_h2d, cudaStreamNonBlocking));
    GPS_CUDA(cudaStreamCreateWithFlags(&c->s_cmp, cudaStreamNonBlocking));
    GPS_CUDA(cudaStreamCreateWithFlags(&c->s_d2h, cudaStreamNonBlocking));
    const uint64_t chunk = std::min<uint64_t>(c->host_chunk_keys, std::max<uint64_t>(c->max_n, 4096));
    for (int i = 0; i < kHostBufs; ++i) {
        GPS_CUDA(cudaMalloc(&c->stage[i], (chunk + 4) * 4));
        GPS_CUDA(cudaEventCreateWithFlags(&c->ev_h2d[i], cudaEventDisableTiming));
        GPS_CUDA(cudaEventCreateWithFlags(&c->ev_cmp[i], cudaEventDisableTiming));
        GPS_CUDA(cudaEventCreateWithFlags(&c->ev_d2h[i], cudaEventDisableTiming));
    }
    GPS_CUDA(cudaMalloc(&c->chain_slots, 2 * sizeof(uint64_t)));
    GPS_CUDA(cudaMemset(c->chain_slots, 0, 2 * sizeof(uint64_t)));
    GPS_CUDA(cudaMalloc(&c->chain_total, 16));
    GPS_CUDA(cudaEventCreateWithFlags(&c->ev_total, cudaEventDisableTiming));
    return GPS_OK;
}

int gps_scan_host_async(gps_ctx* c, int algo, int kind, int dtype, const void* h_in, void* h_out,
                        uint64_t n, uint32_t carry_in_bits, void* h_total_out)
{
    if (!c) return fail(GPS_ERR_INVALID, "null ctx");
    if (n && (!h_in || !h_out)) return fail(GPS_ERR_INVALID, "null host pointer");
    if (n > c->max_n) return fail(GPS_ERR_INVALID, "n exceeds ctx capacity");
    DeviceGuard g(c->device);
    int rc = host_pipeline_init(c);
    if (rc) return rc;
    const uint64_t chunk = std::min<uint64_t>(c->host_chunk_keys, std::max<uint64_t>(c->max_n, 4096));
    const uint64_t n_chunks = std::max<uint64_t>(1, div_up(n, chunk));
    const char* src = static_cast<const char*>(h_in);
    char* dst = static_cast<char*>(h_out);
    uint32_t prev_tag = 0;
    for (uint64_t k = 0; k < n_chunks; ++k) {
        const uint64_t seq = c->host_chunk_seq++;  // staging buffers rotate ACROSS calls, so the first
        const int b = (int)(seq % kHostBufs);      // copy-in of a call overlaps the last copy-out of the previous one
        const uint64_t off = k * chunk;
        const uint64_t len = std::min<uint64_t>(chunk, n - off);
        // the staging buffer is free once its previous contents went back to the host
        if (seq >= (uint64_t)kHostBufs) GPS_CUDA(cudaStreamWaitEvent(c->s_h2d, c->ev_d2h[b], 0));
        if (len) GPS_CUDA(cudaMemcpyAsync(c->stage[b], src + off * 4, len * 4, cudaMemcpyHostToDevice, c->s_h2d));
        GPS_CUDA(cudaEventRecord(c->ev_h2d[b], c->s_h2d));
        GPS_CUDA(cudaStreamWaitEvent(c->s_cmp, c->ev_h2d[b], 0));
        CarrySpec cs;
        if (++c->chain_tag == 0) ++c->chain_tag;  // 0 is the "never written" tag
        const uint32_t tag = c->chain_tag;
        if (k == 0) {
            cs.carry_bits = carry_in_bits;
        } else {  // carry = tagged total published by the previous chunk's last tile
            cs.slots = c->chain_slots + ((k - 1) & 1);
            cs.n_slots = cs.n_carry = 1;
            cs.tag = prev_tag;
        }
        prev_tag = tag;
        cs.publish.n = 1;
        cs.publish.dst[0] = c->chain_slots + (k & 1);
        cs.publish.tag = tag;
        if (k + 1 == n_chunks) {
            cs.total_out = c->chain_total;
            // the previous call's copy of chain_total to the host must have been performed
            if (c->ev_total_pending) GPS_CUDA(cudaStreamWaitEvent(c->s_cmp, c->ev_total, 0));
            c->ev_total_pending = false;
        }
        rc = dispatch_scan(c, algo, kind, dtype, c->stage[b], c->stage[b], len, cs, c->s_cmp);
        if (rc) return rc;
        GPS_CUDA(cudaEventRecord(c->ev_cmp[b], c->s_cmp));
        GPS_CUDA(cudaStreamWaitEvent(c->s_d2h, c->ev_cmp[b], 0));
        if (len) GPS_CUDA(cudaMemcpyAsync(dst + off * 4, c->stage[b], len * 4, cudaMemcpyDeviceToHost, c->s_d2h));
        GPS_CUDA(cudaEventRecord(c->ev_d2h[b], c->s_d2h));
    }
    if (h_total_out) {  // queued behind the last chunk's copy-out, which already waits for its scan
        GPS_CUDA(cudaMemcpyAsync(h_total_out, c->chain_total, 4, cudaMemcpyDeviceToHost, c->s_d2h));
        GPS_CUDA(cudaEventRecord(c->ev_total, c->s_d2h));
        c->ev_total_pending = true;
    }
    return GPS_OK;
}

int gps_host_sync(gps_ctx* c)
{
    if (!c) return fail(GPS_ERR_INVALID, "null ctx");
    if (!c->stage[0]) return GPS_OK;
    DeviceGuard g(c->device);
    GPS_CUDA(cudaStreamSynchronize(c->s_h2d));
    GPS_CUDA(cudaStreamSynchronize(c->s_cmp));
    GPS_CUDA(cudaStreamSynchronize(c->s_d2h));
    return GPS_OK;
}

int gps_scan_host(gps_ctx* c, int algo, int kind, int dtype, const void* h_in, void* h_out,
                  uint64_t n, uint32_t carry_in_bits, void* h_total_out)
{
    const int rc = gps_scan_host_async(c, algo, kind, dtype, h_in, h_out, n, carry_in_bits, h_total_out);
    return rc ? rc : gps_host_sync(c);
}


// ---------------------------------------------------------------------------------
// harness kernels
// ---------------------------------------------------------------------------------
static inline int util_grid(uint64_t n)
{
    return (int)std::max<uint64_t>(1, std::min<uint64_t>(div_up(n, 256 * 8), 148 * 16));
}

int gps_init_ones(void* d, uint64_t n, int dtype, void* stream)
{
    if (n && !d) return fail(GPS_ERR_INVALID, "null pointer");
    if (!n) return GPS_OK;
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == GPS_U32)
        init_ones_kernel<uint32_t><<<util_grid(n), 256, 0, st>>>(static_cast<uint32_t*>(d), n);
    else if (dtype == GPS_F32)
        init_ones_kernel<float><<<util_grid(n), 256, 0, st>>>(static_cast<float*>(d), n);
    else
        return fail(GPS_ERR_INVALID, "bad dtype %d", dtype);
    GPS_CUDA(cudaGetLastError());
    return GPS_OK;
}

int gps_init_random(void* d, uint64_t n, int dtype, uint64_t seed, uint32_t mask, uint64_t first_index,
                    void* stream)
{
    if (n && !d) return fail(GPS_ERR_INVALID, "null pointer");
    if (!n) return GPS_OK;
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == GPS_U32)
        init_random_u32_kernel<<<util_grid(n), 256, 0, st>>>(static_cast<uint32_t*>(d), n, seed, mask, first_index);
    else if (dtype == GPS_F32)
        init_random_f32_kernel<<<util_grid(n), 256, 0, st>>>(static_cast<float*>(d), n, seed, first_index);
    else
        return fail(GPS_ERR_INVALID, "bad dtype %d", dtype);
    GPS_CUDA(cudaGetLastError());
    return GPS_OK;
}

int gps_validate_ones(const void* d, uint64_t n, int kind, uint32_t first_value, uint64_t* h_errs)
{
    if (!h_errs || (n && !d)) return fail(GPS_ERR_INVALID, "null pointer");
    unsigned long long* d_err = nullptr;
    GPS_CUDA(cudaMalloc(&d_err, sizeof(unsigned long long)));
    GPS_CUDA(cudaMemset(d_err, 0, sizeof(unsigned long long)));
    if (n)
        validate_ones_kernel<<<util_grid(n), 256>>>(static_cast<const uint32_t*>(d), n, first_value,
                                                    kind == GPS_INCLUSIVE ? 1u : 0u, d_err);
    unsigned long long h = 0;
    cudaError_t e = cudaMemcpy(&h, d_err, sizeof(h), cudaMemcpyDeviceToHost);
    cudaFree(d_err);
    GPS_CUDA(e);
    *h_errs = h;
    return GPS_OK;
}

int gps_validate_ones_wide(const void* d, uint64_t n, int kind, int dtype, uint64_t* h_errs)
{
    if (!h_errs || (n && !d)) return fail(GPS_ERR_INVALID, "null pointer");
    if (dtype != GPS_U64 && dtype != GPS_U32X4) return fail(GPS_ERR_INVALID, "GPS_U64 or GPS_U32X4");
    unsigned long long* d_err = nullptr;
    GPS_CUDA(cudaMalloc(&d_err, sizeof(unsigned long long)));
    GPS_CUDA(cudaMemset(d_err, 0, sizeof(unsigned long long)));
    const uint32_t inc = kind == GPS_INCLUSIVE ? 1u : 0u;
    if (n && dtype == GPS_U64)
        validate_ones_u64_kernel<<<util_grid(n), 256>>>(static_cast<const uint64_t*>(d), n, inc, d_err);
    else if (n)
        validate_struct_ones_kernel<<<util_grid(4 * n), 256>>>(static_cast<const uint32_t*>(d), 4 * n, inc, d_err);
    unsigned long long h = 0;
    cudaError_t e = cudaMemcpy(&h, d_err, sizeof(h), cudaMemcpyDeviceToHost);
    cudaFree(d_err);
    GPS_CUDA(e);
    *h_errs = h;
    return GPS_OK;
}

static int sort_fingerprint_impl(bool keys64, const void* d_keys, uint64_t n, uint64_t* h_inversions, uint64_t h_sums[2])
{
    if (!h_inversions || !h_sums || (n && !d_keys)) return fail(GPS_ERR_INVALID, "null pointer");
    unsigned long long* d_err = nullptr;
    GPS_CUDA(cudaMalloc(&d_err, 3 * sizeof(unsigned long long)));
    GPS_CUDA(cudaMemset(d_err, 0, 3 * sizeof(unsigned long long)));
    if (n && keys64) sort_fingerprint_kernel<uint64_t><<<util_grid(n), 256>>>(static_cast<const uint64_t*>(d_keys), n, d_err);
    if (n && !keys64) sort_fingerprint_kernel<uint32_t><<<util_grid(n), 256>>>(static_cast<const uint32_t*>(d_keys), n, d_err);
    unsigned long long h[3] = {0, 0, 0};
    cudaError_t e = cudaMemcpy(h, d_err, sizeof(h), cudaMemcpyDeviceToHost);
    cudaFree(d_err);
    GPS_CUDA(e);
    *h_inversions = h[0];
    h_sums[0] = h[1];
    h_sums[1] = h[2];
    return GPS_OK;
}

int gps_sort_fingerprint(const void* d_keys, uint64_t n, uint64_t* h_inversions, uint64_t h_sums[2])
{
    return sort_fingerprint_impl(false, d_keys, n, h_inversions, h_sums);
}

int gps_sort_fingerprint64(const void* d_keys, uint64_t n, uint64_t* h_inversions, uint64_t h_sums[2])
{
    return sort_fingerprint_impl(true, d_keys, n, h_inversions, h_sums);
}

int gps_compare_u32(const void* d_a, const void* d_b, uint64_t n, uint64_t* h_errs, uint64_t* h_first_bad)
{
    if (!h_errs || (n && (!d_a || !d_b))) return fail(GPS_ERR_INVALID, "null pointer");
    unsigned long long* d_err = nullptr;
    unsigned long long init[2] = {0ull, ~0ull};
    GPS_CUDA(cudaMalloc(&d_err, sizeof(init)));
    GPS_CUDA(cudaMemcpy(d_err, init, sizeof(init), cudaMemcpyHostToDevice));
    if (n)
        compare_u32_kernel<<<util_grid(n), 256>>>(static_cast<const uint32_t*>(d_a),
                                                  static_cast<const uint32_t*>(d_b), n, d_err);
    unsigned long long h[2] = {0, 0};
    cudaError_t e = cudaMemcpy(h, d_err, sizeof(h), cudaMemcpyDeviceToHost);
    cudaFree(d_err);
    GPS_CUDA(e);
    *h_errs = h[0];
    if (h_first_bad) *h_first_bad = h[0] ? h[1] : n;
    return GPS_OK;
}

int gps_compare_f32(const void* d_a, const void* d_b, const void* d_bound, uint64_t n, float rtol,
                    float atol, uint64_t* h_errs, float* h_max_abs_diff)
{
    if (!h_errs || (n && (!d_a || !d_b))) return fail(GPS_ERR_INVALID, "null pointer");
    unsigned long long* d_err = nullptr;
    GPS_CUDA(cudaMalloc(&d_err, 16));
    GPS_CUDA(cudaMemset(d_err, 0, 16));
    if (n)
        compare_f32_kernel<<<util_grid(n), 256>>>(static_cast<const float*>(d_a),
                                                  static_cast<const float*>(d_b),
                                                  static_cast<const float*>(d_bound), n, rtol, atol,
                                                  d_err, reinterpret_cast<uint32_t*>(d_err + 1));
    unsigned long long h[2] = {0, 0};
    cudaError_t e = cudaMemcpy(h, d_err, sizeof(h), cudaMemcpyDeviceToHost);
    cudaFree(d_err);
    GPS_CUDA(e);
    *h_errs = h[0];
    if (h_max_abs_diff) {
        uint32_t bits = (uint32_t)h[1];
        memcpy(h_max_abs_diff, &bits, 4);
    }
    return GPS_OK;
}

// ---------------------------------------------------------------------------------
// memory / stream / timing helpers
// ---------------------------------------------------------------------------------
int gps_malloc(void** d_ptr, uint64_t bytes)
{
    if (!d_ptr) return fail(GPS_ERR_INVALID, "null pointer");
    GPS_CUDA(cudaMalloc(d_ptr, bytes ? bytes : 16));
    return GPS_OK;
}
int gps_free(void* d_ptr)
{
    GPS_CUDA(cudaFree(d_ptr));
    return GPS_OK;
}
int gps_malloc_host(void** h_ptr, uint64_t bytes)
{
    if (!h_ptr) return fail(GPS_ERR_INVALID, "null pointer");
    GPS_CUDA(cudaMallocHost(h_ptr, bytes ? bytes : 16));
    return GPS_OK;
}
int gps_free_host(void* h_ptr)
{
    GPS_CUDA(cudaFreeHost(h_ptr));
    return GPS_OK;
}
int gps_memcpy_h2d(void* d, const void* h, uint64_t bytes, void* stream)
{
    GPS_CUDA(cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, (cudaStream_t)stream));
    return GPS_OK;
}
int gps_memcpy_d2h(void* h, const void* d, uint64_t bytes, void* stream)
{
    GPS_CUDA(cudaMemcpyAsync(h, d, bytes, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    return GPS_OK;
}
int gps_memcpy_d2d(void* dst, const void* src, uint64_t bytes, void* stream)
{
    GPS_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    return GPS_OK;
}
int gps_memset(void* d, int byte, uint64_t bytes, void* stream)
{
    GPS_CUDA(cudaMemsetAsync(d, byte, bytes, (cudaStream_t)stream));
    return GPS_OK;
}
int gps_stream_sync(void* stream)
{
    GPS_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    return GPS_OK;
}
int gps_device_sync(void)
{
    GPS_CUDA(cudaDeviceSynchronize());
    return GPS_OK;
}
int gps_event_create(void** ev)
{
    if (!ev) return fail(GPS_ERR_INVALID, "null pointer");
    cudaEvent_t e;
    GPS_CUDA(cudaEventCreate(&e));
    *ev = e;
    return GPS_OK;
}
int gps_event_destroy(void* ev)
{
    GPS_CUDA(cudaEventDestroy((cudaEvent_t)ev));
    return GPS_OK;
}
int gps_event_record(void* ev, void* stream)
{
    GPS_CUDA(cudaEventRecord((cudaEvent_t)ev, (cudaStream_t)stream));
    return GPS_OK;
}
int gps_event_elapsed_ms(void* start, void* stop, float* ms)
{
    if (!ms) return fail(GPS_ERR_INVALID, "null pointer");
    GPS_CUDA(cudaEventSynchronize((cudaEvent_t)stop));
    GPS_CUDA(cudaEventElapsedTime(ms, (cudaEvent_t)start, (cudaEvent_t)stop));
    return GPS_OK;
}

// ---------------------------------------------------------------------------------
// CUDA IPC
// ---------------------------------------------------------------------------------
int gps_ipc_export(void* d_ptr, unsigned char handle[GPS_IPC_HANDLE_BYTES])
{
    static_assert(sizeof(cudaIpcMemHandle_t) == GPS_IPC_HANDLE_BYTES, "IPC handle size");
    cudaIpcMemHandle_t h;
    GPS_CUDA(cudaIpcGetMemHandle(&h, d_ptr));
    memcpy(handle, &h, sizeof(h));
    return GPS_OK;
}
int gps_ipc_open(const unsigned char handle[GPS_IPC_HANDLE_BYTES], void** d_ptr)
{
    if (!d_ptr) return fail(GPS_ERR_INVALID, "null pointer");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof(h));
    GPS_CUDA(cudaIpcOpenMemHandle(d_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return GPS_OK;
}
int gps_ipc_close(void* d_ptr)
{
    GPS_CUDA(cudaIpcCloseMemHandle(d_ptr));
    return GPS_OK;
}

}  // extern "C"

// ---------------------------------------------------------------------------------
// single-process sharded scan
// ---------------------------------------------------------------------------------
struct gps_shard_group {
    int n = 0;
    std::vector<int> devices;
    std::vector<gps_ctx*> ctx;
    std::vector<cudaStream_t> stream;
    std::vector<uint64_t*> slots;   // per GPU: 2 parities x n slots
    std::vector<cudaEvent_t> ev0, ev1;
    uint32_t tag = 0;
    // host-buffer pipeline (gps_sharded_scan_host*): chunk q of the stream of chunks goes to GPU
    // q % n; its scan publishes the running total into slot [(q / n) & 1] of GPU (q + 1) % n
    uint64_t host_seq = 0;          // chunks issued so far, over all calls
    uint32_t host_tag = 0;
    std::vector<int> io;               // GPUs (indices into devices) that carry the host traffic, in chain order
    std::vector<cudaEvent_t> ev_call;  // per GPU; [ev_call_idx] = last scan of the previous call
    int ev_call_idx = -1;              // (orders call boundaries), -1 = none yet
};

namespace {
constexpr int kHostChainSlot0 = 2 * kMaxPublish;  // two more words behind the 2 x kMaxPublish carry slots

// status words of every ctx of the group after its work has drained: a tripped look-back or
// slot watchdog means wrong results -- report it, clear it and bring the ticket base back in step
int sharded_status(gps_shard_group* g)
{
    int rc = GPS_OK;
    for (int i = 0; i < g->n; ++i) {
        gps_ctx* c = g->ctx[i];
        cudaSetDevice(g->devices[i]);
        uint32_t st[2] = {0, 0};  // {counter, status} are adjacent
        if (cudaMemcpy(st, c->s.counter, sizeof(st), cudaMemcpyDeviceToHost) != cudaSuccess) {
            rc = fail(GPS_ERR_CUDA, "device %d: status read failed: %s", g->devices[i], cudaGetErrorString(cudaGetLastError()));
            continue;
        }
        if (st[1] & STATUS_RESCUED) {  // informational (a late tile was rescued; results are valid), as in gps_check
            c->rescued = 1;
            st[1] &= ~STATUS_RESCUED;
            if (!st[1]) cudaMemset(c->s.status, 0, sizeof(uint32_t));
        }
        if (st[1]) {
            cudaMemset(c->s.status, 0, sizeof(uint32_t));
            c->s.counter_base = st[0];
            rc = fail(GPS_ERR_WATCHDOG, "device %d: %s (status 0x%x): results invalid", g->devices[i],
                      (st[1] & STATUS_SLOT_TIMEOUT) ? "a carry slot never arrived" : "look-back poll budget exceeded", st[1]);
        }
    }
    return rc;
}
}  // namespace

extern "C" {

int gps_sharded_create(gps_shard_group** out, int n_gpus, const int* devices, uint64_t max_n_per_gpu)
{
    if (!out || !devices || n_gpus < 1 || n_gpus > kMaxPublish)
        return fail(GPS_ERR_INVALID, "bad sharded_create arguments");
    *out = nullptr;
    gps_shard_group* g = new gps_shard_group();
    g->n = n_gpus;
    g->devices.assign(devices, devices + n_gpus);
    g->ctx.assign(n_gpus, nullptr);
    g->stream.assign(n_gpus, nullptr);
    g->slots.assign(n_gpus, nullptr);
    g->ev0.assign(n_gpus, nullptr);
    g->ev1.assign(n_gpus, nullptr);
    g->ev_call.assign(n_gpus, nullptr);
    for (int i = 0; i < n_gpus; ++i) g->io.push_back(i);  // default: every GPU takes chunks
    int prev = 0;
    cudaGetDevice(&prev);
    int rc = GPS_OK;
    for (int i = 0; i < n_gpus && rc == GPS_OK; ++i) {
        rc = gps_create(&g->ctx[i], devices[i], max_n_per_gpu);
        if (rc) break;
        cudaError_t e = cudaSetDevice(devices[i]);
        for (int j = 0; j < n_gpus && e == cudaSuccess; ++j) {
            if (devices[j] == devices[i]) continue;
            int can = 0;
            e = cudaDeviceCanAccessPeer(&can, devices[i], devices[j]);
            if (e == cudaSuccess && !can) {
                rc = fail(GPS_ERR_UNSUPPORTED, "no peer access %d -> %d", devices[i], devices[j]);
                break;
            }
            if (e == cudaSuccess) {
                e = cudaDeviceEnablePeerAccess(devices[j], 0);
                if (e == cudaErrorPeerAccessAlreadyEnabled) {
                    e = cudaSuccess;
                    cudaGetLastError();
                }
            }
        }
        if (rc) break;
        if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&g->stream[i], cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaMalloc(&g->slots[i], (2 * kMaxPublish + 2) * sizeof(uint64_t));
        if (e == cudaSuccess) e = cudaMemset(g->slots[i], 0, (2 * kMaxPublish + 2) * sizeof(uint64_t));
        if (e == cudaSuccess) e = cudaEventCreate(&g->ev0[i]);
        if (e == cudaSuccess) e = cudaEventCreate(&g->ev1[i]);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&g->ev_call[i], cudaEventDisableTiming);
        if (e != cudaSuccess) rc = fail(GPS_ERR_CUDA, "sharded setup on device %d: %s", devices[i], cudaGetErrorString(e));
    }
    cudaSetDevice(prev);
    if (rc) {
        gps_sharded_destroy(g);
        return rc;
    }
    *out = g;
    return GPS_OK;
}

void gps_sharded_destroy(gps_shard_group* g)
{
    if (!g) return;
    int prev = 0;
    cudaGetDevice(&prev);
    for (int i = 0; i < g->n; ++i) {
        cudaSetDevice(g->devices[i]);
        if (g->ev_call[i]) cudaEventDestroy(g->ev_call[i]);
        if (g->stream[i]) cudaStreamDestroy(g->stream[i]);
        if (g->slots[i]) cudaFree(g->slots[i]);
        if (g->ev0[i]) cudaEventDestroy(g->ev0[i]);
        if (g->ev1[i]) cudaEventDestroy(g->ev1[i]);
        if (g->ctx[i]) gps_destroy(g->ctx[i]);
    }
    cudaSetDevice(prev);
    delete g;
}

int gps_sharded_scan(gps_shard_group* g, int algo, int kind, int dtype, const void* const* d_in,
                     void* const* d_out, const uint64_t* n_per_gpu, int carry_mode, float* h_ms)
{
    if (!g || !d_in || !d_out || !n_per_gpu) return fail(GPS_ERR_INVALID, "null argument");
    if (carry_mode != 0 && carry_mode != 1) return fail(GPS_ERR_INVALID, "bad carry_mode");
    const int G = g->n;
    int prev = 0;
    cudaGetDevice(&prev);
    const uint32_t tag = ++g->tag ? g->tag : ++g->tag;
    const uint32_t par = tag & 1u;
    int rc = GPS_OK;
    const auto t0 = std::chrono::steady_clock::now();
    // pass 1: reduce every slice; publish its tagged total into slot [par][i] of EVERY
    // GPU (P2P stores over NVLink).  Every GPU later polls all G slots, so nobody can
    // run a whole step ahead and overwrite a slot that is still being read.
    std::vector<void*> totals(G, nullptr);
    for (int i = 0; i < G && !rc; ++i) {
        cudaSetDevice(g->devices[i]);
        cudaEventRecord(g->ev0[i], g->stream[i]);
        uint64_t* dst[kMaxPublish];
        uint32_t n_dst = 0;
        if (carry_mode == 0)
            for (int j = 0; j < G; ++j) dst[n_dst++] = g->slots[j] + par * kMaxPublish + i;
        else
            dst[n_dst++] = g->slots[i] + par * kMaxPublish + i;
        rc = gps_reduce_publish(g->ctx[i], dtype, d_in[i], n_per_gpu[i], dst, n_dst, tag, nullptr, g->stream[i]);
    }
    if (!rc && carry_mode == 1) {  // host-staged exchange (debug / no-P2P boxes)
        std::vector<uint64_t> words(G, 0);
        for (int i = 0; i < G; ++i) {
            cudaSetDevice(g->devices[i]);
            cudaStreamSynchronize(g->stream[i]);
            cudaMemcpy(&words[i], g->slots[i] + par * kMaxPublish + i, 8, cudaMemcpyDeviceToHost);
        }
        for (int i = 0; i < G; ++i) {
            cudaSetDevice(g->devices[i]);
            cudaMemcpyAsync(g->slots[i] + par * kMaxPublish, words.data(), 8 * G, cudaMemcpyHostToDevice, g->stream[i]);
            cudaStreamSynchronize(g->stream[i]);
        }
    }
    // pass 2: scan with the carry folded into tile 0's look-back
    for (int i = 0; i < G && !rc; ++i) {
        cudaSetDevice(g->devices[i]);
        rc = gps_scan_carry_slots(g->ctx[i], algo, kind, dtype, d_in[i], d_out[i], n_per_gpu[i],
                                  g->slots[i] + par * kMaxPublish, (uint32_t)G, (uint32_t)i, tag,
                                  nullptr, g->stream[i]);
        cudaEventRecord(g->ev1[i], g->stream[i]);
    }
    float worst = 0.f;
    for (int i = 0; i < G; ++i) {
        cudaSetDevice(g->devices[i]);
        cudaError_t e = cudaStreamSynchronize(g->stream[i]);
        if (e != cudaSuccess && !rc) rc = fail(GPS_ERR_CUDA, "device %d: %s", g->devices[i], cudaGetErrorString(e));
        float ms = 0.f;
        if (e == cudaSuccess && cudaEventElapsedTime(&ms, g->ev0[i], g->ev1[i]) == cudaSuccess)
            worst = std::max(worst, ms);
    }
    const auto t1 = std::chrono::steady_clock::now();
    (void)t0;
    (void)t1;
    if (h_ms) *h_ms = worst;
    if (!rc) rc = sharded_status(g);  // a carry slot that never arrived must not pass silently
    cudaSetDevice(prev);
    return rc;
}

gps_ctx* gps_sharded_ctx(gps_shard_group* g, int index)
{
    return (g && index >= 0 && index < g->n) ? g->ctx[index] : nullptr;
}

// One host array, all GPUs of the group, one call.  The array is cut into chunks of
// "host_chunk_keys"; chunk q of the running sequence goes to GPU q % G: copy-in, scan, copy-out on
// that GPU's three streams over its rotating staging buffers, exactly as gps_scan_host does on one
// GPU.  The scan of chunk q seeds itself from the tagged running total its predecessor's last tile
// stored into THIS GPU's slot over NVLink (st.relaxed.sys -> ld.relaxed.sys in tile 0's carry
// acquisition) and publishes its own inclusive total to the next GPU's slot: the chunks' scans
// form one chain across the GPUs (~20 us per 64 MiB link) while G copy-ins and G copy-outs run
// concurrently -- 8 B/key of HBM traffic per key, no reduce pre-pass, no host synchronisation and
// no collective.  (The device-resident sharded scan needs its reduce pass because all slices are
// already there and must be scanned at once; here PCIe is three orders of magnitude slower than
// the chain.)
int gps_sharded_scan_host_async(gps_shard_group* g, int algo, int kind, int dtype, const void* h_in,
                                void* h_out, uint64_t n, uint32_t carry_in_bits, void* h_total_out)
{
    if (!g) return fail(GPS_ERR_INVALID, "null group");
    if (n && (!h_in || !h_out)) return fail(GPS_ERR_INVALID, "null host pointer");
    if (kind != GPS_INCLUSIVE && kind != GPS_EXCLUSIVE) return fail(GPS_ERR_INVALID, "bad kind %d", kind);
    if (dtype != GPS_U32 && dtype != GPS_F32) return fail(GPS_ERR_INVALID, "bad dtype %d", dtype);
    const int G = (int)g->io.size();  // GPUs that take chunks (all of them unless calibrated / set otherwise)
    int prev = 0;
    cudaGetDevice(&prev);
    struct Restore {
        int dev;
        ~Restore() { cudaSetDevice(dev); }
    } restore{prev};
    uint64_t chunk = ~0ull;
    for (int j = 0; j < G; ++j) {
        const int i = g->io[j];
        gps_ctx* c = g->ctx[i];
        GPS_CUDA(cudaSetDevice(g->devices[i]));
        if (int rc = host_pipeline_init(c)) return rc;
        chunk = std::min(chunk, std::min<uint64_t>(c->host_chunk_keys, std::max<uint64_t>(c->max_n, 4096)));
    }
    const uint64_t n_chunks = std::max<uint64_t>(1, div_up(n, chunk));
    const char* src = static_cast<const char*>(h_in);
    char* dst = static_cast<char*>(h_out);
    uint32_t prev_tag = 0;
    for (uint64_t k = 0; k < n_chunks; ++k) {
        const uint64_t q = g->host_seq++;
        const int gi = g->io[q % (uint64_t)G];
        gps_ctx* c = g->ctx[gi];
        GPS_CUDA(cudaSetDevice(g->devices[gi]));
        const uint64_t seq = c->host_chunk_seq++;
        const int b = (int)(seq % kHostBufs);
        const uint64_t off = k * chunk;
        const uint64_t len = std::min<uint64_t>(chunk, n - off);
        if (seq >= (uint64_t)kHostBufs) GPS_CUDA(cudaStreamWaitEvent(c->s_h2d, c->ev_d2h[b], 0));
        if (len) GPS_CUDA(cudaMemcpyAsync(c->stage[b], src + off * 4, len * 4, cudaMemcpyHostToDevice, c->s_h2d));
        GPS_CUDA(cudaEventRecord(c->ev_h2d[b], c->s_h2d));
        GPS_CUDA(cudaStreamWaitEvent(c->s_cmp, c->ev_h2d[b], 0));
        CarrySpec cs;
        if (++g->host_tag == 0) ++g->host_tag;  // 0 is the "never written" tag
        const uint32_t tag = g->host_tag;
        if (k == 0) {
            cs.carry_bits = carry_in_bits;
            // call boundary: the chain restarts here, so order it behind the previous call's last scan
            // (a slot is only reused once its previous reader has run, which the chain guarantees
            // inside a call and this event guarantees across calls)
            if (g->ev_call_idx >= 0) GPS_CUDA(cudaStreamWaitEvent(c->s_cmp, g->ev_call[g->ev_call_idx], 0));
        } else {  // the slot the previous chunk's scan publishes into, on THIS GPU
            cs.slots = g->slots[gi] + kHostChainSlot0 + (((q - 1) / (uint64_t)G) & 1u);
            cs.n_slots = cs.n_carry = 1;
            cs.tag = prev_tag;
            // two "GPUs" of the group that are one physical device (tests on a single-GPU box): a
            // scan spinning on its slot could keep its predecessor from ever becoming resident,
            // so there the chain is also a stream dependency
            const int pi = g->io[(q - 1) % (uint64_t)G];
            if (pi != gi && g->devices[pi] == g->devices[gi])
                GPS_CUDA(cudaStreamWaitEvent(c->s_cmp, g->ctx[pi]->ev_cmp[(g->ctx[pi]->host_chunk_seq - 1) % kHostBufs], 0));
        }
        prev_tag = tag;
        if (k + 1 < n_chunks) {
            cs.publish.n = 1;
            cs.publish.dst[0] = g->slots[g->io[(q + 1) % (uint64_t)G]] + kHostChainSlot0 + ((q / (uint64_t)G) & 1u);
            cs.publish.tag = tag;
        } else {
            cs.total_out = c->chain_total;
            if (c->ev_total_pending) GPS_CUDA(cudaStreamWaitEvent(c->s_cmp, c->ev_total, 0));
            c->ev_total_pending = false;
        }
        if (int rc = dispatch_scan(c, algo, kind, dtype, c->stage[b], c->stage[b], len, cs, c->s_cmp)) return rc;
        GPS_CUDA(cudaEventRecord(c->ev_cmp[b], c->s_cmp));
        GPS_CUDA(cudaStreamWaitEvent(c->s_d2h, c->ev_cmp[b], 0));
        if (len) GPS_CUDA(cudaMemcpyAsync(dst + off * 4, c->stage[b], len * 4, cudaMemcpyDeviceToHost, c->s_d2h));
        GPS_CUDA(cudaEventRecord(c->ev_d2h[b], c->s_d2h));
        if (k + 1 == n_chunks) {
            if (h_total_out) {
                GPS_CUDA(cudaMemcpyAsync(h_total_out, c->chain_total, 4, cudaMemcpyDeviceToHost, c->s_d2h));
                GPS_CUDA(cudaEventRecord(c->ev_total, c->s_d2h));
                c->ev_total_pending = true;
            }
            GPS_CUDA(cudaEventRecord(g->ev_call[gi], c->s_cmp));
            g->ev_call_idx = gi;
        }
    }
    return GPS_OK;
}

namespace {
// memcpy-only trial: the GPUs of `set` copy `per_gpu` bytes in from h_src and out to h_dst at the
// same time (pieces of the pipeline's chunk size, both directions concurrently, each GPU on its
// own section of the host buffers); returns GB/s per direction over the whole set, <= 0 on error
double host_io_trial(gps_shard_group* g, const std::vector<int>& set, const char* h_src, char* h_dst, uint64_t per_gpu,
                     uint64_t piece)
{
    const auto t0 = std::chrono::steady_clock::now();
    for (size_t j = 0; j < set.size(); ++j) {
        gps_ctx* c = g->ctx[set[j]];
        if (cudaSetDevice(g->devices[set[j]]) != cudaSuccess) return -1.0;
        for (uint64_t off = 0; off < per_gpu; off += piece) {
            const uint64_t len = std::min(piece, per_gpu - off);
            if (cudaMemcpyAsync(c->stage[0], h_src + j * per_gpu + off, len, cudaMemcpyHostToDevice, c->s_h2d) != cudaSuccess ||
                cudaMemcpyAsync(h_dst + j * per_gpu + off, c->stage[1], len, cudaMemcpyDeviceToHost, c->s_d2h) != cudaSuccess)
                return -1.0;
        }
    }
    for (size_t j = 0; j < set.size(); ++j) {
        gps_ctx* c = g->ctx[set[j]];
        cudaSetDevice(g->devices[set[j]]);
        if (cudaStreamSynchronize(c->s_h2d) != cudaSuccess || cudaStreamSynchronize(c->s_d2h) != cudaSuccess) return -1.0;
    }
    const double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    return (double)per_gpu * set.size() / s / 1e9;
}
}  // namespace

// Which GPUs should carry the host traffic?  On a multi-socket box the host <-> device paths are
// not equal and they interfere: on the 8 x B200 VM this was developed on, all eight GPUs copying
// both ways reach 72 GB/s per direction in total, GPUs 4-7 alone 95, GPUs 0-3 alone 53
// (profiles/r02_pcie_aggregate.json, r02_pcie_combos.jsonl) -- the scan itself needs a fraction of
// one GPU.  This measures, memcpy only, on the caller's own buffers: the whole group, then its two
// halves, descending into a half while it beats its parent by more than 3 %, and routes the chunks
// of gps_sharded_scan_host through the winner.  h_dst is overwritten with junk (pass the output
// array of the scan that follows).  Synchronous; a few hundred milliseconds.
int gps_sharded_calibrate_host(gps_shard_group* g, const void* h_src, void* h_dst, uint64_t bytes, float* h_gbs_each_way)
{
    if (!g || !h_src || !h_dst) return fail(GPS_ERR_INVALID, "null argument");
    int prev = 0;
    cudaGetDevice(&prev);
    struct Restore {
        int dev;
        ~Restore() { cudaSetDevice(dev); }
    } restore{prev};
    uint64_t piece = ~0ull;
    for (int i = 0; i < g->n; ++i) {
        GPS_CUDA(cudaSetDevice(g->devices[i]));
        if (int rc = host_pipeline_init(g->ctx[i])) return rc;
        if (int rc = gps_host_sync(g->ctx[i])) return rc;
        piece = std::min(piece, 4 * std::min<uint64_t>(g->ctx[i]->host_chunk_keys, std::max<uint64_t>(g->ctx[i]->max_n, 4096)));
    }
    std::vector<int> best(g->n);
    for (int i = 0; i < g->n; ++i) best[i] = i;
    const uint64_t volume = std::min<uint64_t>(bytes, 4ull << 30);  // per trial and direction
    if (volume < 8 * piece) {  // too small to tell anything: keep every GPU
        g->io = best;
        if (h_gbs_each_way) *h_gbs_each_way = 0.f;
        return GPS_OK;
    }
    auto measure = [&](const std::vector<int>& set) {
        const uint64_t per_gpu = (volume / set.size()) & ~(uint64_t)255;
        double r = 0;
        for (int rep = 0; rep < 2; ++rep)
            r = std::max(r, host_io_trial(g, set, static_cast<const char*>(h_src), static_cast<char*>(h_dst), per_gpu, piece));
        return r;
    };
    double best_rate = measure(best);
    if (best_rate <= 0) return fail(GPS_ERR_CUDA, "host I/O calibration: copy failed: %s", cudaGetErrorString(cudaGetLastError()));
    while (best.size() >= 2) {
        const std::vector<int> lo(best.begin(), best.begin() + best.size() / 2), hi(best.begin() + best.size() / 2, best.end());
        const double r_lo = measure(lo), r_hi = measure(hi);
        const std::vector<int>& half = r_hi > r_lo ? hi : lo;
        const double r_half = std::max(r_lo, r_hi);
        if (r_half <= 1.03 * best_rate) break;
        best = half;
        best_rate = r_half;
    }
    g->io = best;
    if (h_gbs_each_way) *h_gbs_each_way = (float)best_rate;
    return GPS_OK;
}

int gps_sharded_set_host_io(gps_shard_group* g, const int* gpu_indices, int n)
{
    if (!g || !gpu_indices || n < 1 || n > g->n) return fail(GPS_ERR_INVALID, "bad host I/O set");
    std::vector<int> io(gpu_indices, gpu_indices + n);
    for (int i = 0; i < n; ++i) {
        if (io[i] < 0 || io[i] >= g->n) return fail(GPS_ERR_INVALID, "host I/O index %d out of range", io[i]);
        for (int j = 0; j < i; ++j)
            if (io[j] == io[i]) return fail(GPS_ERR_INVALID, "host I/O index %d twice", io[i]);
    }
    if (int rc = gps_sharded_host_sync(g)) return rc;  // the chain layout changes: nothing may be in flight
    g->io = io;
    return GPS_OK;
}

int gps_sharded_get_host_io(gps_shard_group* g, int* gpu_indices, int* n)
{
    if (!g || !gpu_indices || !n) return fail(GPS_ERR_INVALID, "null argument");
    *n = (int)g->io.size();
    for (int i = 0; i < *n; ++i) gpu_indices[i] = g->io[i];
    return GPS_OK;
}

int gps_sharded_host_sync(gps_shard_group* g)
{
    if (!g) return fail(GPS_ERR_INVALID, "null group");
    int prev = 0;
    cudaGetDevice(&prev);
    int rc = GPS_OK;
    for (int i = 0; i < g->n; ++i) {
        cudaSetDevice(g->devices[i]);
        const int r = gps_host_sync(g->ctx[i]);
        if (r && !rc) rc = r;
    }
    if (!rc) rc = sharded_status(g);
    cudaSetDevice(prev);
    return rc;
}

int gps_sharded_scan_host(gps_shard_group* g, int algo, int kind, int dtype, const void* h_in, void* h_out,
                          uint64_t n, uint32_t carry_in_bits, void* h_total_out)
{
    const int rc = gps_sharded_scan_host_async(g, algo, kind, dtype, h_in, h_out, n, carry_in_bits, h_total_out);
    return rc ? rc : gps_sharded_host_sync(g);
}

}  // extern "C"
