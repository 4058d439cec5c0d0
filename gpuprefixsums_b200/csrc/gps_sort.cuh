// gps_sort.cuh -- LSD radix sort of uint32 keys (optionally with a 32-bit payload): the
// downstream consumer of the chained scan (the reference's README.md:7
// names sorting as what its scans are for -- the sort itself lives in the author's sibling
// repository GPUSorting, which is NOT part of /root/reference, so nothing here follows
// reference source; the algorithm is the published "Onesweep" scheme, Adinets & Merrill 2022).
//
// One pass over the keys builds all digit histograms up front; then ONE kernel per 8-bit
// digit reads each key once and writes it once.  Inside that kernel the 256 per-digit
// offsets of a tile come from a chained scan with decoupled look-back across tiles -- the
// scan of this library, 256 lanes wide: thread d walks the descriptors of digit d.  Re-used
// from the scan kernels: the 64-bit descriptor {value:32 | epoch:30 | status:2} (no clearing
// between passes or sorts: every pass takes a fresh epoch), the atomic tile ticket (forward
// progress without relying on dispatch order), and the tile fetched by one TMA bulk load into
// shared memory.
//   tile pipeline:  TMA -> smem -> registers (warp-striped, KPT keys per thread)
//                   -> warp-level multi-split ranking (match.any, per-warp digit counters)
//                   -> tile digit counts -> publish AGGREGATE -> keys to smem in digit order
//                   -> look-back per digit -> publish INCLUSIVE -> coalesced-by-run scatter
// Stable, so pairs come out in the unique stable order and 4 passes sort 32 bits.
#pragma once
#include "gps_common.cuh"

namespace gps {

constexpr int kRadix = 256;

struct SortParams {
    const void* keys_in;       // uint32_t or uint64_t keys (template parameter K of the kernels)
    void* keys_out;
    const uint32_t* vals_in;   // pairs only
    uint32_t* vals_out;
    uint32_t n;
    uint32_t shift;            // 0, 8, 16, 24 (64-bit keys: ... 56)
    uint64_t* desc;            // num_tiles * 256 descriptors, shared by all passes
    const uint32_t* digit_base;  // 256 exclusive digit offsets of THIS pass
    uint32_t* counter;         // tile-claim counter (never reset)
    uint32_t counter_base;
    uint32_t epoch;
    uint32_t num_tiles;
    uint32_t* status;          // watchdog word
    uint32_t xform_in;         // key type whose order-preserving map is applied on load (first pass), else 0
    uint32_t xform_out;        // key type whose inverse map is applied on store (last pass), else 0
    uint32_t digit_mask;       // 0xFF, or (1 << r) - 1 when only r < 8 bits of this digit take part
                               // (last pass of a sort on end_bit % 8 != 0 bits); honoured by the XFORM
                               // instantiation only, which the host selects for such a pass
};

// Order-preserving maps onto uint32 (keys travel transformed between the passes):
//   1 = int32: flip the sign bit; 2 = float32: flip every bit of negative values, the sign bit
//   of the others (so -0.0 < +0.0 and NaNs sort by their bit pattern at both ends, as CUB does)
//   | KEY_DESC: descending order = the complement of the ascending map (still stable)
enum : uint32_t { KEY_U32 = 0, KEY_I32 = 1, KEY_F32 = 2, KEY_TYPE_MASK = 0xFFu, KEY_DESC = 0x100u };
__host__ __device__ __forceinline__ uint32_t sort_key_map(uint32_t k, uint32_t type)
{
    const uint32_t t = type & KEY_TYPE_MASK;
    if (t == KEY_I32) k ^= 0x80000000u;
    if (t == KEY_F32) k ^= ((k >> 31) ? 0xFFFFFFFFu : 0x80000000u);
    return (type & KEY_DESC) ? ~k : k;
}
__host__ __device__ __forceinline__ uint32_t sort_key_unmap(uint32_t k, uint32_t type)
{
    if (type & KEY_DESC) k = ~k;
    const uint32_t t = type & KEY_TYPE_MASK;
    if (t == KEY_I32) k ^= 0x80000000u;
    if (t == KEY_F32) k ^= ((k >> 31) ? 0x80000000u : 0xFFFFFFFFu);
    return k;
}

// 64-bit keys: the same maps on 64 bits (1 = int64, 2 = float64)
__host__ __device__ __forceinline__ uint64_t sort_key_map(uint64_t k, uint32_t type)
{
    const uint32_t t = type & KEY_TYPE_MASK;
    if (t == KEY_I32) k ^= 0x8000000000000000ull;
    if (t == KEY_F32) k ^= ((k >> 63) ? 0xFFFFFFFFFFFFFFFFull : 0x8000000000000000ull);
    return (type & KEY_DESC) ? ~k : k;
}
__host__ __device__ __forceinline__ uint64_t sort_key_unmap(uint64_t k, uint32_t type)
{
    if (type & KEY_DESC) k = ~k;
    const uint32_t t = type & KEY_TYPE_MASK;
    if (t == KEY_I32) k ^= 0x8000000000000000ull;
    if (t == KEY_F32) k ^= ((k >> 63) ? 0x8000000000000000ull : 0xFFFFFFFFFFFFFFFFull);
    return k;
}

// digit = byte (shift / 8) of the key: one PRMT (selector 0x4440 + byte index: byte 0 <- that
// byte of the key, bytes 1..3 <- byte 0 of the zero operand)
__device__ __forceinline__ uint32_t sort_digit(uint32_t key, uint32_t shift)
{
    return __byte_perm(key, 0u, 0x4440u + (shift >> 3));
}
__device__ __forceinline__ uint32_t sort_digit(uint64_t key, uint32_t shift)  // shift is warp-uniform
{
    return sort_digit(shift >= 32u ? (uint32_t)(key >> 32) : (uint32_t)key, shift & 31u);
}

// 256 descriptors per tile are published by 8 full warps: plain coalesced stores by default
// (one 256-byte line pair per warp), the scan kernels' atom.exch as an option
#ifndef GPS_SORT_PUBLISH_ATOMIC
#define GPS_SORT_PUBLISH_ATOMIC 0
#endif
__device__ __forceinline__ void desc_publish_plain(uint64_t* ptr, uint64_t v)
{
#if GPS_SORT_PUBLISH_ATOMIC
    asm volatile("atom.relaxed.gpu.global.exch.b64 %0, [%1], %2;" : "=l"(v) : "l"(ptr), "l"(v) : "memory");
#else
    st_relaxed_gpu(ptr, v);
#endif
}

// ---- all digit histograms in one sweep: hist[pass][digit] -----------------------------------
// Shared-memory atomics on 4 x 256 counters are bank-conflict bound for random keys (32 random
// addresses per warp instruction: ~3.5 wavefronts each; first version of this kernel: 0.41 ms
// at 2^28 keys = 9.3 lanes/clk/SM).  Here every lane owns a private column of the table
// (counter index = (pass * 256 + digit) * 32 + lane, so bank == lane): conflict free by
// construction.  128 KB of shared memory, one persistent CTA per SM.
constexpr int kHistThreads = 1024;
constexpr size_t kHistSmemBytes = (size_t)4 * kRadix * 32 * sizeof(uint32_t);
constexpr int kMaxSortPasses = 8;  // 64-bit keys

// K = key type; one launch covers up to 4 digit passes [pass0, pass0 + passes) (the table holds 4 x 256
// counters per lane): 32-bit keys need one launch, 64-bit keys two
template <typename K>
__global__ void __launch_bounds__(kHistThreads, 1)
    sort_hist_kernel(const K* __restrict__ keys, uint32_t n, uint32_t* hist, int pass0, int passes, int total_passes,
                     uint32_t key_type, uint32_t last_mask)
{
    extern __shared__ __align__(16) uint32_t s_hist[];  // [4 * 256][32]
    constexpr uint32_t KPV = 16 / sizeof(K);             // keys per 16-byte vector
    const uint32_t tid = threadIdx.x, lane = tid & 31u;
    for (uint32_t i = tid; i < 4u * kRadix * 32u; i += kHistThreads) s_hist[i] = 0u;
    __syncthreads();
    auto add = [&](K k) {
        k = sort_key_map(k, key_type);
#pragma unroll
        for (int q = 0; q < 4; ++q)
            if (q < passes) {  // the last pass of the sort may cover fewer than 8 bits (end_bit % 8 != 0)
                const uint32_t d = (uint32_t)(k >> (8 * (pass0 + q))) & (pass0 + q == total_passes - 1 ? last_mask : 0xFFu);
                atomicAdd(&s_hist[((q * kRadix + d) << 5) + lane], 1u);
            }
    };
    auto add_vec = [&](const uint4& a) {
        if constexpr (sizeof(K) == 4) {
            add((K)a.x); add((K)a.y); add((K)a.z); add((K)a.w);
        } else {
            add((K)(((uint64_t)a.y << 32) | a.x)); add((K)(((uint64_t)a.w << 32) | a.z));
        }
    };
    const uint4* k4 = reinterpret_cast<const uint4*>(keys);
    const uint32_t nvec = n / KPV;
    const uint32_t stride = gridDim.x * kHistThreads;
    uint32_t v = blockIdx.x * kHistThreads + tid;
    // two independent loads in flight per thread
    for (; v + stride < nvec; v += 2u * stride) {
        const uint4 a = ldg_stream(k4 + v);
        const uint4 b = ldg_stream(k4 + v + stride);
        add_vec(a);
        add_vec(b);
    }
    if (v < nvec) add_vec(ldg_stream(k4 + v));
    if (blockIdx.x == 0 && tid < (n % KPV)) add(keys[(uint64_t)nvec * KPV + tid]);
    __syncthreads();
    // thread t folds the 32 columns of counter t (rotated start: conflict free)
    if (tid < (uint32_t)passes * kRadix) {
        uint32_t c = 0;
#pragma unroll
        for (uint32_t i = 0; i < 32u; ++i) c += s_hist[(tid << 5) + ((i + tid) & 31u)];
        if (c) atomicAdd(hist + pass0 * kRadix + tid, c);
    }
}

// ---- exclusive scan of each pass's 256-bin histogram: grid = passes, block = 256 -------------
__global__ void __launch_bounds__(kRadix) sort_digit_base_kernel(const uint32_t* hist, uint32_t* digit_base)
{
    __shared__ uint32_t s_warp[kRadix / 32];
    const uint32_t d = threadIdx.x, lane = d & 31u, warp = d >> 5;
    const uint32_t c = hist[blockIdx.x * kRadix + d];
    const uint32_t inc = warp_inclusive_scan(c, lane);
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    uint32_t wp = 0;
#pragma unroll
    for (int j = 0; j < kRadix / 32; ++j)
        if ((uint32_t)j < warp) wp += s_warp[j];
    digit_base[blockIdx.x * kRadix + d] = wp + inc - c;
}

// ---- one digit pass ------------------------------------------------------------------------
// Lanes of the warp holding the same 8-bit digit.  Measured per digit pass at 2^28 keys on
// B200 (tools/gpu_sort.sh): match.any.sync 2.39 ms; eight ballots written in C++ 1.5 ms (nvcc
// spends ~8 SASS instructions per bit); eight ballots in PTX, 4 instructions per bit: bit test
// into a predicate, ballot, select the lane's own bit as a mask, one LOP3 -- see DESIGN.md.
#ifndef GPS_SORT_MATCH
#define GPS_SORT_MATCH 2  // 0 = ballots in C++, 1 = match.any.sync, 2 = ballots in PTX
#endif

template <int B>
__device__ __forceinline__ void sort_peers_bit(uint32_t& peers, uint32_t d)
{
    // peers &= ~(ballot(bit B of d) ^ (own bit B ? ~0 : 0))   [LOP3 0x90 = a & ~(b ^ c)]
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b32 t, bal, sx;\n\t"
        "and.b32 t, %1, %2;\n\t"
        "setp.ne.u32 p, t, 0;\n\t"
        "vote.sync.ballot.b32 bal, p, 0xffffffff;\n\t"
        "selp.b32 sx, -1, 0, p;\n\t"
        "lop3.b32 %0, %0, bal, sx, 0x90;\n\t}"
        : "+r"(peers)
        : "r"(d), "n"(1 << B));
}

__device__ __forceinline__ uint32_t sort_peers(uint32_t d)
{
#if GPS_SORT_MATCH == 1
    return __match_any_sync(kFullMask, d);
#elif GPS_SORT_MATCH == 2
    uint32_t peers = kFullMask;
    sort_peers_bit<0>(peers, d);
    sort_peers_bit<1>(peers, d);
    sort_peers_bit<2>(peers, d);
    sort_peers_bit<3>(peers, d);
    sort_peers_bit<4>(peers, d);
    sort_peers_bit<5>(peers, d);
    sort_peers_bit<6>(peers, d);
    sort_peers_bit<7>(peers, d);
    return peers;
#else
    uint32_t peers = kFullMask;
#pragma unroll
    for (int b = 0; b < 8; ++b) {
        const uint32_t bal = __ballot_sync(kFullMask, (d >> b) & 1u);
        peers &= ((d >> b) & 1u) ? bal : ~bal;
    }
    return peers;
#endif
}

// Decoupled Fallback for the digit look-back (template flag FB; the plain kernel carries none of it).
// The reference's scans guarantee forward progress without fair scheduling by letting a blocked
// tile recompute its predecessor (EmulatedDeadlocking.cuh:63-151, csdldf_struct.wgsl:161-291); the
// sort's 256-wide look-back gets the same: when some thread's walk has found the same tile not
// ready max_spin rounds in a row, the whole CTA re-counts the digits of the NEAREST such tile from
// the pass's input keys (a pass is out of place, the input is intact) and every thread whose walk
// is blocked exactly there inserts its digit's count with atomicCAS -- or takes what the owner
// published meanwhile.  stall_mask emulates stalled tiles for the tests, stats = {spins,
// fallbacks, insertions, rounds}.
struct SortFallbackParams {
    uint32_t stall_mask;  // tiles with (tile & mask) == 0 stall; 0xFFFFFFFF = none
    uint32_t max_spin;
    uint32_t* stats;      // nullable
};

// XFORM: apply the key maps of p.xform_in / p.xform_out (first / last pass of int32 and float32
// sorts); the uint32 instantiation carries none of that code
template <typename K, int NT, int KPT, bool PAIRS, int REGS = 64, int LB = 8, bool XFORM = false, bool FB = false>
__global__ void __launch_bounds__(NT, min_ctas_per_sm(NT, REGS))
    sort_pass_kernel(const __grid_constant__ SortParams p, const SortFallbackParams f)
{
    constexpr uint32_t KW = sizeof(K) / 4;  // 32-bit words per key
    const K* const keys_in = static_cast<const K*>(p.keys_in);
    K* const keys_out = static_cast<K*>(p.keys_out);
    pdl_entry();
    static_assert(!FB || NT == kRadix, "the fallback rounds are block-wide: every thread owns one digit");
    static_assert(NT >= kRadix && NT % 32 == 0, "one thread per digit for the look-back");
    constexpr int NW = NT / 32;
    constexpr uint32_t TILE = NT * KPT;
    constexpr uint32_t WARP_KEYS = 32 * KPT;
    // dynamic shared memory: keys | payload (pairs) | per-warp digit counters -> placement offsets
    extern __shared__ __align__(128) uint32_t s_dyn[];
    K* const s_keys = reinterpret_cast<K*>(s_dyn);
    uint32_t* const s_vals = s_dyn + KW * TILE;  // payload (always 32 bits): fetched by TMA, permuted in place
    uint32_t(*const s_whist)[kRadix] = reinterpret_cast<uint32_t(*)[kRadix]>(s_dyn + (KW + (PAIRS ? 1 : 0)) * TILE);
    __shared__ uint32_t s_goff[kRadix];       // global position of sorted-tile index 0 of each digit run
    __shared__ uint32_t s_wtot[kRadix / 32];
    __shared__ __align__(8) uint64_t s_bar;
    __shared__ uint32_t s_tile;

    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    // the digit of a (mapped) key in this pass; only the XFORM instantiation pays for the mask
    auto digit_of = [&](K k) -> uint32_t {
        const uint32_t d = sort_digit(k, p.shift);
        return XFORM ? (d & p.digit_mask) : d;
    };

    if (tid == 0) {
        mbar_init(&s_bar, 1);
        fence_mbar_init();
        const uint32_t t = atomicAdd(p.counter, 1u) - p.counter_base;
        s_tile = t;
        if (((uint64_t)t + 1u) * TILE <= p.n) {
            mbar_arrive_expect_tx(&s_bar, TILE * 4u * (KW + (PAIRS ? 1u : 0u)));
            bulk_load(s_keys, keys_in + (uint64_t)t * TILE, TILE * 4u * KW, &s_bar);
            if constexpr (PAIRS) bulk_load(s_vals, p.vals_in + (uint64_t)t * TILE, TILE * 4u, &s_bar);
        }
    }
#pragma unroll
    for (uint32_t i = tid; i < (uint32_t)NW * kRadix; i += NT) (&s_whist[0][0])[i] = 0u;
    __syncthreads();
    const uint32_t tile = s_tile;
    if (tile >= p.num_tiles) {  // ticket accounting out of step (an earlier launch failed)
        if (tid == 0) atomicOr(p.status, STATUS_WATCHDOG);
        return;
    }
    const uint64_t tile_base = (uint64_t)tile * TILE;
    const bool full = tile_base + TILE <= p.n;
    const uint32_t valid = full ? TILE : (uint32_t)(p.n - tile_base);

    if (full) {
        if (!mbar_wait_bounded(&s_bar, 0u)) {
            atomicOr(p.status, STATUS_WATCHDOG);
            return;
        }
    } else {
        // last tile: pad with (the preimage of) 0xFFFFFFFF -- the largest digit of every pass (255, or
        // digit_mask), ranked after every real key of the tile, so the padding occupies exactly the
        // sorted-tile indices >= valid
        const K pad = XFORM ? sort_key_unmap((K)~(K)0, p.xform_in) : (K)~(K)0;
#pragma unroll 1
        for (uint32_t i = tid; i < TILE; i += NT) {
            s_keys[i] = i < valid ? keys_in[tile_base + i] : pad;
            if constexpr (PAIRS) s_vals[i] = i < valid ? p.vals_in[tile_base + i] : 0u;
        }
        __syncthreads();
    }

    // ---- keys to registers, warp-striped: key j of lane l is element warp*32*KPT + j*32 + l
    K key[KPT];
#pragma unroll
    for (int j = 0; j < KPT; ++j) {
        key[j] = s_keys[warp * WARP_KEYS + j * 32 + lane];
        if constexpr (XFORM) key[j] = sort_key_map(key[j], p.xform_in);
    }
    // ---- warp-level multi-split: rank of each key among the warp's keys of the same digit,
    // in (j, lane) order == original order (stability)
    // ranks / positions are < TILE <= 65536: two per register (registers decide the occupancy)
    static_assert(TILE <= 65536 && KPT % 2 == 0, "16-bit packed positions");
    uint32_t pos2[KPT / 2] = {};
    auto get_pos = [&](int j) -> uint32_t { return (j & 1) ? (pos2[j >> 1] >> 16) : (pos2[j >> 1] & 0xFFFFu); };
    auto set_pos = [&](int j, uint32_t v) {
        pos2[j >> 1] = (j & 1) ? __byte_perm(pos2[j >> 1], v, 0x5410)   // high half <- v, low half kept
                               : __byte_perm(v, pos2[j >> 1], 0x7610);  // low half <- v, high half kept
    };
#pragma unroll
    for (int j = 0; j < KPT; ++j) {
        const uint32_t d = digit_of(key[j]);
        const uint32_t peers = sort_peers(d);
        const uint32_t below = __popc(peers & ((1u << lane) - 1u));
        const uint32_t pre = s_whist[warp][d];
        __syncwarp();
        if (below == 0u) s_whist[warp][d] = pre + __popc(peers);
        __syncwarp();
        set_pos(j, pre + below);
    }
    __syncthreads();  // every warp counted; s_keys no longer needed in its loaded order

    // ---- thread d: digit d's count in the tile, warp-exclusive offsets, block scan over digits
    uint32_t count = 0, cta_excl = 0;
    if (tid < (uint32_t)kRadix) {
#pragma unroll
        for (int w = 0; w < NW; ++w) {
            const uint32_t c = s_whist[w][tid];
            s_whist[w][tid] = count;
            count += c;
        }
        if constexpr (FB) {
            // stall emulation: refuse to publish until a successor's fallback has inserted this digit's
            // word (tile 0 is rescued with INCLUSIVE); the last tile never stalls
            const bool stalled = f.stall_mask != 0xFFFFFFFFu && (tile & f.stall_mask) == 0u && p.num_tiles > 1u &&
                                 tile != p.num_tiles - 1u;
            if (stalled) {
                uint32_t polls = 0;
                while ((uint32_t)(ld_relaxed_gpu(p.desc + (uint64_t)tile * kRadix + tid) >> 34) != p.epoch)
                    if (watchdog_tick(polls, p.status)) break;
            }
        }
        if (tile != 0) desc_publish_plain(p.desc + (uint64_t)tile * kRadix + tid, pack_desc(count, p.epoch, ST_AGGREGATE));
        const uint32_t inc = warp_inclusive_scan(count, lane);
        if (lane == 31) s_wtot[warp] = inc;
        cta_excl = inc - count;
    }
    __syncthreads();
    if (tid < (uint32_t)kRadix) {
#pragma unroll
        for (int j = 0; j < kRadix / 32; ++j)
            if ((uint32_t)j < warp) cta_excl += s_wtot[j];
#pragma unroll
        for (int w = 0; w < NW; ++w) s_whist[w][tid] += cta_excl;
    }
    __syncthreads();

    // ---- keys into shared memory in digit order
#pragma unroll
    for (int j = 0; j < KPT; ++j) {
        const uint32_t q = get_pos(j) + s_whist[warp][digit_of(key[j])];
        set_pos(j, q);
        s_keys[q] = key[j];
    }
    if constexpr (PAIRS) {
        // the payload takes the same permutation, in place in its own buffer (no registers are
        // held across the look-back; the permutation does not depend on it)
        uint32_t val[KPT];
#pragma unroll
        for (int j = 0; j < KPT; ++j) val[j] = s_vals[warp * WARP_KEYS + j * 32 + lane];
        __syncthreads();
#pragma unroll
        for (int j = 0; j < KPT; ++j) s_vals[get_pos(j)] = val[j];
    }

    // ---- chained scan across tiles, one look-back per digit (thread d walks digit d)
    if constexpr (FB) {
        __shared__ uint32_t s_fbt[3];  // per round (mod 3): nearest tile some walk is stuck at, + 1
        uint32_t* const s_fbhist = &s_whist[0][0];  // the per-warp counters are dead by now (after the barrier below)
        uint32_t prefix = 0, n_spins = 0, n_fallbacks = 0, n_inserts = 0, n_rounds = 0;
        if (tid < 3u) s_fbt[tid] = 0u;
        __syncthreads();
        if (tile != 0) {  // block-uniform
            uint32_t polls = 0, spins = 0;
            int64_t t = (int64_t)tile - 1;
            bool done = false;
            const uint32_t tag = p.epoch << 2;
            for (uint32_t round = 0;; ++round) {
                if (tid == 0) s_fbt[(round + 1u) % 3u] = 0u;  // last read two rounds ago
                bool want_fb = false;
                if (!done) {
                    const uint64_t* dp = p.desc + (uint64_t)t * kRadix + tid;
                    uint32_t lo[LB], hi[LB];
#pragma unroll
                    for (int i = 0; i < LB; ++i) {
                        lo[i] = 0u;
                        hi[i] = tag | ST_INCLUSIVE;  // virtual entry below tile 0
                        if (t - i >= 0) {
                            const uint64_t d = ld_relaxed_gpu(dp - (int64_t)i * kRadix);
                            lo[i] = (uint32_t)d;
                            hi[i] = (uint32_t)(d >> 32);
                        }
                    }
                    uint32_t consumed = 0;
                    bool open = true;
#pragma unroll
                    for (int i = 0; i < LB; ++i) {
                        const uint32_t st = hi[i] ^ tag;
                        const bool ok = open && st < 4u && st != 0u;
                        if (ok) {
                            prefix += lo[i];
                            ++consumed;
                            done = st == ST_INCLUSIVE;
                        }
                        open = ok && !done;
                    }
                    t -= consumed;
                    ++n_rounds;
                    if (consumed == 0u) {
                        ++n_spins;
                        if (watchdog_tick(polls, p.status)) done = true;  // give up (reported by gps_check)
                        want_fb = !done && ++spins >= f.max_spin;
                    } else {
                        spins = 0;
                    }
                }
                if (want_fb) atomicMax(&s_fbt[round % 3u], (uint32_t)t + 1u);
                if (__syncthreads_and(done)) break;
                const uint32_t fbt = s_fbt[round % 3u];
                if (fbt == 0u) continue;  // block-uniform: nobody is stuck long enough
                // ---- Decoupled Fallback on tile fb (never the last tile => full): re-count its digits
                const uint32_t fb = fbt - 1u;
                s_fbhist[tid] = 0u;
                __syncthreads();
#pragma unroll 4
                for (int k = 0; k < KPT; ++k) {
                    K kk = keys_in[(uint64_t)fb * TILE + (uint32_t)k * NT + tid];
                    if constexpr (XFORM) kk = sort_key_map(kk, p.xform_in);
                    atomicAdd(&s_fbhist[digit_of(kk)], 1u);
                }
                __syncthreads();
                if (tid == 0) ++n_fallbacks;
                if (!done && t == (int64_t)fb) {  // my walk is stuck exactly there
                    const uint32_t cnt = s_fbhist[tid];
                    const uint32_t st_mine = fb ? ST_AGGREGATE : ST_INCLUSIVE;
                    const uint64_t mine = pack_desc(cnt, p.epoch, st_mine);
                    uint64_t* ptr = p.desc + (uint64_t)fb * kRadix + tid;
                    uint64_t seen = ld_relaxed_gpu(ptr);
                    uint32_t use = cnt, use_st = st_mine;
                    bool inserted = false;
                    while ((uint32_t)(seen >> 34) != p.epoch) {
                        const uint64_t old = atomicCAS(reinterpret_cast<unsigned long long*>(ptr), (unsigned long long)seen,
                                                       (unsigned long long)mine);
                        if (old == seen) {
                            inserted = true;
                            break;
                        }
                        seen = old;
                    }
                    if (inserted) {
                        ++n_inserts;
                    } else {  // the owner (or another rescuer) got there first
                        use = (uint32_t)seen;
                        use_st = (uint32_t)(seen >> 32) & ST_MASK;
                    }
                    prefix += use;
                    done = use_st == ST_INCLUSIVE;
                    t -= 1;
                    spins = 0;
                }
                __syncthreads();  // s_fbhist is zeroed again by the next fallback
            }
        }
        desc_publish_plain(p.desc + (uint64_t)tile * kRadix + tid, pack_desc(prefix + count, p.epoch, ST_INCLUSIVE));
        s_goff[tid] = p.digit_base[tid] + prefix - cta_excl;
        if (f.stats) {
            if (n_spins) atomicAdd(f.stats + 0, n_spins);
            if (n_fallbacks) atomicAdd(f.stats + 1, n_fallbacks);
            if (n_inserts) atomicAdd(f.stats + 2, n_inserts);
            if (tid == 0 && n_rounds) atomicAdd(f.stats + 3, n_rounds);
        }
    } else
    if (tid < (uint32_t)kRadix) {
        // Thread d walks digit d's descriptors backwards, LB descriptors per L2 round trip: the
        // loads of one round are independent, the walk consumes them nearest first and stops
        // at the first INCLUSIVE (or re-polls from the first one that is not ready).  With
        // hundreds of tiles in flight and a tile finishing every ~20 ns, the nearest INCLUSIVE
        // is usually several tiles back.  Measured: LB = 4 is the best (50.0 Gkeys/s; LB = 1
        // 48.7, LB = 8 45.7, LB = 16 37.8 -- deeper rounds mostly fetch descriptors that are
        // never consumed).
        uint32_t prefix = 0;
        if (tile != 0) {
            uint32_t polls = 0;
            int64_t t = (int64_t)tile - 1;
            bool done = false;
            const uint32_t tag = p.epoch << 2;  // hi word of a valid descriptor = tag | status
            const uint64_t* dp = p.desc + (uint64_t)t * kRadix + tid;
            while (!done) {
                uint32_t lo[LB], hi[LB];
#pragma unroll
                for (int i = 0; i < LB; ++i) {
                    lo[i] = 0u;
                    hi[i] = tag | ST_INCLUSIVE;  // virtual entry below tile 0
                    if (t - i >= 0) {
                        const uint64_t d = ld_relaxed_gpu(dp - (int64_t)i * kRadix);
                        lo[i] = (uint32_t)d;
                        hi[i] = (uint32_t)(d >> 32);
                    }
                }
                uint32_t consumed = 0;
                bool open = true;  // still consuming this round
#pragma unroll
                for (int i = 0; i < LB; ++i) {
                    const uint32_t st = hi[i] ^ tag;  // 1 / 2 when valid, >= 4 otherwise
                    const bool ok = open && st < 4u && st != 0u;
                    if (ok) {
                        prefix += lo[i];
                        ++consumed;
                        done = st == ST_INCLUSIVE;
                    }
                    open = ok && !done;
                }
                t -= consumed;
                dp -= (int64_t)consumed * kRadix;
                if (consumed == 0u && watchdog_tick(polls, p.status)) break;
            }
        }
        desc_publish_plain(p.desc + (uint64_t)tile * kRadix + tid, pack_desc(prefix + count, p.epoch, ST_INCLUSIVE));
        s_goff[tid] = p.digit_base[tid] + prefix - cta_excl;
    }
    __syncthreads();

    // ---- scatter: consecutive sorted-tile indices of one digit are consecutive in memory
#pragma unroll
    for (int k = 0; k < KPT; ++k) {
        const uint32_t i = k * NT + tid;
        const K kk = s_keys[i];
        const uint32_t g = s_goff[digit_of(kk)] + i;
        if (full || i < valid) {
            keys_out[g] = XFORM ? sort_key_unmap(kk, p.xform_out) : kk;
            if constexpr (PAIRS) p.vals_out[g] = s_vals[i];
        }
    }
}

}  // namespace gps
