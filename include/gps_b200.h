/*
 * gps_b200.h -- C ABI of the B200-native device-wide prefix-sum library
 * (libgps_b200.so).  Plain pointers and sizes only; no C++/torch types.
 *
 * The reference (b0nes164/GPUPrefixSums, CUDA port) has no FFI layer: its operator
 * surface is four header-defined C++ dispatcher classes that launch kernels with
 * <<< >>> directly.  Every entry point below replaces one of those private launch /
 * harness methods; the C++ classes with the reference's names live in
 * gpuprefixsums_b200/host/ (the .hpp files) and call only this ABI.  File:line citations are
 * relative to /root/reference/GPUPrefixSumsCUDA/.
 *
 * Conventions
 *   - all data pointers named d_* are DEVICE pointers owned by the caller and must be
 *     16-byte aligned; h_* are host pointers.  n counts keys (4 bytes each), 64-bit.
 *   - d_out may equal d_in (the reference scans in place) but must not partially
 *     overlap it.
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).
 *     Calls are asynchronous on that stream unless stated otherwise.
 *   - a gps_ctx owns the scratch (tile descriptors, tile counter, partials, status
 *     word) for one device; use one ctx per host thread / stream.  Scratch is
 *     epoch-tagged: no per-launch memset (the reference clears its descriptors inside
 *     the timed region, ChainedScanDecoupledLookbackDispatcher.cuh:172-175).
 *   - every function returns GPS_OK (0) or a negative gps_status; gps_last_error()
 *     gives a thread-local message.  There is NO CPU fallback: without a usable CUDA
 *     device every compute entry point fails with GPS_ERR_CUDA.
 */
#ifndef GPS_B200_H_
#define GPS_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GPS_ABI_VERSION 1

typedef struct gps_ctx gps_ctx;
typedef struct gps_shard_group gps_shard_group;

enum gps_status {
    GPS_OK = 0,
    GPS_ERR_INVALID = -1,   /* bad enum / null pointer / n > ctx capacity            */
    GPS_ERR_ALIGNMENT = -2, /* device pointer not 16-byte aligned                    */
    GPS_ERR_CUDA = -3,      /* a CUDA runtime call failed (message in last_error)    */
    GPS_ERR_WATCHDOG = -4,  /* a look-back spin exceeded its poll budget (reported by */
                            /* gps_check; the kernel bails out instead of hanging)   */
    GPS_ERR_MISMATCH = -5,  /* validation helper found wrong results                 */
    GPS_ERR_UNSUPPORTED = -6 /* not an sm_100 device; a captured sort larger than any sort run before */
};

enum gps_algo {
    GPS_ALGO_CSDL = 0,          /* Chained Scan w/ Decoupled Lookback, tuned default  */
    GPS_ALGO_CSDL_REG = 1,      /*   register path: 128-bit coalesced LDG/STG          */
    GPS_ALGO_CSDL_TMA = 2,      /*   TMA bulk copies into a shared-memory ring        */
    GPS_ALGO_CSDL_FALLBACK = 3, /* CSDL + Decoupled Fallback (no emulated stalls)     */
    GPS_ALGO_RTS = 4,           /* three-kernel Reduce-then-Scan                      */
    GPS_ALGO_CSDL_SMEM = 5,     /*   one TMA bulk load per CTA, tile resident in smem  */
    GPS_ALGO_CSDL_BLOCKED = 6   /*   same, blocked tile: one warp scan per tile, TMA store */
};
enum gps_kind { GPS_INCLUSIVE = 0, GPS_EXCLUSIVE = 1 };
enum gps_dtype {
    GPS_U32 = 0,
    GPS_F32 = 1,
    /* wide element types: gps_scan_wide only */
    GPS_U64 = 2,
    GPS_F64 = 3,
    GPS_U32X4 = 4 /* struct scan: four interleaved uint32 lanes per 16-byte element */
};

/* ---- library / device ------------------------------------------------------- */
int gps_abi_version(void);
const char* gps_last_error(void);
int gps_device_count(int* count);
/* name (<=255 chars), SM count, L2 bytes, total memory of `device` */
int gps_device_info(int device, char* name, int* sm_count, uint64_t* l2_bytes,
                    uint64_t* mem_bytes);

/* ---- context: replaces the dispatcher constructors/destructors --------------
 * ChainedScanDecoupledLookbackDispatcher.cuh:28-34,158-163 (m_index,
 * m_threadBlockReduction), ReduceThenScanDispatcher.cuh:27-33,152-156.  Unlike the
 * reference the ctx does not own the data buffer. */
/* Threading / streams: a ctx is thread-compatible, not thread-safe (one host thread at a time).
 * Its scratch is consumed in launch order; launches may go to any stream -- when consecutive
 * launches of one ctx use different streams the library orders them with an event (so they
 * serialise).  For concurrent scans use one ctx per stream.
 * CUDA graphs: gps_scan, gps_scan_wide, gps_scan_emulated_deadlock, gps_sort_* and gps_reduce*
 * may be captured and replayed.  Reduce-then-Scan keeps no launch state; for the chained kernels
 * and the sort a captured launch runs on a shadow scratch of the ctx: the first captured launch
 * of a capture records a clear of it into the graph, the following ones take consecutive epochs
 * and ticket bases, so every replay starts from the same state and plain launches of the ctx are
 * unaffected.  Make one call of each kind outside the capture
 * first (function attributes and the sort's lazily sized scratch cannot be set up while
 * capturing); captured launches of ONE ctx must be ordered with respect to each other. */
int gps_create(gps_ctx** ctx, int device, uint64_t max_n);
void gps_destroy(gps_ctx* ctx);
/* Synchronises `stream` and reports GPS_ERR_WATCHDOG if any launch since the last
 * check tripped the look-back watchdog (then clears the flag). */
int gps_check(gps_ctx* ctx, void* stream);
/* Tuning knobs (key/value); unknown keys return GPS_ERR_INVALID.  Keys:
 *   "csdl_variant"  0 = auto, 1 = register path, 2 = TMA ring, 3 = smem-resident tile
 *                   (lane-strided), 4 = smem-resident tile (blocked, TMA store)
 *   "fallback_variant" Decoupled Fallback kernels: 0 = auto (blocked smem-resident tile from
 *                   2^27 keys; below, the lane-strided smem tile above 2^21 keys when no stalls
 *                   are emulated, else the register tile), 1 = register tile, 2 = smem-resident
 *                   tile, 3 = blocked smem-resident tile
 *   "ticket_free"   1 (default) = GPS_ALGO_CSDL takes tile = blockIdx.x instead of an atomic ticket
 *                   where that does not rest on in-order CTA dispatch: (a) grids of at most one wave
 *                   of a register-tile kernel (<256,4> up to ~3.6 M keys on 148 SMs, <128,8> from
 *                   there to ~5.4 M) -- every CTA becomes resident without another CTA of the grid
 *                   finishing first, and a walk whose nearest predecessor stays unpublished for
 *                   ~10 ms re-reduces that tile itself and inserts it with CAS (warp-level Decoupled
 *                   Fallback on the cold path; gps_get_option "rescued"), so nothing is assumed of
 *                   the dispatch order or of the kernels sharing the device (in place the rescue is
 *                   off: those launches need other kernels on the device to terminate independently
 *                   of the scan, or in-order dispatch -- what cub::DeviceScan relies on); (b) the
 *                   Decoupled Fallback kernel -- no CTA of it waits for another -- for the sizes up
 *                   to 2^24 keys that (a) does not cover.  6..24 % faster there.
 *                   2 = the same dispatch with the lean one-wave kernels (no rescue code: another
 *                   0.07..0.25 us per launch, (a) then rests on the assumption stated for in place).
 *                   0 = atomic tickets everywhere
 *   "rts_variant"   Reduce-then-Scan propagate kernel: 0 = auto (blocked tile from 2^27 keys),
 *                   1 = register tile, 2 = blocked smem-resident tile
 *   "tile_order"    0 = atomic ticket (default), 1 = blockIdx.x order (as cub::DeviceScan)
 *   "debug_delay_tile" test hook: the CTA of tile (value - 1) of the register-tile kernels starts
 *                   ~40 ms late (0 = off) -- what out-of-order dispatch would look like
 *   "desc_protocol" 0 = atom.exch + ld.relaxed (default), 1 = st.release + ld.acquire,
 *                   2 = st.relaxed + ld.relaxed
 *   "lookback"      0 = warp-parallel (default, LookbackWarp), 1 = single thread
 *                   (LookbackSingle, ChainedScanDecoupledLookback.cuh:28-47)
 *   "pdl"           1 = launch the scan kernels with programmatic stream serialization (default:
 *                   their CTAs become resident while the stream's previous kernel drains and
 *                   wait for it to complete before touching memory), 0 = plain launches
 *   "poll_sleep_ns" nanosleep between unsuccessful descriptor polls (default 0)
 *   "l2_hints"      L2 cache policies on the tile loads / stores of the smem-resident kernels:
 *                   -1 = auto (default: evict_first loads when input + output are about the size of
 *                   the L2, measured +6 % at 2^24 keys), 0 = off, bit 0 = loads evict_first, bit 1 =
 *                   stores evict_last
 *   "slot_wait_polls" polls a scan may spend waiting for ONE carry slot (gps_scan_carry_slots,
 *                   the host pipelines) before it gives up: default 0 = 2^26 polls, over a minute
 *                   -- the writer may be a peer process that is legitimately late.  A slot that
 *                   never arrives contributes nothing and gps_check / gps_sharded_* report
 *                   GPS_ERR_WATCHDOG
 *   "host_chunk_keys" chunk size of gps_scan_host (default 2^24; before first use) */
int gps_set_option(gps_ctx* ctx, const char* key, int64_t value);
/* Reads an option back: "ticket_free", "pdl", "tile_order", and "rescued" -- 1 if a gps_check since
 * the last query saw that a ticket-free one-wave scan rescued a predecessor tile whose CTA was late
 * (warp-level Decoupled Fallback; the results are valid, the flag is informational and reading clears it). */
int gps_get_option(gps_ctx* ctx, const char* key, int64_t* value);
/* Scratch a context of capacity max_n needs on the device, in bytes (descriptors + RTS
 * partials + counters) -- the "temp storage size" of the Unity port's caller-owned-buffer
 * API (GPUPrefixSumsUnity/Runtime/ReduceThenScan.cs:156-279,
 * ChainedScanDecoupledLookbackDecoupledFallback.cs:120-186). */
uint64_t gps_scratch_bytes(uint64_t max_n);
/* gps_create on CALLER-OWNED temp storage: d_scratch is a device buffer of at least
 * gps_scratch_bytes(max_n) bytes, 256-byte aligned, on `device`.  It is zeroed here once and
 * never freed by the library; it must outlive the ctx and must not be touched while a scan
 * of this ctx is in flight.  d_scratch == NULL behaves like gps_create. */
int gps_create_with_scratch(gps_ctx** ctx, int device, uint64_t max_n, void* d_scratch,
                            uint64_t scratch_bytes);

/* ---- scans -------------------------------------------------------------------
 * gps_scan replaces DispatchKernelsInclusive / DispatchKernelsExclusive of
 *   ChainedScanDecoupledLookbackDispatcher.cuh:177-195 (algo CSDL*) and
 *   ReduceThenScanDispatcher.cuh:165-187               (algo RTS).
 * carry_in_bits: value (uint32, or float32 bit pattern) added to every output -- the
 * prefix seeded into tile 0; 0 reproduces the reference.  d_total_out (nullable)
 * receives carry + sum of all n keys (one key-sized value).  n == 0 is a no-op that
 * still writes d_total_out. */
int gps_scan(gps_ctx* ctx, int algo, int kind, int dtype, const void* d_in, void* d_out,
             uint64_t n, uint32_t carry_in_bits, void* d_total_out, void* stream);

/* Scans over wide elements: GPS_U64, GPS_F64 (8-byte keys) and GPS_U32X4,
 * the four-lane struct scan of GPUPrefixSumsWebGPUapis/SharedShaders/csdldf_struct.wgsl:61-316
 * (validate.wgsl:50-60).  n counts ELEMENTS (8 or 16 bytes each); n * element size must not
 * exceed 4 * max_n bytes of the ctx.  h_carry (nullable) points to one element on the HOST
 * that seeds the scan; d_total_out (nullable) receives carry + sum on the device.
 * algo: GPS_ALGO_CSDL = chained scan with decoupled look-back over two-set epoch-tagged
 * descriptors (single pass, 2 x element size of HBM traffic per element), GPS_ALGO_CSDL_FALLBACK =
 * the same with Decoupled Fallback (what csdldf_struct.wgsl is: spin budget, whole-CTA
 * re-reduction of the blocking tile from the input, per-word atomicCAS insertion), GPS_ALGO_RTS =
 * Reduce-then-Scan (3 x, deterministic for float64).  Both use the blocked shared-memory
 * tile of csrc/gps_wide.cuh (TMA bulk load and store).  GPS_U32 / GPS_F32 are accepted too:
 * the same kernel at 32 bits, kept to compare it with gps_scan's kernels.  Asynchronous. */
int gps_scan_wide(gps_ctx* ctx, int algo, int kind, int dtype, const void* d_in, void* d_out,
                  uint64_t n, const void* h_carry, void* d_total_out, void* stream);

/* The wide scan's emulated-deadlock harness (the struct scan's counterpart of
 * gps_scan_emulated_deadlock; GPUPrefixSumsWebGPUapis/SharedShaders/csdldf_struct.wgsl:161-291 with
 * the stall injection of TestVariants/): tiles with (tile & stall_mask) == 0 refuse to publish until
 * a successor's fallback has inserted their aggregate; the last tile never stalls.  d_stats
 * (nullable) accumulates {spins, fallbacks, insertions, look-back windows}.  Bit-exact for every
 * element type, float64 included: a rescuer reproduces the owner's association order. */
int gps_scan_wide_emulated_deadlock(gps_ctx* ctx, int kind, int dtype, const void* d_in, void* d_out,
                                    uint64_t n, const void* h_carry, uint32_t stall_mask,
                                    uint32_t max_spin, uint32_t* d_stats, void* stream);

/* ---- radix sort: the scan's consumer (reference README.md:7 names sorting
 * as the downstream use of its scans -- the sort itself is in the sibling repository
 * GPUSorting, outside /root/reference).  Stable LSD radix sort of uint32 keys, ascending, on
 * the low `end_bit` bits (32 = full sort): the result is the stable order of key & (2^end_bit-1),
 * the bits above end_bit do not take part (cub::DeviceRadixSort's end_bit with begin_bit 0).
 * 8 bits per pass; the last pass covers end_bit % 8 bits when that is not 0.  One sweep builds the digit
 * histograms; each digit pass reads and writes every key once and obtains its 256 per-tile
 * digit offsets from a chained scan with decoupled look-back (csrc/gps_sort.cuh).
 * d_vals_* (all three or none): a 32-bit payload moved with its key.  The result lands in
 * d_*_out; d_*_tmp is a work buffer of n elements distinct from in and out; in may equal out
 * when the number of passes ceil(end_bit/8) is even; the input is otherwise preserved.
 * n < 2^32.  The ctx keeps the sort's own scratch (gps_sort_scratch_bytes(n) bytes, grown on
 * demand).  Asynchronous on `stream`; gps_check reports a tripped watchdog. */
uint64_t gps_sort_scratch_bytes(uint64_t n);
/* key types of gps_sort_keys32: int32 and float32 keys are mapped onto uint32 order on the
 * way in (first pass) and back on the way out (last pass): ascending numeric order, -0.0
 * before +0.0, NaNs at both ends by bit pattern (the convention of cub::DeviceRadixSort);
 * both sort on all 32 bits (end_bit must be 32). */
enum gps_key_type {
    GPS_KEY_U32 = 0,
    GPS_KEY_I32 = 1,
    GPS_KEY_F32 = 2,
    GPS_KEY_DESCENDING = 0x100 /* OR-ed to a type: descending order (stable; end_bit must be 32) */
};
int gps_sort_keys32(gps_ctx* ctx, int key_type, const void* d_keys_in, void* d_keys_out, void* d_keys_tmp,
                    const void* d_vals_in, void* d_vals_out, void* d_vals_tmp, uint64_t n,
                    int end_bit, void* stream);
/* The same sort with Decoupled Fallback in the per-digit look-back (forward progress without fair
 * scheduling, as the reference's scans have it: EmulatedDeadlocking.cuh:63-151): a tile whose walk
 * finds the same predecessor not ready max_spin rounds in a row re-counts that predecessor's digits
 * from the pass's input and inserts them with atomicCAS.  stall_mask emulates stalled tiles
 * ((tile & mask) == 0, never the last; 0xFFFFFFFF = none: the production variant); d_stats
 * (nullable) accumulates {spins, fallbacks, insertions, look-back rounds} over all passes. */
int gps_sort_keys32_fallback(gps_ctx* ctx, int key_type, const void* d_keys_in, void* d_keys_out,
                             void* d_keys_tmp, const void* d_vals_in, void* d_vals_out,
                             void* d_vals_tmp, uint64_t n, int end_bit, uint32_t stall_mask,
                             uint32_t max_spin, uint32_t* d_stats, void* stream);
/* 64-bit keys: uint64 (GPS_KEY_U32 = "unsigned"), int64 (GPS_KEY_I32 = "signed") and float64
 * (GPS_KEY_F32 = "floating point") by the same order-preserving maps, | GPS_KEY_DESCENDING; 8 digit
 * passes, 32-bit payload (d_vals_*, all three or none), end_bit 1..64 for unsigned keys.  Same
 * kernels, instantiated for 64-bit keys with half as many keys per thread; same buffer rules. */
int gps_sort_keys64(gps_ctx* ctx, int key_type, const void* d_keys_in, void* d_keys_out, void* d_keys_tmp,
                    const void* d_vals_in, void* d_vals_out, void* d_vals_tmp, uint64_t n,
                    int end_bit, void* stream);
int gps_sort_keys64_fallback(gps_ctx* ctx, int key_type, const void* d_keys_in, void* d_keys_out,
                             void* d_keys_tmp, const void* d_vals_in, void* d_vals_out,
                             void* d_vals_tmp, uint64_t n, int end_bit, uint32_t stall_mask,
                             uint32_t max_spin, uint32_t* d_stats, void* stream);
/* = gps_sort_keys32(GPS_KEY_U32, ...) */
int gps_sort_u32(gps_ctx* ctx, const void* d_keys_in, void* d_keys_out, void* d_keys_tmp,
                 const void* d_vals_in, void* d_vals_out, void* d_vals_tmp, uint64_t n,
                 int end_bit, void* stream);

/* Same, with the carry read ON THE DEVICE by tile 0 from 64-bit carry slots
 * {value bits : 32 | tag : 32} (low word = value).  Tile 0 polls all `n_slots` slots
 * with ld.relaxed.sys (strong, system scope; the word is self-contained) until their tag equals `tag`, sums the values of the first
 * `n_carry` (<= n_slots) of them and seeds the look-back chain -- the multi-GPU carry
 * "folded into the look-back" (slots may be written by peer GPUs over NVLink, see
 * gps_reduce_publish).  Polling more slots than are summed (all G ranks, not only the
 * upstream ones) is the flow control that lets the slots be reused every other call.
 * New capability; the reference is single-GPU. */
int gps_scan_carry_slots(gps_ctx* ctx, int algo, int kind, int dtype, const void* d_in,
                         void* d_out, uint64_t n, const uint64_t* d_slots, uint32_t n_slots,
                         uint32_t n_carry, uint32_t tag, void* d_total_out, void* stream);

/* Sum of n keys -> *d_total_out (key-sized).  The Reduce pass of
 * ReduceThenScan.cuh:15-53 followed by a one-block fold of the partials. */
int gps_reduce(gps_ctx* ctx, int dtype, const void* d_in, uint64_t n, void* d_total_out,
               void* stream);

/* Reduce + publish: writes {total bits | tag<<32} with st.relaxed.sys (system scope) to each of the
 * `n_dst` device addresses in h_dst_slots (host array of device pointers; may point
 * into peer-GPU memory mapped with gps_ipc_open / peer access). */
int gps_reduce_publish(gps_ctx* ctx, int dtype, const void* d_in, uint64_t n,
                       uint64_t* const* h_dst_slots, uint32_t n_dst, uint32_t tag,
                       void* d_total_out, void* stream);

/* Emulated-deadlock harness kernel: replaces
 * EmulatedDeadlockingDispatcher.cuh:152-158 -> EmulatedDeadlocking.cuh:26-168.
 * Tiles with (tile & stall_mask) == 0 stall until a successor's Decoupled Fallback
 * has published their aggregate (atomic CAS); the LAST tile never stalls (the
 * reference hangs there).  stall_mask = 0xFFFFFFFF disables stalling
 * (pure Decoupled Fallback).  max_spin = unsuccessful polls before a fallback (the
 * reference's MAX_SPIN_COUNT 4, EmulatedDeadlocking.cuh:15).  d_stats (nullable)
 * accumulates {total spins, fallbacks initiated, successful insertions, lookbacks}
 * (GPUPrefixSumsWebGPUapis/SharedShaders/TestVariants/csdldf_stats.wgsl). */
int gps_scan_emulated_deadlock(gps_ctx* ctx, int kind, int dtype, const void* d_in,
                               void* d_out, uint64_t n, uint32_t stall_mask,
                               uint32_t max_spin, uint32_t* d_stats, void* stream);

/* Occupancy-estimation variant of the Decoupled Fallback scan (GPUPrefixSumsWebGPUapis/SharedShaders/
 * TestVariants/csdldf_occ.wgsl:72-82, read back by Dawn/main.cpp:663-667,833-836): one CTA per tile is
 * launched, every CTA keeps claiming tiles until none are left and counts itself once if its first
 * claim was a real tile; *h_occupancy = that count = an estimate of how many CTAs of the kernel the
 * device holds at once (for inputs with more tiles than that).  d_out receives the correct scan.
 * Synchronous. */
int gps_scan_estimate_occupancy(gps_ctx* ctx, int kind, int dtype, const void* d_in, void* d_out,
                                uint64_t n, uint32_t* h_occupancy, void* stream);

/* The reference-facing call with HOST buffers: pipelined H2D -> scan -> D2H in
 * chunks on internal streams, chunk totals chained on the device.  Synchronous.
 * h_in / h_out should be pinned for full PCIe bandwidth (pageable works). */
int gps_scan_host(gps_ctx* ctx, int algo, int kind, int dtype, const void* h_in,
                  void* h_out, uint64_t n, uint32_t carry_in_bits, void* h_total_out);
/* The same without the final wait: returns once everything is queued.  Consecutive calls share
 * the staging-buffer rotation, so the first copy-in of a call overlaps the last copy-out of
 * the previous one.  gps_host_sync waits for all queued host-buffer work of the ctx. */
int gps_scan_host_async(gps_ctx* ctx, int algo, int kind, int dtype, const void* h_in,
                  void* h_out, uint64_t n, uint32_t carry_in_bits, void* h_total_out);
int gps_host_sync(gps_ctx* ctx);

/* ---- harness kernels: UtilityKernels.cuh:15-38 -------------------------------- */
int gps_init_ones(void* d, uint64_t n, int dtype, void* stream);               /* InitOne  */
/* key(i) = hash(seed, first_index + i) & mask (uint32) or uniform [0,1) (float32);
 * same generator as oracle_init_random_* so host and device agree */
int gps_init_random(void* d, uint64_t n, int dtype, uint64_t seed, uint32_t mask,
                    uint64_t first_index, void* stream);
/* ValidateInclusive / ValidateExclusive: counts d[i] != first_value + i (+1 if
 * inclusive) on uint32; synchronous, *h_errs receives the count */
int gps_validate_ones(const void* d, uint64_t n, int kind, uint32_t first_value,
                      uint64_t* h_errs);
/* word-for-word compare of two device arrays (uint32), synchronous */
/* All-ones closed forms of the wide scans: GPS_U64 out[i] == i + 1 / i; GPS_U32X4 the flat
 * u32 view out[i] == (i >> 2) + 1 / (i >> 2) (GPUPrefixSumsWebGPUapis/SharedShaders/
 * validate.wgsl:50-60).  n counts elements.  Synchronous. */
int gps_validate_ones_wide(const void* d, uint64_t n, int kind, int dtype, uint64_t* h_errs);
/* Sort checker: *h_inversions = number of i with keys[i-1] > keys[i]; h_sums = two order-
 * independent fingerprints of the key multiset (compare input against output).  Synchronous. */
int gps_sort_fingerprint(const void* d_keys, uint64_t n, uint64_t* h_inversions, uint64_t h_sums[2]);
/* the same on uint64 keys (gps_sort_keys64) */
int gps_sort_fingerprint64(const void* d_keys, uint64_t n, uint64_t* h_inversions, uint64_t h_sums[2]);
int gps_compare_u32(const void* d_a, const void* d_b, uint64_t n, uint64_t* h_errs,
                    uint64_t* h_first_bad);
/* float32: counts |a-b| > rtol*|d_bound[i]| + atol (d_bound nullable -> |b|) */
int gps_compare_f32(const void* d_a, const void* d_b, const void* d_bound, uint64_t n,
                    float rtol, float atol, uint64_t* h_errs, float* h_max_abs_diff);

/* ---- memory / stream / timing helpers so host code needs no CUDA headers ------ */
int gps_set_device(int device);
int gps_malloc(void** d_ptr, uint64_t bytes);
int gps_free(void* d_ptr);
int gps_malloc_host(void** h_ptr, uint64_t bytes); /* pinned */
int gps_free_host(void* h_ptr);
int gps_memcpy_h2d(void* d_dst, const void* h_src, uint64_t bytes, void* stream);
int gps_memcpy_d2h(void* h_dst, const void* d_src, uint64_t bytes, void* stream);
int gps_memcpy_d2d(void* d_dst, const void* d_src, uint64_t bytes, void* stream);
int gps_memset(void* d, int byte, uint64_t bytes, void* stream);
int gps_stream_sync(void* stream);
int gps_device_sync(void);
int gps_event_create(void** ev);
int gps_event_destroy(void* ev);
int gps_event_record(void* ev, void* stream);
int gps_event_elapsed_ms(void* start, void* stop, float* ms); /* syncs on stop */

/* ---- CUDA IPC for the cross-process NVLink carry (one process per GPU) -------- */
#define GPS_IPC_HANDLE_BYTES 64
int gps_ipc_export(void* d_ptr, unsigned char handle[GPS_IPC_HANDLE_BYTES]);
int gps_ipc_open(const unsigned char handle[GPS_IPC_HANDLE_BYTES], void** d_ptr);
int gps_ipc_close(void* d_ptr);

/* ---- single-process sharded scan over several GPUs (C++ hosts) ---------------
 * New capability.  GPU g holds keys [g*n/G, (g+1)*n/G) in
 * d_in[g]/d_out[g].  Pass 1 reduces each slice and stores its total into every
 * downstream GPU's carry slots over NVLink (peer access); pass 2 scans each slice
 * with the carry folded into tile 0's look-back.  carry_mode: 0 = P2P stores,
 * 1 = host-staged exchange (debug).  Synchronous.  h_ms (nullable) receives the
 * max-over-GPUs device time.  Returns GPS_ERR_WATCHDOG when a look-back or a carry-slot
 * wait gave up on any GPU (results invalid). */
int gps_sharded_create(gps_shard_group** grp, int n_gpus, const int* devices,
                       uint64_t max_n_per_gpu);
void gps_sharded_destroy(gps_shard_group* grp);
int gps_sharded_scan(gps_shard_group* grp, int algo, int kind, int dtype,
                     const void* const* d_in, void* const* d_out, const uint64_t* n_per_gpu,
                     int carry_mode, float* h_ms);
/* The ctx of GPU `index` of the group (for gps_set_option, e.g. "host_chunk_keys"); NULL if out
 * of range.  Owned by the group. */
gps_ctx* gps_sharded_ctx(gps_shard_group* grp, int index);

/* The reference-facing call with HOST buffers on SEVERAL GPUs: one host array in, one host array
 * out (pinned for full PCIe bandwidth), all GPUs of the group working on it at once -- the
 * multi-GPU form of gps_scan_host, and like it the replacement of a dispatcher's
 * DispatchKernels* over caller memory (ChainedScanDecoupledLookbackDispatcher.cuh:177-195).
 * The array is cut into chunks of "host_chunk_keys" keys dealt round-robin to the GPUs; every
 * GPU pipelines copy-in, scan and copy-out of its chunks on three streams, and the scan of a
 * chunk takes its carry from the tagged running total its predecessor's last tile stored into
 * this GPU's memory over NVLink (no reduce pre-pass, no host synchronisation, no collective).
 * gps_sharded_scan_host waits; the _async form returns once everything is queued, consecutive
 * calls overlap their copies, gps_sharded_host_sync waits for all of them and reports
 * GPS_ERR_WATCHDOG if a look-back or carry-slot wait gave up.  h_total_out (nullable) receives
 * carry + sum of all keys. */
int gps_sharded_scan_host(gps_shard_group* grp, int algo, int kind, int dtype, const void* h_in,
                          void* h_out, uint64_t n, uint32_t carry_in_bits, void* h_total_out);
int gps_sharded_scan_host_async(gps_shard_group* grp, int algo, int kind, int dtype,
                                const void* h_in, void* h_out, uint64_t n, uint32_t carry_in_bits,
                                void* h_total_out);
int gps_sharded_host_sync(gps_shard_group* grp);
/* Which GPUs carry the host traffic of gps_sharded_scan_host (default: all of them, round-robin).
 * Host <-> device paths of a multi-socket box are neither equal nor independent (on the 8 x B200
 * VM of profiles/r02_pcie_aggregate.json all eight GPUs copying both ways total 72 GB/s per
 * direction, GPUs 4-7 alone 95), and the scan needs a fraction of one GPU, so fewer GPUs can be
 * faster.  gps_sharded_calibrate_host measures it, memcpy only, on the caller's buffers (h_src is
 * read, h_dst -- `bytes` each, pinned -- is overwritten with junk: pass the arrays of the scan
 * that follows): the whole group, then halves, descending while a half beats its parent by 3 %;
 * the winner takes the chunks from then on.  *h_gbs_each_way (nullable) = its copy rate per
 * direction.  Synchronous, a few hundred ms.  set / get: indices into the group's device list,
 * in chain order. */
int gps_sharded_calibrate_host(gps_shard_group* grp, const void* h_src, void* h_dst, uint64_t bytes,
                               float* h_gbs_each_way);
int gps_sharded_set_host_io(gps_shard_group* grp, const int* gpu_indices, int n);
int gps_sharded_get_host_io(gps_shard_group* grp, int* gpu_indices, int* n);

#ifdef __cplusplus
}
#endif
#endif /* GPS_B200_H_ */
