/*
 * oracle/scan_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement of the reference's serial prefix-sum validation loops.  Only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library; the product path (gpuprefixsums_b200/) never does.
 *
 * What it restates (paths relative to /root/reference):
 *   - inclusive serial scan: GPUPrefixSumsD3D12/Shaders/Utility.hlsl:99-111
 *     (ValidateRandomInclusive: `sum += in[i]; expect out[i] == sum`) and
 *     GPUPrefixSumsD3D12/Shaders/Survey.hlsl:49-54 (SerialInclusive).
 *   - exclusive serial scan: Utility.hlsl:113-125 (ValidateRandomExclusive:
 *     `expect out[i] == sum; sum += in[i]`) and Survey.hlsl:56-67.
 *   - all-ones closed form: GPUPrefixSumsCUDA/UtilityKernels.cuh:22-38
 *     (ValidateInclusive: scan[i] == i + 1, ValidateExclusive: scan[i] == i).
 *   - random test input: Utility.hlsl:33-66 (InitRandom, Hybrid Tausworthe,
 *     values & 1023, one RNG draw per "thread", 65536 threads, grid-stride).
 * All uint32 arithmetic wraps mod 2^32, exactly as the reference's `uint`.
 *
 * The CUDA directory of the reference has no host (CPU) scan at all, and no
 * float32 scan anywhere; the float32 oracle below is new: it accumulates in
 * float64 (a float32 serial sum stagnates above 2^25).
 *
 * Parity pinning: tests/test_oracle.py checks this file against (a) the
 * closed-form all-ones identity for the reference's own test sizes and (b)
 * tests/golden/ref_cuda_*.npz, outputs of the reference's own CUDA kernels
 * (oracle/_ref/libref_shim.so) run on a B200 (generator: oracle/gen_golden.py).
 */
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ---- serial scans (Utility.hlsl:99-125, Survey.hlsl:49-67) ------------------- */

/* returns the grand total (wraps mod 2^32); carry seeds the running sum */
uint32_t oracle_scan_u32(const uint32_t *in, uint32_t *out, uint64_t n, int exclusive,
                         uint32_t carry)
{
    uint32_t sum = carry;
    if (exclusive) {
        for (uint64_t i = 0; i < n; ++i) {
            const uint32_t t = in[i];
            out[i] = sum;
            sum += t;
        }
    } else {
        for (uint64_t i = 0; i < n; ++i) {
            sum += in[i];
            out[i] = sum;
        }
    }
    return sum;
}

/* float32 input, float64 accumulation and float64 output (the comparison is done
 * in float64 against rtol * running sum of |x|, written to out_abs if non-NULL) */
double oracle_scan_f32_f64(const float *in, double *out, double *out_abs, uint64_t n,
                           int exclusive, double carry)
{
    double sum = carry, asum = carry < 0 ? -carry : carry;
    for (uint64_t i = 0; i < n; ++i) {
        const double t = (double)in[i];
        if (exclusive) {
            out[i] = sum;
            if (out_abs) out_abs[i] = asum;
        }
        sum += t;
        asum += t < 0 ? -t : t;
        if (!exclusive) {
            out[i] = sum;
            if (out_abs) out_abs[i] = asum;
        }
    }
    return sum;
}

/* Checks a candidate float32 scan against the float64 serial scan of `in`:
 * counts positions with |cand - ref| > rtol * (carry + sum_{j<=i}|x_j|) + atol.
 * max_rel_out receives the largest observed |cand-ref| / max(sum|x|, tiny). */
uint64_t oracle_check_f32(const float *in, const float *cand, uint64_t n, int exclusive,
                          double carry, double rtol, double atol, double *max_rel_out)
{
    double sum = carry, asum = carry < 0 ? -carry : carry, max_rel = 0.0;
    uint64_t errs = 0;
    for (uint64_t i = 0; i < n; ++i) {
        const double t = (double)in[i];
        if (!exclusive) {
            sum += t;
            asum += t < 0 ? -t : t;
        }
        double d = (double)cand[i] - sum;
        if (d < 0) d = -d;
        if (!(d <= rtol * asum + atol)) ++errs; /* also catches NaN */
        const double rel = d / (asum > 1e-30 ? asum : 1e-30);
        if (rel > max_rel) max_rel = rel;
        if (exclusive) {
            sum += t;
            asum += t < 0 ? -t : t;
        }
    }
    if (max_rel_out) *max_rel_out = max_rel;
    return errs;
}

/* Word-for-word comparison of a candidate uint32 scan with the serial scan of
 * `in` (Utility.hlsl:99-125 run as a checker).  Returns the mismatch count and the
 * index of the first mismatch (or n). */
uint64_t oracle_check_u32(const uint32_t *in, const uint32_t *cand, uint64_t n, int exclusive,
                          uint32_t carry, uint64_t *first_bad)
{
    uint32_t sum = carry;
    uint64_t errs = 0, first = n;
    for (uint64_t i = 0; i < n; ++i) {
        if (!exclusive) sum += in[i];
        if (cand[i] != sum) {
            if (!errs) first = i;
            ++errs;
        }
        if (exclusive) sum += in[i];
    }
    if (first_bad) *first_bad = first;
    return errs;
}

/* ---- wide elements ------------------------------------------------ */
/* uint64: same serial loops, arithmetic mod 2^64 */
uint64_t oracle_scan_u64(const uint64_t *in, uint64_t *out, uint64_t n, int exclusive, uint64_t carry)
{
    uint64_t sum = carry;
    for (uint64_t i = 0; i < n; ++i) {
        if (exclusive) out[i] = sum;
        sum += in[i];
        if (!exclusive) out[i] = sum;
    }
    return sum;
}

/* four-lane struct scan: GPUPrefixSumsWebGPUapis/SharedShaders/csdldf_struct.wgsl scans an
 * array of vec4<u32> lane by lane; validate.wgsl:50-60 checks all-ones input against
 * out[i] == (i >> 2) + 1 over the flat u32 view.  n counts vec4 elements. */
void oracle_scan_u32x4(const uint32_t *in, uint32_t *out, uint64_t n, int exclusive, const uint32_t carry[4],
                       uint32_t total[4])
{
    uint32_t sum[4] = {carry[0], carry[1], carry[2], carry[3]};
    for (uint64_t i = 0; i < n; ++i)
        for (int l = 0; l < 4; ++l) {
            if (exclusive) out[4 * i + l] = sum[l];
            sum[l] += in[4 * i + l];
            if (!exclusive) out[4 * i + l] = sum[l];
        }
    for (int l = 0; l < 4; ++l) total[l] = sum[l];
}

/* float64: serial scan accumulated in long double (x87 80-bit on this platform), returns
 * the count of candidate values with |cand - ref| > rtol * sum|x| + atol */
uint64_t oracle_check_f64(const double *in, const double *cand, uint64_t n, int exclusive, double carry,
                          double rtol, double atol, double *max_rel_out)
{
    long double sum = carry, asum = carry < 0 ? -carry : carry, max_rel = 0.0L;
    uint64_t errs = 0;
    for (uint64_t i = 0; i < n; ++i) {
        const long double t = in[i];
        if (!exclusive) {
            sum += t;
            asum += t < 0 ? -t : t;
        }
        long double d = (long double)cand[i] - sum;
        if (d < 0) d = -d;
        if (!(d <= (long double)rtol * asum + atol)) ++errs;
        const long double rel = d / (asum > 1e-300L ? asum : 1e-300L);
        if (rel > max_rel) max_rel = rel;
        if (exclusive) {
            sum += t;
            asum += t < 0 ? -t : t;
        }
    }
    if (max_rel_out) *max_rel_out = (double)max_rel;
    return errs;
}

/* ---- radix sort consumer -------------------------------------------- */
/* Stable LSD radix sort of uint32 keys on their low `end_bit` bits, 8 bits per pass, with an
 * optional 32-bit payload: the textbook counting-sort-per-digit loop (count, exclusive scan
 * of the 256 counts -- the prefix sum the reference's README.md:7 says its scans exist for --
 * then a stable scatter).  The sort itself is not part of /root/reference (sibling repository
 * GPUSorting), so this is pinned in tests/test_oracle.py against numpy's stable sort instead
 * of reference vectors.  tmp_* are caller-provided work arrays of n elements; the result
 * lands in keys_out / vals_out.  vals_* may all be NULL. */
/* int32 / float32 keys: the usual order-preserving maps onto uint32 (flip the sign bit; for
 * floats flip every bit of negative values) around the same radix passes */
static uint32_t oracle_key_map(uint32_t k, int type)
{
    const int t = type & 0xFF;
    if (t == 1) k ^= 0x80000000u;
    if (t == 2) k ^= ((k >> 31) ? 0xFFFFFFFFu : 0x80000000u);
    return (type & 0x100) ? ~k : k; /* 0x100: descending = complement of the ascending map */
}
static uint32_t oracle_key_unmap(uint32_t k, int type)
{
    const int t = type & 0xFF;
    if (type & 0x100) k = ~k;
    if (t == 1) k ^= 0x80000000u;
    if (t == 2) k ^= ((k >> 31) ? 0x80000000u : 0xFFFFFFFFu);
    return k;
}
void oracle_sort_u32(const uint32_t *keys_in, uint32_t *keys_out, uint32_t *keys_tmp, const uint32_t *vals_in,
                     uint32_t *vals_out, uint32_t *vals_tmp, uint64_t n, int end_bit);
/* key_type 0 = uint32, 1 = int32, 2 = float32 (bit patterns); keys_map is a work array of n keys */
void oracle_sort_keys32(int key_type, const uint32_t *keys_in, uint32_t *keys_out, uint32_t *keys_tmp,
                        uint32_t *keys_map, const uint32_t *vals_in, uint32_t *vals_out, uint32_t *vals_tmp, uint64_t n)
{
    for (uint64_t i = 0; i < n; ++i) keys_map[i] = oracle_key_map(keys_in[i], key_type);
    oracle_sort_u32(keys_map, keys_out, keys_tmp, vals_in, vals_out, vals_tmp, n, 32);
    for (uint64_t i = 0; i < n; ++i) keys_out[i] = oracle_key_unmap(keys_out[i], key_type);
}

void oracle_sort_u32(const uint32_t *keys_in, uint32_t *keys_out, uint32_t *keys_tmp, const uint32_t *vals_in,
                     uint32_t *vals_out, uint32_t *vals_tmp, uint64_t n, int end_bit)
{
    /* stable order of key & (2^end_bit - 1): the last digit covers end_bit % 8 bits when that is
     * not 0 (tests/test_oracle.py pins this on numpy's stable argsort of the masked keys) */
    const int passes = (end_bit + 7) / 8;
    const uint32_t *sk = keys_in, *sv = vals_in;
    for (int pass = 0; pass < passes; ++pass) {
        const int to_out = ((passes - 1 - pass) & 1) == 0;
        uint32_t *dk = to_out ? keys_out : keys_tmp, *dv = to_out ? vals_out : vals_tmp;
        const int shift = 8 * pass;
        const uint32_t mask = (pass == passes - 1 && (end_bit & 7)) ? ((1u << (end_bit & 7)) - 1u) : 0xFFu;
        uint64_t count[256] = {0}, sum = 0;
        for (uint64_t i = 0; i < n; ++i) ++count[(sk[i] >> shift) & mask];
        for (int d = 0; d < 256; ++d) {
            const uint64_t c = count[d];
            count[d] = sum;
            sum += c;
        }
        for (uint64_t i = 0; i < n; ++i) {
            const uint64_t pos = count[(sk[i] >> shift) & mask]++;
            dk[pos] = sk[i];
            if (sv) dv[pos] = sv[i];
        }
        sk = dk;
        sv = sv ? dv : sv;
    }
}

/* ---- 64-bit keys: the same textbook LSD loop, 8 bits per pass; key_type 0 = uint64, 1 = int64,
 * 2 = float64 bit patterns, | 0x100 = descending (the maps of the 32-bit sort on 64 bits) ---- */
static uint64_t oracle_key_map64(uint64_t k, int type)
{
    const int t = type & 0xFF;
    if (t == 1) k ^= 0x8000000000000000ull;
    if (t == 2) k ^= ((k >> 63) ? 0xFFFFFFFFFFFFFFFFull : 0x8000000000000000ull);
    return (type & 0x100) ? ~k : k;
}
static uint64_t oracle_key_unmap64(uint64_t k, int type)
{
    const int t = type & 0xFF;
    if (type & 0x100) k = ~k;
    if (t == 1) k ^= 0x8000000000000000ull;
    if (t == 2) k ^= ((k >> 63) ? 0x8000000000000000ull : 0xFFFFFFFFFFFFFFFFull);
    return k;
}
/* keys_tmp: work array of n keys; vals (nullable, all three or none): 32-bit payload */
void oracle_sort_keys64(int key_type, const uint64_t *keys_in, uint64_t *keys_out, uint64_t *keys_tmp,
                        const uint32_t *vals_in, uint32_t *vals_out, uint32_t *vals_tmp, uint64_t n, int end_bit)
{
    const int passes = (end_bit + 7) / 8;
    if (passes == 0 || n == 0) return;
    /* pass 0 reads the mapped input; ping-pong so that the last pass lands in keys_out */
    uint64_t *bufs[2] = {keys_out, keys_tmp};
    uint32_t *vbufs[2] = {vals_out, vals_tmp};
    int cur = (passes & 1) ? 1 : 0; /* index of the buffer holding the CURRENT source after the initial map */
    uint64_t *sk = bufs[cur];
    uint32_t *sv = vals_in ? vbufs[cur] : 0;
    for (uint64_t i = 0; i < n; ++i) {
        sk[i] = oracle_key_map64(keys_in[i], key_type);
        if (sv) sv[i] = vals_in[i];
    }
    for (int pass = 0; pass < passes; ++pass) {
        uint64_t *dk = bufs[cur ^ 1];
        uint32_t *dv = sv ? vbufs[cur ^ 1] : 0;
        const int shift = 8 * pass;
        const uint32_t mask = (pass == passes - 1 && (end_bit & 7)) ? ((1u << (end_bit & 7)) - 1u) : 0xFFu;
        uint64_t count[256] = {0}, sum = 0;
        for (uint64_t i = 0; i < n; ++i) ++count[(sk[i] >> shift) & mask];
        for (int d = 0; d < 256; ++d) {
            const uint64_t c = count[d];
            count[d] = sum;
            sum += c;
        }
        for (uint64_t i = 0; i < n; ++i) {
            const uint64_t pos = count[(sk[i] >> shift) & mask]++;
            dk[pos] = sk[i];
            if (sv) dv[pos] = sv[i];
        }
        cur ^= 1;
        sk = dk;
        sv = dv;
    }
    /* passes flips from the start buffer: ends in bufs[0] == keys_out */
    for (uint64_t i = 0; i < n; ++i) keys_out[i] = oracle_key_unmap64(keys_out[i], key_type);
}

/* ---- all-ones closed form (UtilityKernels.cuh:22-38) ------------------------- */
uint64_t oracle_validate_ones(const uint32_t *scan, uint64_t n, int exclusive)
{
    uint64_t errs = 0;
    for (uint64_t i = 0; i < n; ++i)
        errs += scan[i] != (uint32_t)(exclusive ? i : i + 1);
    return errs;
}

/* ---- reference random test input (Utility.hlsl:38-66) ------------------------ */
/* One Hybrid-Tausworthe draw per dispatch thread id (65536 threads = 256 groups x
 * 256), computed once outside the grid-stride loop exactly as the shader does, so
 * element i receives the value of thread (i mod 65536). */
void oracle_init_random_d3d12(uint32_t *out, uint64_t n, uint32_t seed)
{
    const uint32_t inc = 256u * 256u;
    for (uint32_t id = 0; id < inc && id < n; ++id) {
        uint32_t z1 = (id << 2) * seed;
        uint32_t z2 = ((id << 2) + 1) * seed;
        uint32_t z3 = ((id << 2) + 2) * seed;
        uint32_t z4 = ((id << 2) + 3) * seed;
        z1 = ((z1 & 4294967294U) << 12) ^ (((z1 << 13) ^ z1) >> 19);
        z2 = ((z2 & 4294967288U) << 4) ^ (((z2 << 2) ^ z2) >> 25);
        z3 = ((z3 & 4294967280U) << 17) ^ (((z3 << 3) ^ z3) >> 11);
        z4 = z4 * 1664525u + 1013904223U;
        const uint32_t t = (z1 ^ z2 ^ z3 ^ z4) & 1023u;
        for (uint64_t i = id; i < n; i += inc) out[i] = t;
    }
}

/* ---- seeded test input shared with the CUDA side ----------------------------- */
/* Counter-based: key(i) = mix64(seed + i * golden) -- identical to
 * gps::random_key() in gpuprefixsums_b200/csrc/gps_util_kernels.cuh, so host and
 * device can generate the same array independently (needed for the 2^33 sharded
 * case where the host never materialises the input). */
static inline uint64_t mix64(uint64_t z)
{
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}

uint32_t oracle_random_key(uint64_t seed, uint64_t i, uint32_t mask)
{
    return (uint32_t)(mix64(seed + (i + 1) * 0x9E3779B97F4A7C15ULL) >> 32) & mask;
}

void oracle_init_random_u32(uint32_t *out, uint64_t n, uint64_t seed, uint32_t mask,
                            uint64_t first_index)
{
    for (uint64_t i = 0; i < n; ++i) out[i] = oracle_random_key(seed, first_index + i, mask);
}

/* float32 uniform in [0,1): top 24 bits of the same hash */
void oracle_init_random_f32(float *out, uint64_t n, uint64_t seed, uint64_t first_index)
{
    for (uint64_t i = 0; i < n; ++i)
        out[i] = (float)(oracle_random_key(seed, first_index + i, 0xFFFFFFFFu) >> 8) *
                 (1.0f / 16777216.0f);
}

/* ---- multi-threaded host scan (CPU baseline "with all the host threads") ----- */
/* Chunked two-pass restatement: each thread runs the serial loop above on a
 * contiguous chunk (pass 1: chunk totals, pass 2: serial scan seeded with the sum of
 * the preceding chunk totals).  Bit-identical to oracle_scan_u32 for uint32. */
typedef struct {
    const uint32_t *in;
    uint32_t *out;
    uint64_t begin, end;
    int exclusive;
    uint32_t total, carry;
} chunk_job;

static void *chunk_reduce(void *arg)
{
    chunk_job *j = (chunk_job *)arg;
    uint32_t s = 0;
    for (uint64_t i = j->begin; i < j->end; ++i) s += j->in[i];
    j->total = s;
    return NULL;
}

static void *chunk_scan(void *arg)
{
    chunk_job *j = (chunk_job *)arg;
    oracle_scan_u32(j->in + j->begin, j->out + j->begin, j->end - j->begin, j->exclusive,
                    j->carry);
    return NULL;
}

uint32_t oracle_scan_u32_mt(const uint32_t *in, uint32_t *out, uint64_t n, int exclusive,
                            uint32_t carry, int threads)
{
    if (threads <= 1 || n < (uint64_t)threads * 4096)
        return oracle_scan_u32(in, out, n, exclusive, carry);
    chunk_job *jobs = (chunk_job *)calloc((size_t)threads, sizeof(chunk_job));
    pthread_t *tids = (pthread_t *)calloc((size_t)threads, sizeof(pthread_t));
    const uint64_t per = (n + (uint64_t)threads - 1) / (uint64_t)threads;
    for (int t = 0; t < threads; ++t) {
        jobs[t].in = in;
        jobs[t].out = out;
        jobs[t].begin = per * (uint64_t)t < n ? per * (uint64_t)t : n;
        jobs[t].end = per * (uint64_t)(t + 1) < n ? per * (uint64_t)(t + 1) : n;
        jobs[t].exclusive = exclusive;
    }
    for (int t = 0; t < threads; ++t) pthread_create(&tids[t], NULL, chunk_reduce, &jobs[t]);
    for (int t = 0; t < threads; ++t) pthread_join(tids[t], NULL);
    uint32_t run = carry;
    for (int t = 0; t < threads; ++t) {
        jobs[t].carry = run;
        run += jobs[t].total;
    }
    for (int t = 0; t < threads; ++t) pthread_create(&tids[t], NULL, chunk_scan, &jobs[t]);
    for (int t = 0; t < threads; ++t) pthread_join(tids[t], NULL);
    free(jobs);
    free(tids);
    return run;
}
