// tools/gps_tune.cu -- tuning / comparison harness (developer tool, not the product).
// Sweeps kernel configurations of the register and TMA CSDL paths, RTS, copy
// ceilings (cudaMemcpy D2D, LDG/STG copy kernel, TMA bulk copy kernel) and
// cub::DeviceScan on one GPU; verifies every variant against CUB's output word for word
// and prints one JSON object per variant.
//   gps_tune [--log2n 28] [--n N] [--iters 50] [--filter substr] [--float] [--exclusive]
#include <cub/device/device_scan.cuh>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <memory>
#include <string>
#include <vector>

#include "../gpuprefixsums_b200/csrc/gps_launch.cuh"
#include "../gpuprefixsums_b200/csrc/gps_util_kernels.cuh"

using namespace gps;

#define CK(call)                                                                      \
    do {                                                                              \
        cudaError_t e_ = (call);                                                      \
        if (e_ != cudaSuccess) {                                                      \
            fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
            exit(2);                                                                  \
        }                                                                             \
    } while (0)

// ---- copy ceilings ----------------------------------------------------------------
template <int NT, int V>
__global__ void __launch_bounds__(NT) copy_ldg_kernel(const uint4* in, uint4* out, uint64_t nvec)
{
    const uint64_t base = (uint64_t)blockIdx.x * (NT * V) + (threadIdx.x >> 5) * (32 * V) + (threadIdx.x & 31);
    uint4 r[V];
#pragma unroll
    for (int k = 0; k < V; ++k)
        if (base + 32 * k < nvec) r[k] = ldg_stream(in + base + 32 * k);
#pragma unroll
    for (int k = 0; k < V; ++k)
        if (base + 32 * k < nvec) stg_stream(out + base + 32 * k, r[k]);
}

// persistent TMA bulk copy: global -> smem ring -> global, one thread drives it
template <int TILE_BYTES, int STAGES>
__global__ void __launch_bounds__(32) copy_tma_kernel(const unsigned char* in, unsigned char* out,
                                                      uint32_t num_tiles, uint32_t* counter,
                                                      uint32_t counter_base)
{
    extern __shared__ __align__(128) unsigned char ring[];
    __shared__ __align__(8) uint64_t s_full[STAGES];
    if (threadIdx.x != 0) return;
    for (int s = 0; s < STAGES; ++s) mbar_init(&s_full[s], 1);
    fence_mbar_init();
    uint32_t tiles[STAGES];
    auto produce = [&](int stage) {
        const uint32_t t = atomicAdd(counter, 1u) - counter_base;
        tiles[stage] = t;
        if (t < num_tiles) {
            mbar_arrive_expect_tx(&s_full[stage], TILE_BYTES);
            bulk_load(ring + (size_t)stage * TILE_BYTES, in + (uint64_t)t * TILE_BYTES, TILE_BYTES, &s_full[stage]);
        }
    };
    // static register indexing is not needed here (one thread, local array is fine)
    for (int s = 0; s < STAGES - 1; ++s) produce(s);
    for (uint32_t it = 0;; ++it) {
        const int stage = it % STAGES;
        const uint32_t tile = tiles[stage];
        if (tile >= num_tiles) break;
        mbar_wait(&s_full[stage], (it / STAGES) & 1u);
        bulk_store(out + (uint64_t)tile * TILE_BYTES, ring + (size_t)stage * TILE_BYTES, TILE_BYTES);
        bulk_commit();
        bulk_wait_read<1>();
        produce((it + STAGES - 1) % STAGES);
    }
    bulk_wait<0>();
}

// ---- harness ------------------------------------------------------------------------
struct Env {
    uint64_t n = 1ull << 28;
    int iters = 50;
    bool is_float = false;
    bool exclusive = false;
    void *in = nullptr, *out = nullptr, *ref = nullptr, *work = nullptr;
    Scratch s;
    uint32_t* tma_copy_counter = nullptr;
    uint32_t tma_copy_base = 0;
    cudaEvent_t e0, e1;
    double copy_gbs = 0;
    bool prof_enabled = false;
    unsigned long long* prof_buf = nullptr;
    double sm_ghz = 1.965;
    void* flush_buf = nullptr;  // --flush: written before every timed launch (cold L2), launches timed one by one
    size_t flush_bytes = 0;
};

struct Result {
    double ms_mean = 0;
    uint64_t errs = 0;
    bool verified = false;
    bool ok = true;
};

static uint64_t compare(Env& env, const void* got)
{
    unsigned long long* d_err;
    unsigned long long init[2] = {0ull, ~0ull}, h[2];
    CK(cudaMalloc(&d_err, sizeof(init)));
    CK(cudaMemcpy(d_err, init, sizeof(init), cudaMemcpyHostToDevice));
    if (!env.is_float) {
        compare_u32_kernel<<<2048, 256>>>((const uint32_t*)got, (const uint32_t*)env.ref, env.n, d_err);
    } else {
        // float: tolerance vs CUB (a peer, not an oracle): rtol 1e-4 of the value
        compare_f32_kernel<<<2048, 256>>>((const float*)got, (const float*)env.ref, nullptr, env.n, 1e-4f, 1e-3f,
                                          d_err, (uint32_t*)(d_err + 1));
    }
    CK(cudaMemcpy(h, d_err, sizeof(h), cudaMemcpyDeviceToHost));
    CK(cudaFree(d_err));
    if (h[0] && !env.is_float) fprintf(stderr, "    first mismatch at %llu\n", h[1]);
    return h[0];
}

using LaunchFn = std::function<cudaError_t(const void* in, void* out)>;

static Result run_variant(Env& env, const char* name, bool verify, const LaunchFn& fn, double bytes_per_key = 8.0,
                          bool inplace_too = true)
{
    Result r;
    // correctness: out-of-place once, then in place once
    CK(cudaMemset(env.out, 0xAB, env.n * 4));
    cudaError_t e = fn(env.in, env.out);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
        printf("{\"variant\": \"%s\", \"error\": \"%s\"}\n", name, cudaGetErrorString(e));
        fflush(stdout);
        r.ok = false;
        return r;
    }
    uint32_t status = 0;
    CK(cudaMemcpy(&status, env.s.status, 4, cudaMemcpyDeviceToHost));
    if (status) CK(cudaMemset(env.s.status, 0, 4));
    if (verify) {
        r.errs = compare(env, env.out);
        r.verified = true;
        if (inplace_too) {
            CK(cudaMemcpy(env.work, env.in, env.n * 4, cudaMemcpyDeviceToDevice));
            e = fn(env.work, env.work);
            if (e == cudaSuccess) e = cudaDeviceSynchronize();
            if (e != cudaSuccess) {
                printf("{\"variant\": \"%s\", \"error\": \"in-place: %s\"}\n", name, cudaGetErrorString(e));
                r.ok = false;
                return r;
            }
            r.errs += compare(env, env.work);
        }
    }
    // timing: out-of-place, back to back (2 GiB of traffic per launch at 2^28 >> L2)
    for (int i = 0; i < 3; ++i) fn(env.in, env.out);
    CK(cudaDeviceSynchronize());
    float ms = 0;
    if (env.flush_buf) {  // cold L2: flush, then time ONE launch with its own pair of events
        double sum = 0;
        for (int i = 0; i < env.iters; ++i) {
            CK(cudaMemsetAsync(env.flush_buf, i & 0xFF, env.flush_bytes));
            CK(cudaEventRecord(env.e0));
            fn(env.in, env.out);
            CK(cudaEventRecord(env.e1));
            CK(cudaEventSynchronize(env.e1));
            CK(cudaEventElapsedTime(&ms, env.e0, env.e1));
            sum += ms;
        }
        r.ms_mean = sum / env.iters;
    } else {
        CK(cudaEventRecord(env.e0));
        for (int i = 0; i < env.iters; ++i) fn(env.in, env.out);
        CK(cudaEventRecord(env.e1));
        CK(cudaEventSynchronize(env.e1));
        CK(cudaEventElapsedTime(&ms, env.e0, env.e1));
        r.ms_mean = ms / env.iters;
    }
    // in place timing
    double ms_inplace = 0;
    if (inplace_too) {
        for (int i = 0; i < 2; ++i) fn(env.work, env.work);
        CK(cudaDeviceSynchronize());
        CK(cudaEventRecord(env.e0));
        for (int i = 0; i < env.iters; ++i) fn(env.work, env.work);
        CK(cudaEventRecord(env.e1));
        CK(cudaEventSynchronize(env.e1));
        CK(cudaEventElapsedTime(&ms, env.e0, env.e1));
        ms_inplace = ms / env.iters;
    }
    uint32_t status2 = 0;
    CK(cudaMemcpy(&status2, env.s.status, 4, cudaMemcpyDeviceToHost));
    if (status2) CK(cudaMemset(env.s.status, 0, 4));
    status |= status2;
    const double gkeys = env.n / (r.ms_mean * 1e-3) / 1e9;
    const double gbs = gkeys * bytes_per_key;
    // one extra launch with the in-kernel profile counters enabled (kernels that support them)
    double prof_us[kProfFields] = {0};
    if (env.prof_enabled) {
        CK(cudaMemset(env.prof_buf, 0, sizeof(unsigned long long) * kProfSlots * kProfFields));
        env.s.prof = env.prof_buf;
        fn(env.in, env.out);
        CK(cudaDeviceSynchronize());
        env.s.prof = nullptr;
        std::vector<unsigned long long> h(kProfSlots * kProfFields);
        CK(cudaMemcpy(h.data(), env.prof_buf, h.size() * 8, cudaMemcpyDeviceToHost));
        unsigned long long tot[kProfFields] = {0};
        for (int r = 0; r < kProfSlots; ++r)
            for (int f = 0; f < kProfFields; ++f) tot[f] += h[r * kProfFields + f];
        if (tot[PROF_TILES]) {
            for (int f = 1; f <= PROF_FINISH; ++f) prof_us[f] = (double)tot[f] / tot[PROF_TILES] / env.sm_ghz / 1e3;
            prof_us[PROF_WINDOWS] = (double)tot[PROF_WINDOWS] / tot[PROF_TILES];
            prof_us[PROF_POLLS] = (double)tot[PROF_POLLS] / tot[PROF_TILES];
            prof_us[PROF_TILES] = (double)tot[PROF_TILES];
        }
    }
    if (prof_us[PROF_TILES] > 0)
        printf("{\"profile\": \"%s\", \"tiles\": %.0f, \"us_claim\": %.3f, \"us_load\": %.3f, \"us_reduce\": %.3f, "
               "\"us_lookback\": %.3f, \"us_finish\": %.3f, \"windows\": %.2f, \"polls\": %.2f}\n",
               name, prof_us[PROF_TILES], prof_us[PROF_CLAIM], prof_us[PROF_LOAD], prof_us[PROF_REDUCE],
               prof_us[PROF_LOOKBACK], prof_us[PROF_FINISH], prof_us[PROF_WINDOWS], prof_us[PROF_POLLS]);
    printf("{\"variant\": \"%s\", \"n\": %llu, \"ms\": %.5f, \"gkeys_s\": %.2f, \"gb_s\": %.1f, "
           "\"ms_inplace\": %.5f, \"gkeys_s_inplace\": %.2f, \"verified\": %s, \"errors\": %llu, \"watchdog\": %u}\n",
           name, (unsigned long long)env.n, r.ms_mean, gkeys, gbs, ms_inplace,
           ms_inplace > 0 ? env.n / (ms_inplace * 1e-3) / 1e9 : 0.0, r.verified ? "true" : "false",
           (unsigned long long)r.errs, status);
    fflush(stdout);
    return r;
}

template <typename T, int NT, int V, bool INCL, int WIN = 2>
static void add_reg(Env& env, std::vector<std::pair<std::string, LaunchFn>>& v, const char* suffix = "")
{
    char name[96];
    snprintf(name, sizeof(name), "csdl_reg<%d,%d>/w%d%s", NT, V, WIN, suffix);
    v.emplace_back(name, [&env](const void* in, void* out) {
        CarrySpec cs;
        if (env.s.prof) return launch_csdl_reg<T, NT, V, INCL, WIN, true>(env.s, in, out, env.n, cs, 0);
        return launch_csdl_reg<T, NT, V, INCL, WIN, false>(env.s, in, out, env.n, cs, 0);
    });
}

template <typename T, int NT, int V, bool INCL, int G = 4, int REGS = 40>
static void add_smem(Env& env, std::vector<std::pair<std::string, LaunchFn>>& v)
{
    char name[96];
    snprintf(name, sizeof(name), "csdl_smem<%d,%d>/g%d/r%d", NT, V, G, REGS);
    auto cfg = std::make_shared<bool>(false);
    v.emplace_back(name, [&env, cfg](const void* in, void* out) {
        CarrySpec cs;
        return launch_csdl_smem<T, NT, V, INCL, 1, G, REGS>(env.s, in, out, env.n, cs, *cfg, 0);
    });
}

template <typename T, int NT, int STAGES, bool INCL, int WIN = 2>
static void add_tma(Env& env, std::vector<std::pair<std::string, LaunchFn>>& v, int max_ctas)
{
    char name[96];
    snprintf(name, sizeof(name), "csdl_tma<%d,S%d>/w%d/cta%d", NT, STAGES, WIN, max_ctas);
    auto occ = std::make_shared<int>(0);
    auto occp = std::make_shared<int>(0);
    v.emplace_back(name, [&env, occ, occp, max_ctas](const void* in, void* out) {
        CarrySpec cs;
        if (env.s.prof)
            return launch_csdl_tma<T, NT, STAGES, INCL, WIN, true>(env.s, in, out, env.n, cs, *occp, max_ctas, 0);
        return launch_csdl_tma<T, NT, STAGES, INCL, WIN, false>(env.s, in, out, env.n, cs, *occ, max_ctas, 0);
    });
}

template <typename T, int NT, int V, bool INCL>
static void add_rts(Env& env, std::vector<std::pair<std::string, LaunchFn>>& v)
{
    char name[96];
    snprintf(name, sizeof(name), "rts<%d,%d>", NT, V);
    v.emplace_back(name, [&env](const void* in, void* out) {
        CarrySpec cs;
        return launch_rts<T, NT, V, INCL>(env.s, in, out, env.n, cs, 0);
    });
}

template <typename T, int NT, int V, bool INCL>
static void add_fallback(Env& env, std::vector<std::pair<std::string, LaunchFn>>& v, uint32_t mask, uint32_t spin)
{
    char name[96];
    snprintf(name, sizeof(name), "csdl_fallback<%d,%d>/mask%x/spin%u", NT, V, mask, spin);
    v.emplace_back(name, [&env, mask, spin](const void* in, void* out) {
        CarrySpec cs;
        FallbackParams f{mask, spin, nullptr};
        return launch_csdl_fallback<T, NT, V, INCL>(env.s, in, out, env.n, cs, f, 0);
    });
}

template <typename T, bool INCL>
static void build_variants(Env& env, std::vector<std::pair<std::string, LaunchFn>>& v)
{
    add_reg<T, 256, 4, INCL, 1>(env, v);
    add_reg<T, 128, 4, INCL, 1>(env, v);
    add_reg<T, 256, 8, INCL, 1>(env, v);
    add_reg<T, 128, 10, INCL, 1>(env, v);
    add_reg<T, 128, 12, INCL, 1>(env, v);
    add_reg<T, 128, 16, INCL, 1>(env, v);
    add_reg<T, 96, 16, INCL, 1>(env, v);
    add_reg<T, 96, 20, INCL, 1>(env, v);
    add_reg<T, 64, 16, INCL, 1>(env, v);
    add_reg<T, 64, 20, INCL, 1>(env, v);
    add_reg<T, 64, 24, INCL, 1>(env, v);
    add_reg<T, 64, 28, INCL, 1>(env, v);
    add_reg<T, 64, 32, INCL, 1>(env, v);
    add_reg<T, 32, 24, INCL, 1>(env, v);
    add_reg<T, 32, 32, INCL, 1>(env, v);
    add_reg<T, 32, 40, INCL, 1>(env, v);
    add_smem<T, 128, 8, INCL>(env, v);
    add_smem<T, 128, 8, INCL, 2>(env, v);
    add_smem<T, 128, 8, INCL, 8, 48>(env, v);
    add_smem<T, 128, 8, INCL, 4, 32>(env, v);
    add_smem<T, 128, 8, INCL, 2, 32>(env, v);
    add_smem<T, 128, 12, INCL>(env, v);
    add_smem<T, 128, 12, INCL, 4, 48>(env, v);
    add_smem<T, 128, 12, INCL, 6, 48>(env, v);
    add_smem<T, 128, 16, INCL>(env, v);
    add_smem<T, 96, 8, INCL>(env, v);
    add_smem<T, 96, 12, INCL>(env, v);
    add_smem<T, 160, 8, INCL>(env, v);
    add_smem<T, 192, 8, INCL>(env, v);
    add_smem<T, 256, 8, INCL>(env, v);
    add_smem<T, 256, 4, INCL, 4, 32>(env, v);
    add_smem<T, 64, 16, INCL>(env, v);
    add_tma<T, 512, 2, INCL, 1>(env, v, 0);
    add_tma<T, 256, 2, INCL, 1>(env, v, 0);
    add_tma<T, 384, 2, INCL, 1>(env, v, 0);
    add_tma<T, 768, 2, INCL, 1>(env, v, 0);
    add_rts<T, 256, 4, INCL>(env, v);
    add_rts<T, 512, 4, INCL>(env, v);
    add_rts<T, 256, 8, INCL>(env, v);
    add_rts<T, 128, 12, INCL>(env, v);
    add_rts<T, 128, 16, INCL>(env, v);
    add_rts<T, 64, 24, INCL>(env, v);
    add_fallback<T, 128, 12, INCL>(env, v, kNoStall, 64);
    add_fallback<T, 128, 12, INCL>(env, v, kNoStall, 4);
    add_fallback<T, 128, 12, INCL>(env, v, 7, 4);
    add_fallback<T, 128, 12, INCL>(env, v, 63, 4);
    add_fallback<T, 256, 4, INCL>(env, v, 7, 4);
}

template <typename T, bool INCL>
static cudaError_t cub_scan(const void* in, void* out, uint64_t n, void*& temp, size_t& temp_bytes)
{
    size_t need = 0;
    if (INCL)
        cub::DeviceScan::InclusiveSum(nullptr, need, (const T*)in, (T*)out, (int64_t)n);
    else
        cub::DeviceScan::ExclusiveSum(nullptr, need, (const T*)in, (T*)out, (int64_t)n);
    if (need > temp_bytes) {
        if (temp) cudaFree(temp);
        cudaMalloc(&temp, need);
        temp_bytes = need;
    }
    if (INCL)
        return cub::DeviceScan::InclusiveSum(temp, need, (const T*)in, (T*)out, (int64_t)n);
    return cub::DeviceScan::ExclusiveSum(temp, need, (const T*)in, (T*)out, (int64_t)n);
}

template <typename T, bool INCL>
static int run_all(Env& env, const char* filter)
{
    static void* cub_temp = nullptr;
    static size_t cub_temp_bytes = 0;
    auto wanted = [&](const std::string& name) { return !filter || name.find(filter) != std::string::npos; };
    const uint64_t nvec = env.n / 4;

    // reference output from CUB
    CK((cub_scan<T, INCL>(env.in, env.ref, env.n, cub_temp, cub_temp_bytes)));
    CK(cudaDeviceSynchronize());

    if (wanted("copy_memcpy"))
        run_variant(env, "copy_memcpy_d2d", false,
                    [&](const void* in, void* out) { return cudaMemcpyAsync(out, in, env.n * 4, cudaMemcpyDeviceToDevice, 0); },
                    8.0, false);
    if (wanted("copy_ldg")) {
        run_variant(env, "copy_ldg<256,4>", false, [&](const void* in, void* out) {
            copy_ldg_kernel<256, 4><<<(unsigned)div_up(nvec, 256 * 4), 256>>>((const uint4*)in, (uint4*)out, nvec);
            return cudaGetLastError();
        }, 8.0, false);
        run_variant(env, "copy_ldg<512,4>", false, [&](const void* in, void* out) {
            copy_ldg_kernel<512, 4><<<(unsigned)div_up(nvec, 512 * 4), 512>>>((const uint4*)in, (uint4*)out, nvec);
            return cudaGetLastError();
        }, 8.0, false);
        run_variant(env, "copy_ldg<256,8>", false, [&](const void* in, void* out) {
            copy_ldg_kernel<256, 8><<<(unsigned)div_up(nvec, 256 * 8), 256>>>((const uint4*)in, (uint4*)out, nvec);
            return cudaGetLastError();
        }, 8.0, false);
    }
    if (wanted("copy_tma") && env.n % 8192 == 0) {
        auto tma_copy = [&](auto kern, int tile_bytes, int stages, int ctas_per_sm, const char* name) {
            const int smem = tile_bytes * stages;
            CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
            const uint32_t tiles = (uint32_t)(env.n * 4 / tile_bytes);
            const uint32_t grid = 148 * ctas_per_sm;
            run_variant(env, name, false, [&, kern, tiles, grid, smem, stages](const void* in, void* out) {
                kern<<<grid, 32, smem>>>((const unsigned char*)in, (unsigned char*)out, tiles, env.tma_copy_counter,
                                         env.tma_copy_base);
                env.tma_copy_base += grid * (stages - 1) + tiles;
                return cudaGetLastError();
            }, 8.0, false);
        };
        tma_copy(copy_tma_kernel<32768, 3>, 32768, 3, 2, "copy_tma<32K,S3>x2");
        tma_copy(copy_tma_kernel<16384, 4>, 16384, 4, 3, "copy_tma<16K,S4>x3");
        tma_copy(copy_tma_kernel<32768, 6>, 32768, 6, 1, "copy_tma<32K,S6>x1");
        tma_copy(copy_tma_kernel<8192, 4>, 8192, 4, 6, "copy_tma<8K,S4>x6");
    }
    if (wanted("cub"))
        run_variant(env, "cub_device_scan", true, [&](const void* in, void* out) {
            return cub_scan<T, INCL>(in, out, env.n, cub_temp, cub_temp_bytes);
        });

    if (filter && !strcmp(filter, "sweep")) {
        std::vector<std::pair<std::string, LaunchFn>> sv;
        static int occ_t = 0;
        sv.emplace_back("sweep/cub_device_scan", [&](const void* in, void* out) {
            return cub_scan<T, INCL>(in, out, env.n, cub_temp, cub_temp_bytes);
        });
        sv.emplace_back("sweep/csdl_reg<128,12>", [&](const void* in, void* out) {
            CarrySpec cs;
            return launch_csdl_reg<T, 128, 12, INCL, 1>(env.s, in, out, env.n, cs, 0);
        });
        sv.emplace_back("sweep/csdl_reg<256,4>", [&](const void* in, void* out) {
            CarrySpec cs;
            return launch_csdl_reg<T, 256, 4, INCL, 1>(env.s, in, out, env.n, cs, 0);
        });
        sv.emplace_back("sweep/csdl_reg<128,4>", [&](const void* in, void* out) {
            CarrySpec cs;
            return launch_csdl_reg<T, 128, 4, INCL, 1>(env.s, in, out, env.n, cs, 0);
        });
        static bool cfg_a = false, cfg_b = false, cfg_c = false;
        sv.emplace_back("sweep/csdl_smem<128,8>", [&](const void* in, void* out) {
            CarrySpec cs;
            return launch_csdl_smem<T, 128, 8, INCL, 1>(env.s, in, out, env.n, cs, cfg_a, 0);
        });
        sv.emplace_back("sweep/csdl_smem<128,4>", [&](const void* in, void* out) {
            CarrySpec cs;
            return launch_csdl_smem<T, 128, 4, INCL, 1>(env.s, in, out, env.n, cs, cfg_b, 0);
        });
        sv.emplace_back("sweep/csdl_smem<256,4>", [&](const void* in, void* out) {
            CarrySpec cs;
            return launch_csdl_smem<T, 256, 4, INCL, 1>(env.s, in, out, env.n, cs, cfg_c, 0);
        });
        static bool cfg_d = false, cfg_e = false, cfg_f = false, cfg_g = false;
        sv.emplace_back("sweep/csdl_blocked<256,9>", [&](const void* in, void* out) {
            CarrySpec cs;
            return launch_csdl_blocked<T, 256, 9, INCL, 1>(env.s, in, out, env.n, cs, cfg_d, 0);
        });
        sv.emplace_back("sweep/csdl_blocked<128,9>", [&](const void* in, void* out) {
            CarrySpec cs;
            return launch_csdl_blocked<T, 128, 9, INCL, 1>(env.s, in, out, env.n, cs, cfg_e, 0);
        });
        sv.emplace_back("sweep/csdl_blocked<128,5>", [&](const void* in, void* out) {
            CarrySpec cs;
            return launch_csdl_blocked<T, 128, 5, INCL, 1>(env.s, in, out, env.n, cs, cfg_f, 0);
        });
        sv.emplace_back("sweep/csdl_blocked<256,5>", [&](const void* in, void* out) {
            CarrySpec cs;
            return launch_csdl_blocked<T, 256, 5, INCL, 1>(env.s, in, out, env.n, cs, cfg_g, 0);
        });
        // mid-size ladder (round 2): more tile shapes between 8 and 28 KB
        static bool cfg_h = false, cfg_i = false, cfg_j = false, cfg_k = false, cfg_l = false, cfg_m = false, cfg_n = false;
        sv.emplace_back("sweep/csdl_smem<128,6>", [&](const void* in, void* out) {
            CarrySpec cs;
            return launch_csdl_smem<T, 128, 6, INCL, 1>(env.s, in, out, env.n, cs, cfg_h, 0);
        });
        sv.emplace_back("sweep/csdl_smem<128,12>", [&](const void* in, void* out) {
            CarrySpec cs;
            return launch_csdl_smem<T, 128, 12, INCL, 1>(env.s, in, out, env.n, cs, cfg_i, 0);
        });
        sv.emplace_back("sweep/csdl_smem<64,8>", [&](const void* in, void* out) {
            CarrySpec cs;
            return launch_csdl_smem<T, 64, 8, INCL, 1>(env.s, in, out, env.n, cs, cfg_j, 0);
        });
        sv.emplace_back("sweep/csdl_blocked<128,7>", [&](const void* in, void* out) {
            CarrySpec cs;
            return launch_csdl_blocked<T, 128, 7, INCL, 1>(env.s, in, out, env.n, cs, cfg_k, 0);
        });
        sv.emplace_back("sweep/csdl_blocked<256,7>", [&](const void* in, void* out) {
            CarrySpec cs;
            return launch_csdl_blocked<T, 256, 7, INCL, 1>(env.s, in, out, env.n, cs, cfg_l, 0);
        });
        sv.emplace_back("sweep/csdl_blocked<256,3>", [&](const void* in, void* out) {
            CarrySpec cs;
            return launch_csdl_blocked<T, 256, 3, INCL, 1>(env.s, in, out, env.n, cs, cfg_m, 0);
        });
        sv.emplace_back("sweep/csdl_blocked<128,11>", [&](const void* in, void* out) {
            CarrySpec cs;
            return launch_csdl_blocked<T, 128, 11, INCL, 1>(env.s, in, out, env.n, cs, cfg_n, 0);
        });
        sv.emplace_back("sweep/csdl_tma<512,S2>", [&](const void* in, void* out) {
            CarrySpec cs;
            return launch_csdl_tma<T, 512, 2, INCL, 1>(env.s, in, out, env.n, cs, occ_t, 0, 0);
        });
        sv.emplace_back("sweep/rts<128,12>", [&](const void* in, void* out) {
            CarrySpec cs;
            return launch_rts<T, 128, 12, INCL>(env.s, in, out, env.n, cs, 0);
        });
        int bad = 0;
        for (auto& kv : sv) {
            Result r = run_variant(env, kv.first.c_str(), true, kv.second);
            if (!r.ok || r.errs) ++bad;
        }
        return bad;
    }
    std::vector<std::pair<std::string, LaunchFn>> variants;
    build_variants<T, INCL>(env, variants);
    int bad = 0;
    for (auto& kv : variants) {
        if (!wanted(kv.first)) continue;
        Result r = run_variant(env, kv.first.c_str(), true, kv.second);
        if (!r.ok || r.errs) ++bad;
    }
    // knobs on the two main candidates
    if (wanted("knob")) {
        for (uint32_t sleep_ns : {20u, 100u}) {
            env.s.poll_sleep_ns = sleep_ns;
            char name[96];
            snprintf(name, sizeof(name), "knob/csdl_reg<256,4>/sleep%u", sleep_ns);
            run_variant(env, name, true, [&](const void* in, void* out) {
                CarrySpec cs;
                return launch_csdl_reg<T, 256, 4, INCL>(env.s, in, out, env.n, cs, 0);
            });
            snprintf(name, sizeof(name), "knob/csdl_tma<512,S2>/sleep%u", sleep_ns);
            static int occ = 0;
            run_variant(env, name, true, [&](const void* in, void* out) {
                CarrySpec cs;
                return launch_csdl_tma<T, 512, 2, INCL>(env.s, in, out, env.n, cs, occ, 0, 0);
            });
        }
        env.s.poll_sleep_ns = 0;
        struct { const char* name; uint32_t flags; } protos[] = {
            {"st_release+ld_acquire", PF_PUBLISH_ST_RELEASE | PF_POLL_ACQUIRE},
            {"st_relaxed+ld_relaxed", PF_PUBLISH_ST_RELAXED},
            {"atom_exch+ld_acquire", PF_POLL_ACQUIRE},
            {"atom_exch+ld_volatile", PF_POLL_VOLATILE},
            {"atom_exch+ld_relaxed", 0},
        };
        static int occ2 = 0, occ3 = 0;
        for (auto& pr : protos) {
            env.s.flags = pr.flags;
            char name[96];
            snprintf(name, sizeof(name), "knob/csdl_reg<256,4>/%s", pr.name);
            run_variant(env, name, true, [&](const void* in, void* out) {
                CarrySpec cs;
                return launch_csdl_reg<T, 256, 4, INCL>(env.s, in, out, env.n, cs, 0);
            });
            snprintf(name, sizeof(name), "knob/csdl_tma<512,S2>/%s", pr.name);
            run_variant(env, name, true, [&](const void* in, void* out) {
                CarrySpec cs;
                return launch_csdl_tma<T, 512, 2, INCL>(env.s, in, out, env.n, cs, occ2, 0, 0);
            });
            snprintf(name, sizeof(name), "knob/csdl_tma<256,S2>/%s", pr.name);
            run_variant(env, name, true, [&](const void* in, void* out) {
                CarrySpec cs;
                return launch_csdl_tma<T, 256, 2, INCL>(env.s, in, out, env.n, cs, occ3, 0, 0);
            });
        }
        env.s.flags = 0;
    }
    // experiments on the register path (some produce WRONG results on purpose: not verified)
    if (wanted("exp")) {
        struct { const char* name; uint32_t flags; bool verify; } exps[] = {
            {"exp/reg<256,4>/lookback_wait_all", PF_LOOKBACK_WAIT_ALL, true},
            {"exp/reg<256,4>/static_tile", PF_STATIC_TILE, true},
            {"exp/reg<256,4>/skip_lookback", PF_SKIP_LOOKBACK, false},
            {"exp/reg<256,4>/static_tile+skip_lookback", PF_STATIC_TILE | PF_SKIP_LOOKBACK, false},
            {"exp/reg<256,4>/fence_all", PF_FENCE_ALL, true},
            {"exp/reg<256,4>/ld_allocate", PF_LD_ALLOCATE, true},
        };
        const uint32_t pdl_base = env.s.flags & PF_PDL;  // experiments keep the launch mode of the run
        for (auto& ex : exps) {
            env.s.flags = ex.flags;
            run_variant(env, ex.name, ex.verify, [&](const void* in, void* out) {
                CarrySpec cs;
                return launch_csdl_reg<T, 256, 4, INCL, 1>(env.s, in, out, env.n, cs, 0);
            });
        }
        for (uint32_t fl : {(uint32_t)PF_STATIC_TILE, (uint32_t)PF_SKIP_LOOKBACK, (uint32_t)PF_LD_ALLOCATE, (uint32_t)PF_LOOKBACK_WAIT_ALL}) {
            env.s.flags = fl;
            char name[96];
            snprintf(name, sizeof(name), "exp/reg<64,24>/w1/flags%u", fl);
            run_variant(env, name, fl != PF_SKIP_LOOKBACK, [&](const void* in, void* out) {
                CarrySpec cs;
                return launch_csdl_reg<T, 64, 24, INCL, 1>(env.s, in, out, env.n, cs, 0);
            });
        }
        for (uint32_t fl : {0u, (uint32_t)PF_STATIC_TILE, (uint32_t)PF_SKIP_LOOKBACK, (uint32_t)(PF_SKIP_LOOKBACK | PF_STATIC_TILE), (uint32_t)PF_LOOKBACK_SINGLE, (uint32_t)PF_LOOKBACK_WAIT_ALL}) {
            env.s.flags = fl | pdl_base;
            char name[96];
            static bool cfgx = false;
            snprintf(name, sizeof(name), "exp/smem<128,8>/flags%u", fl);
            run_variant(env, name, !(fl & PF_SKIP_LOOKBACK), [&](const void* in, void* out) {
                CarrySpec cs;
                return launch_csdl_smem<T, 128, 8, INCL, 1>(env.s, in, out, env.n, cs, cfgx, 0);
            });
        }
        // eager look-back (warp 0 walks while the tile is in flight), with poll back-off and the other mid-size knobs
        for (uint32_t sleep_ns : {0u, 100u, 300u}) {
            for (uint32_t fl : {(uint32_t)PF_EAGER_LOOKBACK, (uint32_t)(PF_EAGER_LOOKBACK | PF_L2_IN_EVICT_FIRST),
                                (uint32_t)(PF_EAGER_LOOKBACK | PF_STATIC_TILE), (uint32_t)PF_L2_IN_EVICT_FIRST}) {
                if (sleep_ns && !(fl & PF_EAGER_LOOKBACK)) continue;
                env.s.flags = fl | pdl_base;
                env.s.poll_sleep_ns = sleep_ns;
                char name[96];
                static bool cfge = false;
                snprintf(name, sizeof(name), "exp/smem<128,8>/eager%u/sleep%u", fl, sleep_ns);
                run_variant(env, name, true, [&](const void* in, void* out) {
                    CarrySpec cs;
                    return launch_csdl_smem<T, 128, 8, INCL, 1>(env.s, in, out, env.n, cs, cfge, 0);
                });
            }
        }
        env.s.poll_sleep_ns = 0;
        // persistent one-wave kernel (CTA c scans tiles c, c + grid, ...): no ticket, no CTA launch per tile
        for (uint32_t fl : {0u, (uint32_t)PF_L2_IN_EVICT_FIRST}) {
            env.s.flags = fl | pdl_base;
            char name[96];
            static int res_p = 0;
            snprintf(name, sizeof(name), "exp/smem_persistent<128,8>/flags%u", fl);
            run_variant(env, name, true, [&](const void* in, void* out) {
                CarrySpec cs;
                return launch_csdl_smem_persistent<T, 128, 8, INCL, 1>(env.s, in, out, env.n, cs, res_p, 0);
            });
        }
        // the blocked tile (default from 2^27 keys): ticket vs blockIdx order, plain and Decoupled Fallback
        for (uint32_t fl : {0u, (uint32_t)PF_STATIC_TILE}) {
            env.s.flags = fl | pdl_base;
            char name[96];
            static bool cfgb = false, cfgbf = false;
            snprintf(name, sizeof(name), "exp/blocked<256,9>/flags%u", fl);
            run_variant(env, name, true, [&](const void* in, void* out) {
                CarrySpec cs;
                return launch_csdl_blocked<T, 256, 9, INCL, 1>(env.s, in, out, env.n, cs, cfgb, 0);
            });
            const FallbackParams fbp{kNoStall, 64u, nullptr};
            snprintf(name, sizeof(name), "exp/blocked_fb<256,9>/flags%u/spin64", fl);
            run_variant(env, name, true, [&](const void* in, void* out) {
                CarrySpec cs;
                return launch_csdl_blocked_fallback<T, 256, 9, INCL>(env.s, in, out, env.n, cs, fbp, cfgbf, 0);
            });
        }
        // Decoupled Fallback without tickets (tile = blockIdx.x is deadlock-free there), spin budgets, no stalls
        static uint32_t* fb_stats = nullptr;  // {spins, fallbacks, insertions, windows} summed over the variant's launches
        if (!fb_stats) CK(cudaMalloc(&fb_stats, 4 * sizeof(uint32_t)));
        for (uint32_t spin : {16u, 64u, 256u, 1024u}) {
            for (uint32_t fl : {0u, (uint32_t)PF_STATIC_TILE, (uint32_t)(PF_STATIC_TILE | PF_L2_IN_EVICT_FIRST)}) {
                env.s.flags = fl | pdl_base;
                char name[96];
                static bool cfgf = false;
                snprintf(name, sizeof(name), "exp/smem_fb<128,8>/flags%u/spin%u", fl, spin);
                CK(cudaMemset(fb_stats, 0, 4 * sizeof(uint32_t)));
                run_variant(env, name, true, [&](const void* in, void* out) {
                    CarrySpec cs;
                    FallbackParams f{kNoStall, spin, fb_stats};
                    return launch_csdl_smem_fallback<T, 128, 8, INCL>(env.s, in, out, env.n, cs, f, cfgf, 0);
                });
                uint32_t h[4];
                CK(cudaMemcpy(h, fb_stats, sizeof(h), cudaMemcpyDeviceToHost));
                printf("{\"variant\": \"%s/stats\", \"spins\": %u, \"fallbacks\": %u, \"insertions\": %u, \"windows\": %u}\n", name,
                       h[0], h[1], h[2], h[3]);
            }
        }
        // the same on the register tiles of the small sizes (<128,4> up to 2^17, <256,4> up to 2^22), stats off
        for (uint32_t fl : {0u, (uint32_t)PF_STATIC_TILE}) {
            env.s.flags = fl | pdl_base;
            char name[96];
            snprintf(name, sizeof(name), "exp/reg<256,4>/w1/flags%u", fl);
            run_variant(env, name, true, [&](const void* in, void* out) {
                CarrySpec cs;
                return launch_csdl_reg<T, 256, 4, INCL, 1>(env.s, in, out, env.n, cs, 0);
            });
            snprintf(name, sizeof(name), "exp/reg<128,4>/w1/flags%u", fl);
            run_variant(env, name, true, [&](const void* in, void* out) {
                CarrySpec cs;
                return launch_csdl_reg<T, 128, 4, INCL, 1>(env.s, in, out, env.n, cs, 0);
            });
            snprintf(name, sizeof(name), "exp/reg<128,8>/w1/flags%u", fl);  // 56 registers -> 9 CTAs of 128 threads per SM: one wave = 1332 tiles of 4096 keys
            run_variant(env, name, true, [&](const void* in, void* out) {
                CarrySpec cs;
                return launch_csdl_reg<T, 128, 8, INCL, 1>(env.s, in, out, env.n, cs, 0);
            });
            snprintf(name, sizeof(name), "exp/reg<128,6>/w1/flags%u", fl);  // 48 registers -> 10 CTAs per SM: one wave = 1480 tiles of 3072 keys
            run_variant(env, name, true, [&](const void* in, void* out) {
                CarrySpec cs;
                return launch_csdl_reg<T, 128, 6, INCL, 1>(env.s, in, out, env.n, cs, 0);
            });
            snprintf(name, sizeof(name), "exp/reg<256,4>r36/w1/flags%u", fl);  // 7 CTAs per SM: 2^22 keys = 1024 tiles fit one wave of 1036
            run_variant(env, name, true, [&](const void* in, void* out) {
                CarrySpec cs;
                return launch_csdl_reg<T, 256, 4, INCL, 1, false, 36>(env.s, in, out, env.n, cs, 0);
            });
            snprintf(name, sizeof(name), "exp/reg<256,4>r32/w1/flags%u", fl);  // 8 CTAs per SM
            run_variant(env, name, true, [&](const void* in, void* out) {
                CarrySpec cs;
                return launch_csdl_reg<T, 256, 4, INCL, 1, false, 32>(env.s, in, out, env.n, cs, 0);
            });
            const FallbackParams f{kNoStall, 64u, nullptr};
            snprintf(name, sizeof(name), "exp/reg_fb<256,4>/flags%u/spin64", fl);
            run_variant(env, name, true, [&](const void* in, void* out) {
                CarrySpec cs;
                return launch_csdl_fallback<T, 256, 4, INCL>(env.s, in, out, env.n, cs, f, 0);
            });
            snprintf(name, sizeof(name), "exp/reg_fb<128,4>/flags%u/spin64", fl);
            run_variant(env, name, true, [&](const void* in, void* out) {
                CarrySpec cs;
                return launch_csdl_fallback<T, 128, 4, INCL>(env.s, in, out, env.n, cs, f, 0);
            });
            static bool cfgs = false;
            snprintf(name, sizeof(name), "exp/smem_fb<128,8>/nostats/flags%u/spin64", fl);
            run_variant(env, name, true, [&](const void* in, void* out) {
                CarrySpec cs;
                return launch_csdl_smem_fallback<T, 128, 8, INCL>(env.s, in, out, env.n, cs, f, cfgs, 0);
            });
        }
        env.s.flags = PF_LOOKBACK_WAIT_ALL;
        {
            static int occw = 0;
            run_variant(env, "exp/tma<512,S2>/lookback_wait_all", true, [&](const void* in, void* out) {
                CarrySpec cs;
                return launch_csdl_tma<T, 512, 2, INCL>(env.s, in, out, env.n, cs, occw, 0, 0);
            });
        }
        env.s.flags = 0;
        for (uint32_t ns : {500u, 1000u, 2000u}) {
            env.s.flags = PF_DELAY_FIRST;
            env.s.poll_sleep_ns = ns;
            char name[96];
            snprintf(name, sizeof(name), "exp/reg<256,4>/delay_first%u", ns);
            run_variant(env, name, true, [&](const void* in, void* out) {
                CarrySpec cs;
                return launch_csdl_reg<T, 256, 4, INCL>(env.s, in, out, env.n, cs, 0);
            });
        }
        env.s.flags = pdl_base;
        env.s.poll_sleep_ns = 0;
    }
    return bad;
}

int main(int argc, char** argv)
{
    Env env;
    const char* filter = nullptr;
    bool sweep = false, pdl = false, flush = false;
    int max_log2n = 30, min_log2n = 10;
    for (int i = 1; i < argc; ++i) {
        if (!strcmp(argv[i], "--log2n") && i + 1 < argc) env.n = 1ull << atoi(argv[++i]);
        else if (!strcmp(argv[i], "--n") && i + 1 < argc) env.n = strtoull(argv[++i], nullptr, 10);
        else if (!strcmp(argv[i], "--iters") && i + 1 < argc) env.iters = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--filter") && i + 1 < argc) filter = argv[++i];
        else if (!strcmp(argv[i], "--float")) env.is_float = true;
        else if (!strcmp(argv[i], "--prof")) env.prof_enabled = true;
        else if (!strcmp(argv[i], "--sweep")) sweep = true;
        else if (!strcmp(argv[i], "--max-log2n") && i + 1 < argc) max_log2n = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--exclusive")) env.exclusive = true;
        else if (!strcmp(argv[i], "--pdl")) pdl = true;
        else if (!strcmp(argv[i], "--flush")) flush = true;
        else if (!strcmp(argv[i], "--min-log2n") && i + 1 < argc) min_log2n = atoi(argv[++i]);
        else {
            fprintf(stderr, "unknown argument %s\n", argv[i]);
            return 2;
        }
    }
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    fprintf(stderr, "device: %s, %d SMs, L2 %d MB\n", prop.name, prop.multiProcessorCount, prop.l2CacheSize >> 20);
    const size_t bytes = env.n * 4 + 64;
    CK(cudaMalloc(&env.in, bytes));
    CK(cudaMalloc(&env.out, bytes));
    CK(cudaMalloc(&env.ref, bytes));
    CK(cudaMalloc(&env.work, bytes));
    env.s.sm_count = prop.multiProcessorCount;
    if (pdl) env.s.flags |= PF_PDL;
    env.s.desc_tiles = (uint32_t)div_up(env.n, 2048) + 1;
    CK(cudaMalloc(&env.s.desc, (size_t)env.s.desc_tiles * 8));
    CK(cudaMemset(env.s.desc, 0, (size_t)env.s.desc_tiles * 8));
    CK(cudaMalloc(&env.s.counter, 64));
    CK(cudaMemset(env.s.counter, 0, 64));
    env.s.status = env.s.counter + 1;
    env.tma_copy_counter = env.s.counter + 8;
    env.s.partials_capacity = env.s.desc_tiles;
    CK(cudaMalloc(&env.s.partials, (size_t)env.s.partials_capacity * 4));
    CK(cudaMalloc(&env.prof_buf, sizeof(unsigned long long) * kProfSlots * kProfFields));
    env.sm_ghz = prop.clockRate / 1e6;
    CK(cudaEventCreate(&env.e0));
    CK(cudaEventCreate(&env.e1));
    if (flush) {
        env.flush_bytes = 512ull << 20;  // 4 x the 126 MB L2
        CK(cudaMalloc(&env.flush_buf, env.flush_bytes));
    }
    if (env.is_float)
        init_random_f32_kernel<<<2048, 256>>>((float*)env.in, env.n, 1234, 0);
    else
        init_random_u32_kernel<<<2048, 256>>>((uint32_t*)env.in, env.n, 1234, 0xFFFFFFFFu, 0);
    CK(cudaDeviceSynchronize());

    if (sweep) {
        // size sweep (BASELINE configs[3]): 2^10 .. 2^max keys, default kernels + CUB; buffers
        // were allocated for env.n = 2^max.  n <= 2^23: in + out fit the 126 MB L2 ("L2-resident")
        const uint64_t n_max = env.n;
        int bad_total = 0;
        (void)max_log2n;
        for (int lg = min_log2n; (1ull << lg) <= n_max; ++lg) {
            env.n = 1ull << lg;
            env.iters = lg <= 20 ? 200 : lg <= 24 ? 100 : 30;
            bad_total += env.is_float ? run_all<float, true>(env, filter ? filter : "sweep")
                                      : run_all<uint32_t, true>(env, filter ? filter : "sweep");
        }
        return bad_total ? 1 : 0;
    }
    int bad;
    if (env.is_float)
        bad = env.exclusive ? run_all<float, false>(env, filter) : run_all<float, true>(env, filter);
    else
        bad = env.exclusive ? run_all<uint32_t, false>(env, filter) : run_all<uint32_t, true>(env, filter);
    fprintf(stderr, "variants with errors: %d\n", bad);
    return bad ? 1 : 0;
}
