// pcie_probe.cu -- memcpy-only probe of the host <-> device path on a multi-GPU box.
//
// Answers one question for the end-to-end (host-buffer) numbers of bench.py: how many GB/s can
// N GPUs of this box pull from / push to pinned host memory AT THE SAME TIME, and does that
// depend on how the host buffers are allocated or on who drives the copies?  No kernels of the
// library are involved; nothing here is on the product path.
//
//   gps_pcie_probe --gpus 0,1,2,3 [--dir h2d|d2h|both] [--alloc pinned|wc|huge|pageable]
//                  [--drive threads|procs] [--mib 256] [--secs 0.6] [--bind 0|1]
//
//   --dirs h2d,h2d,d2h,d2h   per-GPU direction (ingress through some GPUs, egress through others)
//   --alloc  pinned   cudaHostAlloc (what torch.pin_memory and gps_malloc_host use)
//            wc       cudaHostAllocWriteCombined for the H2D source
//            huge     2 MiB-aligned anonymous memory + MADV_HUGEPAGE + cudaHostRegister
//            pageable plain malloc (the driver stages through its own pinned buffers)
//   --drive  threads  one process, one host thread per GPU
//            procs    one forked process per GPU (what torchrun does)
//   --bind   1        pin each worker to the CPUs NVML reports as local to its GPU before the
//                     host buffers are allocated and first touched
// Prints one JSON line: per-GPU and aggregate GB/s per direction over a common wall-clock window
// (all workers start together behind a process-shared barrier).
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <pthread.h>
#include <sched.h>
#include <sys/mman.h>
#include <sys/wait.h>
#include <unistd.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

namespace {

struct Shared {
    pthread_barrier_t bar;
    double h2d_gbs[16];
    double d2h_gbs[16];
    int failed[16];
};

struct Opts {
    std::vector<int> gpus;
    std::string dir = "both", alloc = "pinned", drive = "threads";
    std::vector<std::string> dirs;  // optional per-GPU direction (h2d | d2h | both), overrides dir
    size_t mib = 256;
    double secs = 0.6;
    int bind = 0;
};

#define CK(call)                                                                              \
    do {                                                                                      \
        cudaError_t e_ = (call);                                                              \
        if (e_ != cudaSuccess) {                                                              \
            fprintf(stderr, "%s: %s (line %d)\n", #call, cudaGetErrorString(e_), __LINE__);   \
            return false;                                                                     \
        }                                                                                     \
    } while (0)

// CPUs local to GPU `index` according to NVML (dlopen: no link-time dependency)
bool bind_to_gpu_cpus(int index)
{
    void* h = dlopen("libnvidia-ml.so.1", RTLD_NOW);
    if (!h) return false;
    using init_t = int (*)();
    using handle_t = int (*)(unsigned, void**);
    using aff_t = int (*)(void*, unsigned, unsigned long*);
    auto init = (init_t)dlsym(h, "nvmlInit_v2");
    auto get = (handle_t)dlsym(h, "nvmlDeviceGetHandleByIndex_v2");
    auto aff = (aff_t)dlsym(h, "nvmlDeviceGetCpuAffinity");
    if (!init || !get || !aff || init() != 0) return false;
    void* dev = nullptr;
    if (get((unsigned)index, &dev) != 0) return false;
    unsigned long words[16] = {};
    if (aff(dev, 16, words) != 0) return false;
    cpu_set_t set;
    CPU_ZERO(&set);
    int n = 0;
    for (int i = 0; i < 1024; ++i)
        if ((words[i / 64] >> (i % 64)) & 1ul) {
            CPU_SET(i, &set);
            ++n;
        }
    return n > 0 && sched_setaffinity(0, sizeof(set), &set) == 0;
}

bool alloc_host(const Opts& o, size_t bytes, bool is_src, void** out, void** to_free, int* kind)
{
    *kind = 0;
    if (o.alloc == "pinned" || (o.alloc == "wc" && !is_src)) {
        CK(cudaHostAlloc(out, bytes, cudaHostAllocDefault));
        *kind = 1;
    } else if (o.alloc == "wc") {
        CK(cudaHostAlloc(out, bytes, cudaHostAllocWriteCombined));
        *kind = 1;
    } else if (o.alloc == "huge") {
        void* p = mmap(nullptr, bytes + (2u << 20), PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
        if (p == MAP_FAILED) return false;
        *to_free = p;
        void* al = (void*)(((uintptr_t)p + (2u << 20) - 1) & ~(uintptr_t)((2u << 20) - 1));
        madvise(al, bytes, MADV_HUGEPAGE);
        memset(al, 1, bytes);
        CK(cudaHostRegister(al, bytes, cudaHostRegisterDefault));
        *out = al;
        *kind = 2;
        return true;
    } else {  // pageable
        *out = malloc(bytes);
        if (!*out) return false;
        *kind = 3;
    }
    memset(*out, 1, bytes);  // first touch on the worker's CPUs
    return true;
}

bool worker(const Opts& o, int slot, Shared* sh, int* phase)
{
    const int gpu = o.gpus[slot];
    if (o.bind) bind_to_gpu_cpus(gpu);
    CK(cudaSetDevice(gpu));
    const size_t bytes = o.mib << 20;
    const std::string& mydir = slot < (int)o.dirs.size() ? o.dirs[slot] : o.dir;
    const bool up = mydir != "d2h", down = mydir != "h2d";
    void *h_src = nullptr, *h_dst = nullptr, *f1 = nullptr, *f2 = nullptr, *d_a = nullptr, *d_b = nullptr;
    int k1 = 0, k2 = 0;
    if (up && !alloc_host(o, bytes, true, &h_src, &f1, &k1)) return false;
    if (down && !alloc_host(o, bytes, false, &h_dst, &f2, &k2)) return false;
    CK(cudaMalloc(&d_a, bytes));
    CK(cudaMalloc(&d_b, bytes));
    CK(cudaMemset(d_b, 2, bytes));
    cudaStream_t s_up, s_down;
    CK(cudaStreamCreateWithFlags(&s_up, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&s_down, cudaStreamNonBlocking));
    // warm-up
    if (up) CK(cudaMemcpyAsync(d_a, h_src, bytes, cudaMemcpyHostToDevice, s_up));
    if (down) CK(cudaMemcpyAsync(h_dst, d_b, bytes, cudaMemcpyDeviceToHost, s_down));
    CK(cudaDeviceSynchronize());
    cudaEvent_t u0, u1, d0, d1;
    CK(cudaEventCreate(&u0));
    CK(cudaEventCreate(&u1));
    CK(cudaEventCreate(&d0));
    CK(cudaEventCreate(&d1));

    pthread_barrier_wait(&sh->bar);  // common start
    ++*phase;
    const auto t0 = std::chrono::steady_clock::now();
    size_t n_up = 0, n_down = 0;
    CK(cudaEventRecord(u0, s_up));
    CK(cudaEventRecord(d0, s_down));
    // keep two copies queued per direction so the engines never idle
    double el = 0;
    do {
        if (up) {
            CK(cudaMemcpyAsync(d_a, h_src, bytes, cudaMemcpyHostToDevice, s_up));
            ++n_up;
        }
        if (down) {
            CK(cudaMemcpyAsync(h_dst, d_b, bytes, cudaMemcpyDeviceToHost, s_down));
            ++n_down;
        }
        if (n_up + n_down >= 4) {
            if (up) CK(cudaStreamSynchronize(s_up));
            if (down) CK(cudaStreamSynchronize(s_down));
        }
        el = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    } while (el < o.secs);
    CK(cudaEventRecord(u1, s_up));
    CK(cudaEventRecord(d1, s_down));
    CK(cudaDeviceSynchronize());
    float ms_u = 0, ms_d = 0;
    CK(cudaEventElapsedTime(&ms_u, u0, u1));
    CK(cudaEventElapsedTime(&ms_d, d0, d1));
    sh->h2d_gbs[slot] = up && ms_u > 0 ? (double)n_up * bytes / (ms_u * 1e-3) / 1e9 : 0.0;
    sh->d2h_gbs[slot] = down && ms_d > 0 ? (double)n_down * bytes / (ms_d * 1e-3) / 1e9 : 0.0;
    pthread_barrier_wait(&sh->bar);  // nobody tears down while others still measure
    ++*phase;
    if (k1 == 2) cudaHostUnregister(h_src);
    if (k2 == 2) cudaHostUnregister(h_dst);
    if (k1 == 1) cudaFreeHost(h_src);
    if (k2 == 1) cudaFreeHost(h_dst);
    if (k1 == 3) free(h_src);
    if (k2 == 3) free(h_dst);
    (void)f1;
    (void)f2;
    cudaFree(d_a);
    cudaFree(d_b);
    return true;
}

// a worker that fails still keeps its appointments at the barriers
bool run_worker(const Opts& o, int slot, Shared* sh)
{
    int phase = 0;
    const bool ok = worker(o, slot, sh, &phase);
    if (!ok) sh->failed[slot] = 1;
    for (; phase < 2; ++phase) pthread_barrier_wait(&sh->bar);
    return ok;
}

}  // namespace

int main(int argc, char** argv)
{
    Opts o;
    for (int i = 1; i < argc; ++i) {
        const std::string a = argv[i];
        auto val = [&]() -> std::string { return i + 1 < argc ? argv[++i] : ""; };
        if (a == "--gpus") {
            std::string v = val();
            size_t p = 0;
            while (p < v.size()) {
                size_t q = v.find(',', p);
                if (q == std::string::npos) q = v.size();
                o.gpus.push_back(atoi(v.substr(p, q - p).c_str()));
                p = q + 1;
            }
        } else if (a == "--dirs") {  // per GPU, e.g. h2d,h2d,d2h,d2h
            std::string v = val();
            size_t p = 0;
            while (p < v.size()) {
                size_t q = v.find(',', p);
                if (q == std::string::npos) q = v.size();
                o.dirs.push_back(v.substr(p, q - p));
                p = q + 1;
            }
            o.dir = "mixed";
        } else if (a == "--dir") o.dir = val();
        else if (a == "--alloc") o.alloc = val();
        else if (a == "--drive") o.drive = val();
        else if (a == "--mib") o.mib = (size_t)atol(val().c_str());
        else if (a == "--secs") o.secs = atof(val().c_str());
        else if (a == "--bind") o.bind = atoi(val().c_str());
        else {
            fprintf(stderr, "unknown argument %s\n", a.c_str());
            return 2;
        }
    }
    if (o.gpus.empty()) o.gpus.push_back(0);
    const int n = (int)o.gpus.size();
    if (n > 16) return 2;
    Shared* sh = (Shared*)mmap(nullptr, sizeof(Shared), PROT_READ | PROT_WRITE, MAP_SHARED | MAP_ANONYMOUS, -1, 0);
    if (sh == MAP_FAILED) return 1;
    memset(sh, 0, sizeof(Shared));
    pthread_barrierattr_t at;
    pthread_barrierattr_init(&at);
    pthread_barrierattr_setpshared(&at, PTHREAD_PROCESS_SHARED);
    pthread_barrier_init(&sh->bar, &at, (unsigned)n);

    bool ok = true;
    if (o.drive == "procs") {  // fork BEFORE any CUDA call
        std::vector<pid_t> kids;
        for (int s = 0; s < n; ++s) {
            pid_t pid = fork();
            if (pid == 0) {
                const bool w = run_worker(o, s, sh);
                _exit(w ? 0 : 1);
            }
            kids.push_back(pid);
        }
        for (pid_t k : kids) {
            int st = 0;
            waitpid(k, &st, 0);
            ok = ok && WIFEXITED(st) && WEXITSTATUS(st) == 0;
        }
    } else {
        std::vector<std::thread> th;
        std::vector<int> res(n, 0);
        for (int s = 0; s < n; ++s) th.emplace_back([&, s] { res[s] = run_worker(o, s, sh) ? 1 : 0; });
        for (auto& t : th) t.join();
        for (int s = 0; s < n; ++s) ok = ok && res[s];
    }
    double su = 0, sd = 0;
    printf("{\"gpus\": [");
    for (int s = 0; s < n; ++s) printf("%s%d", s ? ", " : "", o.gpus[s]);
    printf("], \"dir\": \"%s\", \"alloc\": \"%s\", \"drive\": \"%s\", \"bind\": %d, \"mib\": %zu, \"ok\": %s, \"h2d_gbs\": [",
           o.dir.c_str(), o.alloc.c_str(), o.drive.c_str(), o.bind, o.mib, ok ? "true" : "false");
    for (int s = 0; s < n; ++s) {
        printf("%s%.1f", s ? ", " : "", sh->h2d_gbs[s]);
        su += sh->h2d_gbs[s];
    }
    printf("], \"d2h_gbs\": [");
    for (int s = 0; s < n; ++s) {
        printf("%s%.1f", s ? ", " : "", sh->d2h_gbs[s]);
        sd += sh->d2h_gbs[s];
    }
    printf("], \"h2d_total\": %.1f, \"d2h_total\": %.1f, \"both_total\": %.1f}\n", su, sd, su + sd);
    return ok ? 0 : 1;
}
