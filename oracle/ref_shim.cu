/*
 * oracle/ref_shim.cu -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Thin C-ABI shim around the UNMODIFIED reference CUDA kernels.  It #includes the
 * reference headers where they lie under /root/reference/GPUPrefixSumsCUDA (include
 * path given by oracle/Makefile; no reference source is copied into this repo) and
 * launches them on caller-owned device buffers so tests can compare the new kernels
 * with the reference's own CUDA scans on a B200.  The launch arithmetic restates the
 * private Dispatch* methods of the reference dispatchers:
 *   ChainedScanDecoupledLookbackDispatcher.cuh:166-195 (align16, ClearMemory, <<<blocks,256>>>)
 *   ReduceThenScanDispatcher.cuh:158-187              (Reduce, SpineScan<32,4><<<1,1024>>>, Propagate)
 *   EmulatedDeadlockingDispatcher.cuh:152-166
 *   CubDispatcher.cuh:42-58                            (cub::DeviceScan::InclusiveSum in place)
 * Built into oracle/_ref/libref_shim.so (git-ignored, travels to the GPU box).
 */
#include <cstdint>
#include <cstdio>

#include "ChainedScanDecoupledLookback.cuh"
#include "EmulatedDeadlocking.cuh"
#include "ReduceThenScan.cuh"
#include "UtilityKernels.cuh"
#include "cub/device/device_scan.cuh"

namespace {
constexpr uint32_t kWarps = 8;         // CSDL_WARPS / RTS_WARPS      (…Dispatcher.cuh:13)
constexpr uint32_t kUint4PerThread = 4; // UINT4_PER_THREAD            (…Dispatcher.cuh:14)
constexpr uint32_t kThreads = kWarps * 32;
constexpr uint32_t kPartitionSize = kThreads * kUint4PerThread * 4;  // 4096 keys

inline uint32_t div_round_up(uint32_t x, uint32_t y) { return (x + y - 1) / y; }
inline uint32_t align4(uint32_t x) { return div_round_up(x, 4) * 4; }

uint32_t* g_index = nullptr;
uint32_t* g_tbr = nullptr;
uint32_t g_tbr_blocks = 0;
void* g_cub_temp = nullptr;
size_t g_cub_temp_bytes = 0;

int ensure_scratch(uint32_t blocks)
{
    if (!g_index && cudaMalloc(&g_index, sizeof(uint32_t)) != cudaSuccess) return -1;
    if (blocks > g_tbr_blocks) {
        if (g_tbr) cudaFree(g_tbr);
        if (cudaMalloc(&g_tbr, (size_t)blocks * sizeof(uint32_t)) != cudaSuccess) return -1;
        g_tbr_blocks = blocks;
    }
    return 0;
}

int ensure_cub_temp(size_t bytes)
{
    if (bytes > g_cub_temp_bytes) {
        if (g_cub_temp) cudaFree(g_cub_temp);
        if (cudaMalloc(&g_cub_temp, bytes) != cudaSuccess) return -1;
        g_cub_temp_bytes = bytes;
    }
    return 0;
}
}  // namespace

extern "C" {

/* algo: 0 = reference CSDL, 1 = reference RTS, 2 = reference EmulatedDeadlock (inclusive
 * only), 3 = cub::DeviceScan.  In place on d_scan, which must hold align4(size) keys
 * (the reference writes the whole last uint4).  Launches on the default stream; no
 * host sync.  Returns 0 or -1. */
int ref_scan_u32(int algo, int exclusive, uint32_t* d_scan, uint32_t size)
{
    const uint32_t alignedSize = align4(size);
    const uint32_t vectorizedSize = alignedSize / 4;
    const uint32_t blocks = div_round_up(alignedSize, kPartitionSize);
    if (size == 0) return 0;
    if (algo == 3) {
        size_t bytes = 0;
        if (exclusive)
            cub::DeviceScan::ExclusiveSum(nullptr, bytes, d_scan, d_scan, size);
        else
            cub::DeviceScan::InclusiveSum(nullptr, bytes, d_scan, size);
        if (ensure_cub_temp(bytes)) return -1;
        if (exclusive)
            cub::DeviceScan::ExclusiveSum(g_cub_temp, bytes, d_scan, d_scan, size);
        else
            cub::DeviceScan::InclusiveSum(g_cub_temp, bytes, d_scan, size);
        return cudaGetLastError() == cudaSuccess ? 0 : -1;
    }
    if (ensure_scratch(blocks)) return -1;
    switch (algo) {
        case 0:
            cudaMemset(g_index, 0, sizeof(uint32_t));
            cudaMemset(g_tbr, 0, (size_t)blocks * sizeof(uint32_t));
            if (exclusive)
                ChainedScanDecoupledLookback::CSDLExclusive<kWarps, kUint4PerThread>
                    <<<blocks, kThreads>>>(d_scan, g_tbr, g_index, vectorizedSize);
            else
                ChainedScanDecoupledLookback::CSDLInclusive<kWarps, kUint4PerThread>
                    <<<blocks, kThreads>>>(d_scan, g_tbr, g_index, vectorizedSize);
            break;
        case 1:
            ReduceThenScan::Reduce<kWarps, kUint4PerThread>
                <<<blocks, kThreads>>>(d_scan, g_tbr, vectorizedSize);
            ReduceThenScan::SpineScan<32, 4><<<1, 1024>>>(g_tbr, blocks);
            if (exclusive)
                ReduceThenScan::PropagateExclusive<kWarps, kUint4PerThread>
                    <<<blocks, kThreads>>>(d_scan, g_tbr, vectorizedSize);
            else
                ReduceThenScan::PropagateInclusive<kWarps, kUint4PerThread>
                    <<<blocks, kThreads>>>(d_scan, g_tbr, vectorizedSize);
            break;
        case 2:
            if (exclusive) return -1;
            /* the reference kernel never terminates when the last tile is a stalled
             * tile; refuse those sizes instead of hanging the box */
            if (blocks > 1 && ((blocks - 1) & 7) == 0) return -2;
            cudaMemset(g_index, 0, sizeof(uint32_t));
            cudaMemset(g_tbr, 0, (size_t)blocks * sizeof(uint32_t));
            EmulatedDeadlocking::EmulatedDeadlock<kWarps, kUint4PerThread>
                <<<blocks, kThreads>>>(d_scan, g_tbr, g_index, vectorizedSize);
            break;
        default:
            return -1;
    }
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

/* float32 peer (not an oracle): CUB only; the reference has no float kernels */
int ref_cub_scan_f32(int exclusive, float* d_scan, uint32_t size)
{
    size_t bytes = 0;
    if (size == 0) return 0;
    if (exclusive)
        cub::DeviceScan::ExclusiveSum(nullptr, bytes, d_scan, d_scan, size);
    else
        cub::DeviceScan::InclusiveSum(nullptr, bytes, d_scan, size);
    if (ensure_cub_temp(bytes)) return -1;
    if (exclusive)
        cub::DeviceScan::ExclusiveSum(g_cub_temp, bytes, d_scan, d_scan, size);
    else
        cub::DeviceScan::InclusiveSum(g_cub_temp, bytes, d_scan, size);
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

/* the reference harness kernels on caller buffers (UtilityKernels.cuh:15-38) */
void ref_init_one(uint32_t* d_scan, uint32_t size) { InitOne<<<256, 256>>>(d_scan, size); }

uint32_t ref_validate(int exclusive, uint32_t* d_scan, uint32_t size)
{
    uint32_t* d_err = nullptr;
    uint32_t h_err = 0xFFFFFFFFu;
    if (cudaMalloc(&d_err, sizeof(uint32_t)) != cudaSuccess) return h_err;
    cudaMemset(d_err, 0, sizeof(uint32_t));
    if (exclusive)
        ValidateExclusive<<<256, 256>>>(d_scan, d_err, size);
    else
        ValidateInclusive<<<256, 256>>>(d_scan, d_err, size);
    cudaMemcpy(&h_err, d_err, sizeof(uint32_t), cudaMemcpyDeviceToHost);
    cudaFree(d_err);
    return h_err;
}

/* Timing with the reference protocol (…Dispatcher.cuh:116-156): batch+1 iterations,
 * first discarded, input restored before each one (from d_src if given, else InitOne),
 * cudaDeviceSynchronize, events around the dispatch (memsets included).  Returns the
 * mean milliseconds per dispatch, or a negative value on error. */
double ref_time_u32(int algo, int exclusive, uint32_t* d_scan, const uint32_t* d_src,
                    uint32_t size, uint32_t batch)
{
    cudaEvent_t start, stop;
    cudaEventCreate(&start);
    cudaEventCreate(&stop);
    double total = 0.0;
    for (uint32_t i = 0; i <= batch; ++i) {
        if (d_src)
            cudaMemcpy(d_scan, d_src, (size_t)size * sizeof(uint32_t), cudaMemcpyDeviceToDevice);
        else
            InitOne<<<256, 256>>>(d_scan, size);
        cudaDeviceSynchronize();
        cudaEventRecord(start);
        if (ref_scan_u32(algo, exclusive, d_scan, size) != 0) return -1.0;
        cudaEventRecord(stop);
        cudaEventSynchronize(stop);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, start, stop);
        if (i) total += ms;
    }
    cudaEventDestroy(start);
    cudaEventDestroy(stop);
    if (cudaGetLastError() != cudaSuccess) return -1.0;
    return total / batch;
}

/* Self-contained timing of the reference's kernels for bench.py's reference arm, which must not
 * load the product library: allocates its own buffers, fills the source with hashed keys in
 * [0,4) (total < 2^30, the reference CSDL's limit), times CSDL / RTS / cub::DeviceScan with the
 * protocol above.  out_ms[3] = mean ms per dispatch (negative on error).  Returns 0 or -1. */
__global__ void ref_shim_fill_small(uint32_t* d, uint32_t n)
{
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        uint32_t h = i * 2654435761u;
        h ^= h >> 15;
        h *= 2246822519u;
        h ^= h >> 13;
        d[i] = h & 3u;
    }
}

int ref_bench_all(uint32_t size, uint32_t batch, double out_ms[3])
{
    uint32_t *src = nullptr, *buf = nullptr;
    out_ms[0] = out_ms[1] = out_ms[2] = -1.0;
    if (cudaMalloc(&src, (size_t)align4(size) * 4) != cudaSuccess) return -1;
    if (cudaMalloc(&buf, (size_t)align4(size) * 4) != cudaSuccess) {
        cudaFree(src);
        return -1;
    }
    ref_shim_fill_small<<<1024, 256>>>(src, size);
    cudaDeviceSynchronize();
    const int algos[3] = {0, 1, 3};
    for (int i = 0; i < 3; ++i) out_ms[i] = ref_time_u32(algos[i], 0, buf, src, size, batch);
    cudaFree(src);
    cudaFree(buf);
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

void ref_shim_release(void)
{
    if (g_index) cudaFree(g_index);
    if (g_tbr) cudaFree(g_tbr);
    if (g_cub_temp) cudaFree(g_cub_temp);
    g_index = g_tbr = nullptr;
    g_cub_temp = nullptr;
    g_tbr_blocks = 0;
    g_cub_temp_bytes = 0;
}

}  // extern "C"
