// gps_cub_baseline.cu -- comparison baseline only: cub::DeviceScan from the CUDA
// toolkit's CCCL (library code, not re-implemented), behind a tiny C ABI so the
// CubDispatcher mirror and bench.py can time it next to the new kernels.
// Replaces the call sites CubDispatcher.cuh:44,58 of the reference.
// Built into libgps_cub.so, separate from the product library.
#include <cub/device/device_scan.cuh>
#include <cuda_runtime.h>
#include <stdint.h>

namespace {
void* g_temp = nullptr;
size_t g_temp_bytes = 0;

template <typename T>
int scan_impl(int exclusive, const T* in, T* out, uint64_t n, cudaStream_t st)
{
    size_t need = 0;
    cudaError_t e = exclusive ? cub::DeviceScan::ExclusiveSum(nullptr, need, in, out, (int64_t)n, st)
                              : cub::DeviceScan::InclusiveSum(nullptr, need, in, out, (int64_t)n, st);
    if (e != cudaSuccess) return -1;
    if (need > g_temp_bytes) {
        if (g_temp) cudaFree(g_temp);
        if (cudaMalloc(&g_temp, need) != cudaSuccess) return -1;
        g_temp_bytes = need;
    }
    e = exclusive ? cub::DeviceScan::ExclusiveSum(g_temp, need, in, out, (int64_t)n, st)
                  : cub::DeviceScan::InclusiveSum(g_temp, need, in, out, (int64_t)n, st);
    return e == cudaSuccess ? 0 : -1;
}
}  // namespace

extern "C" {

// dtype: 0 = uint32, 1 = float32; d_out may equal d_in; asynchronous on `stream`
int gps_cub_scan(int exclusive, int dtype, const void* d_in, void* d_out, uint64_t n, void* stream)
{
    if (n == 0) return 0;
    if (dtype == 0)
        return scan_impl<uint32_t>(exclusive, (const uint32_t*)d_in, (uint32_t*)d_out, n, (cudaStream_t)stream);
    if (dtype == 1)
        return scan_impl<float>(exclusive, (const float*)d_in, (float*)d_out, n, (cudaStream_t)stream);
    return -1;
}

void gps_cub_release(void)
{
    if (g_temp) cudaFree(g_temp);
    g_temp = nullptr;
    g_temp_bytes = 0;
}

}  // extern "C"
