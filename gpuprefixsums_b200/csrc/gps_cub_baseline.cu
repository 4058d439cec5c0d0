// gps_cub_baseline.cu -- comparison baseline only: cub::DeviceScan from the CUDA
// toolkit's CCCL (library code, not re-implemented), behind a tiny C ABI so the
// CubDispatcher mirror and bench.py can time it next to the new kernels.
// Replaces the call sites CubDispatcher.cuh:44,58 of the reference.
// Built into libgps_cub.so, separate from the product library.
#include <cub/device/device_radix_sort.cuh>
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

// cub::DeviceRadixSort (keys, or pairs when d_vals_in != NULL): comparison baseline for
// gps_sort_u32; out of place, asynchronous on `stream`
int gps_cub_sort_u32(const void* d_keys_in, void* d_keys_out, const void* d_vals_in, void* d_vals_out, uint64_t n,
                     int end_bit, void* stream)
{
    if (n == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    const uint32_t* ki = (const uint32_t*)d_keys_in;
    uint32_t* ko = (uint32_t*)d_keys_out;
    const uint32_t* vi = (const uint32_t*)d_vals_in;
    uint32_t* vo = (uint32_t*)d_vals_out;
    size_t need = 0;
    cudaError_t e = vi ? cub::DeviceRadixSort::SortPairs(nullptr, need, ki, ko, vi, vo, (int64_t)n, 0, end_bit, st)
                       : cub::DeviceRadixSort::SortKeys(nullptr, need, ki, ko, (int64_t)n, 0, end_bit, st);
    if (e != cudaSuccess) return -1;
    if (need > g_temp_bytes) {
        if (g_temp) cudaFree(g_temp);
        if (cudaMalloc(&g_temp, need) != cudaSuccess) return -1;
        g_temp_bytes = need;
    }
    e = vi ? cub::DeviceRadixSort::SortPairs(g_temp, need, ki, ko, vi, vo, (int64_t)n, 0, end_bit, st)
           : cub::DeviceRadixSort::SortKeys(g_temp, need, ki, ko, (int64_t)n, 0, end_bit, st);
    return e == cudaSuccess ? 0 : -1;
}

// the same for uint64 keys (32-bit payload)
int gps_cub_sort_u64(const void* d_keys_in, void* d_keys_out, const void* d_vals_in, void* d_vals_out, uint64_t n,
                     int end_bit, void* stream)
{
    if (n == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    const uint64_t* ki = (const uint64_t*)d_keys_in;
    uint64_t* ko = (uint64_t*)d_keys_out;
    const uint32_t* vi = (const uint32_t*)d_vals_in;
    uint32_t* vo = (uint32_t*)d_vals_out;
    size_t need = 0;
    cudaError_t e = vi ? cub::DeviceRadixSort::SortPairs(nullptr, need, ki, ko, vi, vo, (int64_t)n, 0, end_bit, st)
                       : cub::DeviceRadixSort::SortKeys(nullptr, need, ki, ko, (int64_t)n, 0, end_bit, st);
    if (e != cudaSuccess) return -1;
    if (need > g_temp_bytes) {
        if (g_temp) cudaFree(g_temp);
        if (cudaMalloc(&g_temp, need) != cudaSuccess) return -1;
        g_temp_bytes = need;
    }
    e = vi ? cub::DeviceRadixSort::SortPairs(g_temp, need, ki, ko, vi, vo, (int64_t)n, 0, end_bit, st)
           : cub::DeviceRadixSort::SortKeys(g_temp, need, ki, ko, (int64_t)n, 0, end_bit, st);
    return e == cudaSuccess ? 0 : -1;
}

void gps_cub_release(void)
{
    if (g_temp) cudaFree(g_temp);
    g_temp = nullptr;
    g_temp_bytes = 0;
}

}  // extern "C"
