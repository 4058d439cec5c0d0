// gps_util_kernels.cuh -- harness kernels (input initialisation, validation).
//
// Replaces UtilityKernels.cuh:15-38 of the reference (InitOne, ValidateInclusive,
// ValidateExclusive; grid-stride) with 64-bit sizes, a float32 variant, a seeded
// counter-based random input (the reference's CUDA port only ever scans all-ones) and
// word-for-word / tolerance comparison of two device arrays.
#pragma once
#include "gps_common.cuh"

namespace gps {

__host__ __device__ __forceinline__ uint64_t mix64(uint64_t z)
{
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
// identical to oracle_random_key() in oracle/scan_oracle.c
__host__ __device__ __forceinline__ uint32_t random_key(uint64_t seed, uint64_t i, uint32_t mask)
{
    return (uint32_t)(mix64(seed + (i + 1) * 0x9E3779B97F4A7C15ULL) >> 32) & mask;
}

template <typename T>
__global__ void init_ones_kernel(T* d, uint64_t n)
{
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) d[i] = T(1);
}

__global__ void init_random_u32_kernel(uint32_t* d, uint64_t n, uint64_t seed, uint32_t mask,
                                       uint64_t first)
{
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        d[i] = random_key(seed, first + i, mask);
}

__global__ void init_random_f32_kernel(float* d, uint64_t n, uint64_t seed, uint64_t first)
{
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        d[i] = (float)(random_key(seed, first + i, 0xFFFFFFFFu) >> 8) * (1.0f / 16777216.0f);
}

// d[i] == first_value + i (+1 if inclusive), uint32 wraparound
__global__ void validate_ones_kernel(const uint32_t* d, uint64_t n, uint32_t first_value,
                                     uint32_t inclusive, unsigned long long* errs)
{
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    unsigned long long local = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        local += d[i] != (uint32_t)(first_value + (uint32_t)i + inclusive);
    if (local) atomicAdd(errs, local);
}

// errs[0] = mismatch count, errs[1] = smallest mismatching index
__global__ void compare_u32_kernel(const uint32_t* a, const uint32_t* b, uint64_t n,
                                   unsigned long long* errs)
{
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    unsigned long long local = 0, first = ~0ULL;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        if (a[i] != b[i]) {
            ++local;
            if (i < first) first = i;
        }
    if (local) {
        atomicAdd(errs, local);
        atomicMin(errs + 1, first);
    }
}

// counts |a-b| > rtol*|bound| + atol (NaN counts); maxdiff via atomicMax on float bits
__global__ void compare_f32_kernel(const float* a, const float* b, const float* bound, uint64_t n,
                                   float rtol, float atol, unsigned long long* errs,
                                   uint32_t* max_abs_bits)
{
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    unsigned long long local = 0;
    float mx = 0.f;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const float d = fabsf(a[i] - b[i]);
        const float lim = rtol * fabsf(bound ? bound[i] : b[i]) + atol;
        if (!(d <= lim)) ++local;
        if (d > mx) mx = d;
    }
    if (local) atomicAdd(errs, local);
    if (mx > 0.f) atomicMax(max_abs_bits, __float_as_uint(mx));  // non-negative floats order as uints
}

__global__ void write_value_kernel(uint32_t* dst, uint32_t bits) { *dst = bits; }

// struct scan of all-ones input (GPUPrefixSumsWebGPUapis/SharedShaders/validate.wgsl:50-60):
// flat u32 view, out[i] == (i >> 2) + 1 (inclusive) or (i >> 2) (exclusive); n4 = 4 * elements
__global__ void validate_struct_ones_kernel(const uint32_t* d, uint64_t n4, uint32_t inclusive,
                                            unsigned long long* errs)
{
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    unsigned long long local = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride)
        local += d[i] != (uint32_t)(i >> 2) + inclusive;
    if (local) atomicAdd(errs, local);
}

// uint64 scan of all-ones input: d[i] == i + 1 / i
__global__ void validate_ones_u64_kernel(const uint64_t* d, uint64_t n, uint64_t inclusive, unsigned long long* errs)
{
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    unsigned long long local = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        local += d[i] != i + inclusive;
    if (local) atomicAdd(errs, local);
}

// sortedness + multiset fingerprint of a sort: errs[0] = inversions between neighbours,
// errs[1] = sum of keys, errs[2] = sum of key * (key >> 7 | 1) (both mod 2^64)
template <typename K>  // uint32_t or uint64_t keys
__global__ void sort_fingerprint_kernel(const K* d, uint64_t n, unsigned long long* errs)
{
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    unsigned long long inv = 0, s1 = 0, s2 = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const K k = d[i];
        if (i && d[i - 1] > k) ++inv;
        s1 += k;
        s2 += (unsigned long long)k * ((unsigned long long)(k >> 7) | 1ull);
    }
    if (inv) atomicAdd(errs, inv);
    atomicAdd(errs + 1, s1);
    atomicAdd(errs + 2, s2);
}

}  // namespace gps
