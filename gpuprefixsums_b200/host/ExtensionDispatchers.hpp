// ExtensionDispatchers.hpp -- C++ hosts for the rows SURVEY.md 8(f) marks "next", in the shape
// of the reference's dispatcher classes (same TestAll / BatchTiming protocol and stdout
// lines as ChainedScanDecoupledLookbackDispatcher.cuh:38-156), pure C++17 over the C ABI:
//   StructScanDispatcher  four-lane vec4<u32> scan; test = the WebGPU port's struct-scan
//                         protocol (GPUPrefixSumsWebGPUapis/SharedShaders/csdldf_struct.wgsl,
//                         validate.wgsl:50-60: all-ones input, flat out[i] == (i >> 2) + 1)
//   Uint64ScanDispatcher  64-bit keys, all-ones closed form
//   RadixSortDispatcher   the scan's consumer (README.md:7): random keys, checked for
//                         sortedness and an unchanged key multiset on the device
#pragma once
#include <cstdint>
#include <cstdio>

#include "../../include/gps_b200.h"

namespace gps_host {

class ExtensionBase {
   protected:
    const uint64_t k_maxBytes;
    const char* const k_name;
    gps_ctx* m_ctx = nullptr;
    void *m_a = nullptr, *m_b = nullptr, *m_c = nullptr;
    bool m_ok = true;

    bool check(int rc, const char* what)
    {
        if (rc != GPS_OK) {
            printf("\n%s: %s failed (%d): %s\n", k_name, what, rc, gps_last_error());
            m_ok = false;
        }
        return rc == GPS_OK;
    }
    ExtensionBase(uint64_t maxBytes, int buffers, const char* name) : k_maxBytes(maxBytes), k_name(name)
    {
        check(gps_create(&m_ctx, 0, maxBytes / 4), "gps_create");
        check(gps_malloc(&m_a, maxBytes), "gps_malloc");
        if (buffers > 1) check(gps_malloc(&m_b, maxBytes), "gps_malloc");
        if (buffers > 2) check(gps_malloc(&m_c, maxBytes), "gps_malloc");
    }
    ~ExtensionBase()
    {
        if (m_a) gps_free(m_a);
        if (m_b) gps_free(m_b);
        if (m_c) gps_free(m_c);
        if (m_ctx) gps_destroy(m_ctx);
    }
    template <typename F>
    double timeLoop(uint64_t elems, uint32_t batchCount, const char* unit, F&& body)
    {
        void *start = nullptr, *stop = nullptr;
        gps_event_create(&start);
        gps_event_create(&stop);
        float totalTime = 0.0f;
        for (uint32_t i = 0; i <= batchCount; ++i) {
            gps_device_sync();
            gps_event_record(start, nullptr);
            body();
            gps_event_record(stop, nullptr);
            float millis = 0.f;
            gps_event_elapsed_ms(start, stop, &millis);
            if (i) totalTime += millis;
        }
        check(gps_check(m_ctx, nullptr), "gps_check");
        gps_event_destroy(start);
        gps_event_destroy(stop);
        totalTime /= 1000.0f;
        const double speed = (double)elems / totalTime * batchCount;
        printf("Total time elapsed: %f\n", totalTime);
        printf("Estimated speed at %llu %s: %E per sec\n\n", (unsigned long long)elems, unit, speed);
        return speed;
    }

   public:
    bool ok() const { return m_ok; }
};

// wide scans: all-ones protocol, in place like the reference dispatchers
class WideScanDispatcher : public ExtensionBase {
    const int k_dtype;
    const uint32_t k_elemBytes;

   protected:
    int m_timingAlgo = GPS_ALGO_CSDL;
    WideScanDispatcher(uint64_t maxElems, int dtype, uint32_t elemBytes, const char* name)
        : ExtensionBase(maxElems * elemBytes, 1, name), k_dtype(dtype), k_elemBytes(elemBytes)
    {
    }
    bool runOne(uint64_t elems, int kind, int algo)
    {
        // all-ones input: uint32 ones over the flat view for the struct scan, 64-bit ones otherwise
        if (k_dtype == GPS_U32X4) {
            check(gps_init_ones(m_a, elems * 4, GPS_U32, nullptr), "gps_init_ones");
        } else {
            check(gps_memset(m_a, 0, elems * 8, nullptr), "gps_memset");
            check(gps_init_ones_strided(m_a, elems), "init");
        }
        check(gps_scan_wide(m_ctx, algo, kind, k_dtype, m_a, m_a, elems, nullptr, nullptr, nullptr), "gps_scan_wide");
        uint64_t errs = ~0ull;
        if (!check(gps_check(m_ctx, nullptr), "gps_check")) return false;
        if (!check(gps_validate_ones_wide(m_a, elems, kind, k_dtype, &errs), "validate")) return false;
        return errs == 0;
    }
    // little-endian 64-bit ones = uint32 {1, 0} pairs: write ones to the even words
    int gps_init_ones_strided(void* d, uint64_t elems)
    {
        // one cudaMemset2D-free way over the C ABI: scan-independent, uses the harness kernels only
        static const uint32_t one[2] = {1u, 0u};
        if (elems == 0) return GPS_OK;
        int rc = gps_memcpy_h2d(d, one, sizeof(one), nullptr);
        // doubling copy on the device
        for (uint64_t have = 1; rc == GPS_OK && have < elems; have *= 2) {
            const uint64_t take = have < elems - have ? have : elems - have;
            rc = gps_memcpy_d2d(static_cast<char*>(d) + have * 8, d, take * 8, nullptr);
        }
        return rc;
    }

   public:
    void TestAll()
    {
        printf("Beginning %s validation test: \n", k_name);
        uint32_t passed = 0, total = 0;
        const uint64_t tile = 36864 / k_elemBytes;  // elements per tile of the compiled kernel
        // chained scan, chained scan with Decoupled Fallback (what csdldf_struct.wgsl is), Reduce-then-Scan
        for (int algo : {GPS_ALGO_CSDL, GPS_ALGO_CSDL_FALLBACK, GPS_ALGO_RTS})
            for (int kind : {GPS_INCLUSIVE, GPS_EXCLUSIVE}) {
                for (uint64_t n = tile - 3; n <= 2 * tile + 3; n += (n < tile + 3 || n > 2 * tile - 4) ? 1 : tile / 7) {
                    ++total;
                    if (runOne(n, kind, algo)) ++passed;
                    else printf("\n Test failed at size %llu \n", (unsigned long long)n);
                }
                ++total;
                if (runOne(k_maxBytes / k_elemBytes, kind, algo)) ++passed;
                else printf("\n Test failed at size %llu \n", (unsigned long long)(k_maxBytes / k_elemBytes));
                printf(".");
            }
        // the emulated-deadlock harness of the wide scans: reference settings (mask 7, spin 4)
        for (uint64_t n : {tile * 9 + 1, tile * 64, k_maxBytes / k_elemBytes}) {
            ++total;
            check(gps_init_ones_strided(m_a, n), "init");
            if (k_dtype == GPS_U32X4) check(gps_init_ones(m_a, n * 4, GPS_U32, nullptr), "gps_init_ones");
            check(gps_scan_wide_emulated_deadlock(m_ctx, GPS_INCLUSIVE, k_dtype, m_a, m_a, n, nullptr, 7, 4, nullptr, nullptr),
                  "gps_scan_wide_emulated_deadlock");
            uint64_t errs = ~0ull;
            if (check(gps_check(m_ctx, nullptr), "gps_check") &&
                check(gps_validate_ones_wide(m_a, n, GPS_INCLUSIVE, k_dtype, &errs), "validate") && errs == 0)
                ++passed;
            else
                printf("\n Emulated-deadlock test failed at size %llu \n", (unsigned long long)n);
        }
        if (passed == total) printf("\n%u/%u All tests passed.\n\n", passed, total);
        else printf("\n%u/%u Test failed.\n\n", passed, total);
        if (passed != total) m_ok = false;
    }
    double BatchTiming(uint64_t elems, uint32_t batchCount)
    {
        printf("Beginning %s inclusive batch timing test at:\nSize: %llu\nTest size: %u\n", k_name,
               (unsigned long long)elems, batchCount);
        return timeLoop(elems, batchCount, "elements", [&] {
            check(gps_scan_wide(m_ctx, m_timingAlgo, GPS_INCLUSIVE, k_dtype, m_a, m_a, elems, nullptr, nullptr, nullptr),
                  "gps_scan_wide");
        });
    }
};

class StructScanDispatcher : public WideScanDispatcher {
   public:
    // timed with the Decoupled Fallback variant: that is what the reference's struct scan is
    // (csdldf_struct.wgsl), and it runs within 1 % of the plain chained kernel
    explicit StructScanDispatcher(uint64_t maxElems) : WideScanDispatcher(maxElems, GPS_U32X4, 16, "GPS vec4<u32> struct scan")
    {
        m_timingAlgo = GPS_ALGO_CSDL_FALLBACK;
    }
};
class Uint64ScanDispatcher : public WideScanDispatcher {
   public:
    explicit Uint64ScanDispatcher(uint64_t maxElems) : WideScanDispatcher(maxElems, GPS_U64, 8, "GPS uint64 scan") {}
};

class RadixSortDispatcher : public ExtensionBase {
    bool runOne(uint64_t n, uint64_t seed, bool pairs)
    {
        check(gps_init_random(m_a, n, GPS_U32, seed, 0xFFFFFFFFu, 0, nullptr), "gps_init_random");
        uint64_t inv = 0, before[2], after[2];
        check(gps_sort_fingerprint(m_a, n, &inv, before), "fingerprint");
        // pairs: the payload is a copy of the keys, so it must come out equal to the sorted keys
        void* vals_in = pairs ? m_a : nullptr;
        void* vals = pairs ? static_cast<char*>(m_c) + k_maxBytes / 2 : nullptr;
        void* vals_tmp = pairs ? m_c : nullptr;
        check(gps_sort_u32(m_ctx, m_a, m_b, pairs ? static_cast<char*>(m_b) + k_maxBytes / 2 : m_c, vals_in, vals, vals_tmp, n,
                           32, nullptr),
              "gps_sort_u32");
        if (!check(gps_check(m_ctx, nullptr), "gps_check")) return false;
        check(gps_sort_fingerprint(m_b, n, &inv, after), "fingerprint");
        bool good = inv == 0 && before[0] == after[0] && before[1] == after[1];
        if (good && pairs) {
            uint64_t errs = ~0ull, first = 0;
            check(gps_compare_u32(m_b, vals, n, &errs, &first), "gps_compare_u32");
            good = errs == 0;
        }
        return good;
    }

   public:
    // three buffers of maxKeys keys; pair tests use half of each for keys and half for the payload
    explicit RadixSortDispatcher(uint64_t maxKeys) : ExtensionBase(maxKeys * 4, 3, "GPS radix sort") {}

    void TestAll()
    {
        printf("Beginning GPS radix sort validation test: \n");
        uint32_t passed = 0, total = 0;
        const uint64_t maxKeys = k_maxBytes / 4;
        for (int pairs = 0; pairs < 2; ++pairs) {
            for (uint64_t n = 10237; n <= 10243; ++n) {
                ++total;
                if (runOne(n, 100 + n, pairs)) ++passed;
                else printf("\n Test failed at size %llu \n", (unsigned long long)n);
            }
            for (uint64_t n : {(uint64_t)1, (uint64_t)255, (uint64_t)1 << 20, pairs ? maxKeys / 2 : maxKeys}) {
                ++total;
                if (runOne(n, 7 + n, pairs)) ++passed;
                else printf("\n Test failed at size %llu \n", (unsigned long long)n);
            }
            printf(".");
        }
        if (passed == total) printf("\n%u/%u All tests passed.\n\n", passed, total);
        else printf("\n%u/%u Test failed.\n\n", passed, total);
        if (passed != total) m_ok = false;
    }
    double BatchTiming(uint64_t n, uint32_t batchCount)
    {
        printf("Beginning GPS radix sort (keys) batch timing test at:\nSize: %llu\nTest size: %u\n", (unsigned long long)n,
               batchCount);
        check(gps_init_random(m_a, n, GPS_U32, 1, 0xFFFFFFFFu, 0, nullptr), "gps_init_random");
        return timeLoop(n, batchCount, "32-bit keys",
                        [&] { check(gps_sort_u32(m_ctx, m_a, m_b, m_c, nullptr, nullptr, nullptr, n, 32, nullptr), "gps_sort_u32"); });
    }
};

// 64-bit keys (gps_sort_keys64): same protocol; the key buffers are filled as 2 n random 32-bit words
class RadixSort64Dispatcher : public ExtensionBase {
    bool runOne(uint64_t n, uint64_t seed, bool pairs, int key_type)
    {
        check(gps_init_random(m_a, 2 * n, GPS_U32, seed, 0xFFFFFFFFu, 0, nullptr), "gps_init_random");
        uint64_t inv = 0, before[2], after[2];
        check(gps_sort_fingerprint64(m_a, n, &inv, before), "fingerprint");
        // pairs: the payload is the first n 32-bit words of the key buffer; keys and payload must each come out as
        // the same multiset (fingerprints), the keys in order
        char* b = static_cast<char*>(m_b);
        char* c = static_cast<char*>(m_c);
        void* vals_in = pairs ? m_a : nullptr;              // the first n 32-bit words of the key buffer
        void* vals = pairs ? b + k_maxBytes / 2 : nullptr;  // keys take at most half of each buffer in pair mode
        void* vals_tmp = pairs ? c + k_maxBytes / 2 : nullptr;
        uint64_t vinv = 0, vbefore[2] = {0, 0}, vafter[2] = {0, 0};
        if (pairs) check(gps_sort_fingerprint(vals_in, n, &vinv, vbefore), "fingerprint");
        check(gps_sort_keys64(m_ctx, key_type, m_a, m_b, m_c, vals_in, vals, vals_tmp, n, 64, nullptr), "gps_sort_keys64");
        if (!check(gps_check(m_ctx, nullptr), "gps_check")) return false;
        check(gps_sort_fingerprint64(m_b, n, &inv, after), "fingerprint");
        // signed order = unsigned order of key ^ 2^63: inversions are only counted for unsigned keys
        bool good = (key_type != GPS_KEY_U32 || inv == 0) && before[0] == after[0] && before[1] == after[1];
        if (good && pairs) {
            check(gps_sort_fingerprint(vals, n, &vinv, vafter), "fingerprint");
            good = vbefore[0] == vafter[0] && vbefore[1] == vafter[1];
        }
        return good;
    }

   public:
    // three buffers of maxKeys 64-bit keys; pair tests use half of each for keys and half for the payload
    explicit RadixSort64Dispatcher(uint64_t maxKeys) : ExtensionBase(maxKeys * 8, 3, "GPS radix sort, 64-bit keys") {}

    void TestAll()
    {
        printf("Beginning GPS radix sort (64-bit keys) validation test: \n");
        uint32_t passed = 0, total = 0;
        const uint64_t maxKeys = k_maxBytes / 8;
        for (int pairs = 0; pairs < 2; ++pairs) {
            const uint64_t tile = pairs ? 5120 : 6144;
            for (uint64_t n = tile - 3; n <= tile + 3; ++n) {
                ++total;
                if (runOne(n, 100 + n, pairs, GPS_KEY_U32)) ++passed;
                else printf("\n Test failed at size %llu \n", (unsigned long long)n);
            }
            for (uint64_t n : {(uint64_t)1, (uint64_t)255, (uint64_t)1 << 20, pairs ? maxKeys / 4 : maxKeys}) {
                ++total;
                if (runOne(n, 7 + n, pairs, n == 255 ? GPS_KEY_I32 : GPS_KEY_U32)) ++passed;
                else printf("\n Test failed at size %llu \n", (unsigned long long)n);
            }
            printf(".");
        }
        if (passed == total) printf("\n%u/%u All tests passed.\n\n", passed, total);
        else printf("\n%u/%u Test failed.\n\n", passed, total);
        if (passed != total) m_ok = false;
    }
    double BatchTiming(uint64_t n, uint32_t batchCount)
    {
        printf("Beginning GPS radix sort (64-bit keys) batch timing test at:\nSize: %llu\nTest size: %u\n", (unsigned long long)n,
               batchCount);
        check(gps_init_random(m_a, 2 * n, GPS_U32, 1, 0xFFFFFFFFu, 0, nullptr), "gps_init_random");
        return timeLoop(n, batchCount, "64-bit keys", [&] {
            check(gps_sort_keys64(m_ctx, GPS_KEY_U32, m_a, m_b, m_c, nullptr, nullptr, nullptr, n, 64, nullptr), "gps_sort_keys64");
        });
    }
};

}  // namespace gps_host
using gps_host::RadixSort64Dispatcher;
using gps_host::RadixSortDispatcher;
using gps_host::StructScanDispatcher;
using gps_host::Uint64ScanDispatcher;
