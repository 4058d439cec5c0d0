// ScanDispatcherCommon.hpp -- shared host logic of the dispatcher classes that mirror
// the reference's operator surface.  Pure C++17 over the C ABI in include/gps_b200.h:
// no CUDA headers, no kernels, compiled with g++ and linked against libgps_b200.so.
//
// The reference repeats this logic in each dispatcher
// (ChainedScanDecoupledLookbackDispatcher.cuh:38-156, ReduceThenScanDispatcher.cuh:36-150,
// EmulatedDeadlockingDispatcher.cuh:41-122); here it lives once, parameterised by the
// algorithm.  Behaviour kept: the dispatcher owns the data buffer (maxSize rounded up
// to 4 keys) and scans it in place; TestAll* sweeps every size in [4096, 8192] on
// all-ones input and checks the closed form, then 2^26..2^28; BatchTiming* runs
// batchCount+1 iterations, discards the first, re-initialises before each one and times
// the dispatch alone with events; the same stdout lines.  Deliberate fixes: the large
// tests initialise and validate all 1<<i keys (the reference passes `i`, i.e. 26..28
// keys, ...Dispatcher.cuh:61-64), and errors from the library are reported.
#pragma once
#include <cstdint>
#include <cstdio>

#include "../../include/gps_b200.h"

namespace gps_host {

class ScanDispatcherCommon {
   protected:
    static constexpr uint32_t k_partitionSize = 4096;  // the reference's tile: test range only
    const uint32_t k_maxSize;
    const int k_algo;
    const char* const k_name;
    gps_ctx* m_ctx = nullptr;
    uint32_t* m_scan = nullptr;
    bool m_ok = true;

    static uint32_t divRoundUp(uint32_t x, uint32_t y) { return (x + y - 1) / y; }
    static uint32_t align16(uint32_t x) { return divRoundUp(x, 4) * 4; }

    bool check(int rc, const char* what)
    {
        if (rc != GPS_OK) {
            printf("\n%s: %s failed (%d): %s\n", k_name, what, rc, gps_last_error());
            m_ok = false;
        }
        return rc == GPS_OK;
    }

    ScanDispatcherCommon(uint32_t maxSize, int algo, const char* name)
        : k_maxSize(align16(maxSize)), k_algo(algo), k_name(name)
    {
        check(gps_create(&m_ctx, 0, k_maxSize), "gps_create");
        check(gps_malloc(reinterpret_cast<void**>(&m_scan), (uint64_t)k_maxSize * sizeof(uint32_t)),
              "gps_malloc");
    }

    virtual ~ScanDispatcherCommon()
    {
        if (m_scan) gps_free(m_scan);
        if (m_ctx) gps_destroy(m_ctx);
        m_scan = nullptr;
        m_ctx = nullptr;
    }

    // overridden by the emulated-deadlock dispatcher
    virtual int Dispatch(uint32_t size, int kind)
    {
        return gps_scan(m_ctx, k_algo, kind, GPS_U32, m_scan, m_scan, size, 0, nullptr, nullptr);
    }

    void DispatchKernelsInclusive(uint32_t size) { check(Dispatch(size, GPS_INCLUSIVE), "dispatch"); }
    void DispatchKernelsExclusive(uint32_t size) { check(Dispatch(size, GPS_EXCLUSIVE), "dispatch"); }

    bool DispatchValidate(uint32_t size, int kind)
    {
        uint64_t errs = ~0ull;
        if (!check(gps_check(m_ctx, nullptr), "gps_check")) return false;
        if (!check(gps_validate_ones(m_scan, size, kind, 0, &errs), "gps_validate_ones")) return false;
        return errs == 0;
    }
    bool DispatchValidateInclusive(uint32_t size) { return DispatchValidate(size, GPS_INCLUSIVE); }
    bool DispatchValidateExclusive(uint32_t size) { return DispatchValidate(size, GPS_EXCLUSIVE); }

    void TestAll(int kind, const char* banner)
    {
        if (k_maxSize < (1u << 28)) {
            printf("This test requires a minimum initialized size of %u. ", 1u << 28);
            printf("Reinitialize the object to at least %u.\n", 1u << 28);
            return;
        }
        printf("%s", banner);
        uint32_t testsPassed = 0;
        for (uint32_t i = k_partitionSize; i < k_partitionSize * 2 + 1; ++i) {
            gps_init_ones(m_scan, i, GPS_U32, nullptr);
            check(Dispatch(i, kind), "dispatch");
            if (DispatchValidate(i, kind))
                testsPassed++;
            else
                printf("\n Test failed at size %u \n", i);
            if (!(i & 255)) printf(".");
        }
        printf("\n");
        for (uint32_t i = 26; i <= 28; ++i) {
            gps_init_ones(m_scan, 1u << i, GPS_U32, nullptr);
            check(Dispatch(1u << i, kind), "dispatch");
            if (DispatchValidate(1u << i, kind))
                testsPassed++;
            else
                printf("\n Test failed at size %u \n", 1u << i);
        }
        if (testsPassed == k_partitionSize + 3 + 1)
            printf("%u/%u All tests passed.\n\n", testsPassed, testsPassed);
        else
            printf("%u/%u Test failed.\n\n", testsPassed, k_partitionSize + 3 + 1);
    }

    // returns keys/sec (0 on error)
    double BatchTiming(uint32_t size, uint32_t batchCount, const char* banner)
    {
        if (size > k_maxSize) {
            printf("Error, requested test size exceeds max initialized size. \n");
            return 0.0;
        }
        printf("%s", banner);
        printf("Size: %u\n", size);
        printf("Test size: %u\n", batchCount);
        void *start = nullptr, *stop = nullptr;
        gps_event_create(&start);
        gps_event_create(&stop);
        float totalTime = 0.0f;
        for (uint32_t i = 0; i <= batchCount; ++i) {
            gps_init_ones(m_scan, size, GPS_U32, nullptr);
            gps_device_sync();
            gps_event_record(start, nullptr);
            DispatchKernelsInclusive(size);
            gps_event_record(stop, nullptr);
            float millis = 0.f;
            gps_event_elapsed_ms(start, stop, &millis);
            if (i) totalTime += millis;
            if ((i & 15) == 0) printf(". ");
        }
        check(gps_check(m_ctx, nullptr), "gps_check");
        gps_event_destroy(start);
        gps_event_destroy(stop);
        printf("\n");
        totalTime /= 1000.0f;
        printf("Total time elapsed: %f\n", totalTime);
        const double speed = (double)size / totalTime * batchCount;
        printf("Estimated speed at %u 32-bit elements: %E keys/sec\n\n", size, speed);
        return speed;
    }

   public:
    bool ok() const { return m_ok; }
};

}  // namespace gps_host
