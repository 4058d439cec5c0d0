// CubDispatcher.hpp -- drop-in for the reference's CUB baseline class
// (GPUPrefixSumsCUDA/CubDispatcher.cuh:21-77): times cub::DeviceScan::InclusiveSum in
// place with the same protocol.  CUB stays library code (libgps_cub.so wraps the call
// sites CubDispatcher.cuh:44,58); nothing of it is re-implemented.
#pragma once
#include <cstdint>
#include <cstdio>

#include "../../include/gps_b200.h"

extern "C" int gps_cub_scan(int exclusive, int dtype, const void* d_in, void* d_out, uint64_t n, void* stream);

class CubDispatcher {
    const uint32_t k_maxSize;
    uint32_t* m_scan = nullptr;

   public:
    explicit CubDispatcher(uint32_t size) : k_maxSize(size)
    {
        if (gps_malloc(reinterpret_cast<void**>(&m_scan), (uint64_t)k_maxSize * sizeof(uint32_t)) != GPS_OK)
            printf("CubDispatcher: allocation failed: %s\n", gps_last_error());
    }
    ~CubDispatcher()
    {
        if (m_scan) gps_free(m_scan);
        m_scan = nullptr;
    }

    double BatchTimingCubChainedScan(uint32_t size, uint32_t batchCount)
    {
        if (size > k_maxSize) {
            printf("Error, requested test size exceeds max initialized size. \n");
            return 0.0;
        }
        printf("Beginning CUB ChainedScanDecoupledLookback inclusive batch timing test at:\n");
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
            gps_cub_scan(0, GPS_U32, m_scan, m_scan, size, nullptr);
            gps_event_record(stop, nullptr);
            float millis = 0.f;
            gps_event_elapsed_ms(start, stop, &millis);
            if (i) totalTime += millis;
            if ((i & 15) == 0) printf(". ");
        }
        gps_event_destroy(start);
        gps_event_destroy(stop);
        printf("\n");
        totalTime /= 1000.0f;
        printf("Total time elapsed: %f\n", totalTime);
        const double speed = (double)size / totalTime * batchCount;
        printf("Estimated speed at %u 32-bit elements: %E keys/sec\n\n", size, speed);
        return speed;
    }
};
