// EmulatedDeadlockingDispatcher.hpp -- drop-in for the reference class of the same name
// (GPUPrefixSumsCUDA/EmulatedDeadlockingDispatcher.cuh:19-175).  Same public methods;
// partMask / maxSpinCount are HONOURED here (the reference accepts them but its kernel
// uses the compile-time DEADLOCK_MASK 7 / MAX_SPIN_COUNT 4, EmulatedDeadlocking.cuh:15-16,
// which therefore remain the defaults of TestAllInclusive and the 2-argument timing call).
#pragma once
#include <fstream>

#include "ScanDispatcherCommon.hpp"

class EmulatedDeadlockingDispatcher : public gps_host::ScanDispatcherCommon {
    uint32_t m_partMask = 7;
    uint32_t m_maxSpinCount = 4;
    uint32_t* m_stats = nullptr;  // device: {spins, fallbacks, insertions, lookback windows}

    int Dispatch(uint32_t size, int kind) override
    {
        return gps_scan_emulated_deadlock(m_ctx, kind, GPS_U32, m_scan, m_scan, size, m_partMask,
                                          m_maxSpinCount, m_stats, nullptr);
    }

   public:
    explicit EmulatedDeadlockingDispatcher(uint32_t maxSize)
        : ScanDispatcherCommon(maxSize, GPS_ALGO_CSDL_FALLBACK, "EmulatedDeadlocking")
    {
        check(gps_malloc(reinterpret_cast<void**>(&m_stats), 4 * sizeof(uint32_t)), "gps_malloc");
        if (m_stats) gps_memset(m_stats, 0, 4 * sizeof(uint32_t), nullptr);
    }
    ~EmulatedDeadlockingDispatcher()
    {
        if (m_stats) gps_free(m_stats);
        m_stats = nullptr;
    }

    void TestAllInclusive() { TestAll(GPS_INCLUSIVE, "Beginning Emulated Deadlocking validation test: \n"); }

    double BatchTimingInclusive(uint32_t size, uint32_t batchCount, uint32_t partMask,
                                uint32_t maxSpinCount)
    {
        m_partMask = partMask;
        m_maxSpinCount = maxSpinCount;
        if (m_stats) gps_memset(m_stats, 0, 4 * sizeof(uint32_t), nullptr);
        const double speed =
            BatchTiming(size, batchCount, "Beginning Emulated Deadlocking batch timing test at:\n");
        uint32_t h[4] = {0, 0, 0, 0};
        if (m_stats && gps_memcpy_d2h(h, m_stats, sizeof(h), nullptr) == GPS_OK && gps_device_sync() == GPS_OK)
            printf("Fallback stats over %u dispatches: spins %u, fallbacks %u, insertions %u, "
                   "lookback polls %u\n\n",
                   batchCount + 1, h[0], h[1], h[2], h[3]);
        m_partMask = 7;
        m_maxSpinCount = 4;
        return speed;
    }

    void BatchTimingInclusive(uint32_t size, uint32_t batchCount)
    {
        BatchTimingInclusive(size, batchCount, 7, 4);
    }

    void RecordTimingData(uint32_t spinCount)
    {
        std::ofstream writer;
        writer.open("data.txt");
        writer << "Max Spin Count: " << spinCount << "\n";
        for (uint32_t i = 512; i >= 64; i >>= 1) {
            double t = BatchTimingInclusive(1u << 28, 250, i - 1, spinCount);
            writer << t << "\n";
        }
        writer.close();
    }
};
