// ReduceThenScanDispatcher.hpp -- drop-in for the reference class of the same name
// (GPUPrefixSumsCUDA/ReduceThenScanDispatcher.cuh:16-204) over gps_scan(GPS_ALGO_RTS).
#pragma once
#include "ScanDispatcherCommon.hpp"

class ReduceThenScanDispatcher : public gps_host::ScanDispatcherCommon {
   public:
    explicit ReduceThenScanDispatcher(uint32_t maxSize)
        : ScanDispatcherCommon(maxSize, GPS_ALGO_RTS, "ReduceThenScan")
    {
    }
    ~ReduceThenScanDispatcher() = default;

    void TestAllExclusive()
    {
        TestAll(GPS_EXCLUSIVE, "Beginning GPUPrefixSums ReduceThenScan exclusive validation test: \n");
    }
    void TestAllInclusive()
    {
        TestAll(GPS_INCLUSIVE, "Beginning GPUPrefixSums ReduceThenScan inclusive validation test: \n");
    }
    void BatchTimingInclusive(uint32_t size, uint32_t batchCount)
    {
        BatchTiming(size, batchCount,
                    "Beginning GPUPrefixSums ReduceThenScan inclusive batch timing test at:\n");
    }
};
