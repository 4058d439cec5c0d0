// ChainedScanDecoupledLookbackDispatcher.hpp -- drop-in for the reference class of the
// same name (GPUPrefixSumsCUDA/ChainedScanDecoupledLookbackDispatcher.cuh:16-212):
// same constructor, TestAllExclusive / TestAllInclusive / BatchTimingInclusive and
// destructor, now a thin shell over the C ABI (gps_scan with GPS_ALGO_CSDL).
#pragma once
#include "ScanDispatcherCommon.hpp"

class ChainedScanDecoupledLookbackDispatcher : public gps_host::ScanDispatcherCommon {
   public:
    explicit ChainedScanDecoupledLookbackDispatcher(uint32_t maxSize, int algo = GPS_ALGO_CSDL)
        : ScanDispatcherCommon(maxSize, algo, "ChainedScanDecoupledLookback")
    {
    }
    ~ChainedScanDecoupledLookbackDispatcher() = default;

    void TestAllExclusive()
    {
        TestAll(GPS_EXCLUSIVE,
                "Beginning GPUPrefixSums ChainedScanDecoupledLookback exclusive validation test: \n");
    }
    void TestAllInclusive()
    {
        TestAll(GPS_INCLUSIVE,
                "Beginning GPUPrefixSums ChainedScanDecoupledLookback inclusive validation test: \n");
    }
    void BatchTimingInclusive(uint32_t size, uint32_t batchCount)
    {
        BatchTiming(size, batchCount,
                    "Beginning GPUPrefixSums ChainedScanDecoupledLookback inclusive batch timing test "
                    "at:\n");
    }
};
