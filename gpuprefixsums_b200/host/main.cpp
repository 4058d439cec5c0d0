// main.cpp -- the reference's driver (GPUPrefixSumsCUDA/GPUPrefixSumsCUDA.cu:14-38) on
// the new library: constructs each dispatcher at 1<<28, runs TestAll* and
// BatchTimingInclusive(1<<28, 100).  `gps_main --quick` skips the 3 x 4097-size sweeps
// (timing only); `--algo reg|tma` pins the CSDL variant; `--ext` runs ONLY the extension
// dispatchers (wide scans, radix sort: SURVEY.md 8f) with the same test / timing protocol.
#include <cstring>

#include "ChainedScanDecoupledLookbackDispatcher.hpp"
#include "CubDispatcher.hpp"
#include "EmulatedDeadlockingDispatcher.hpp"
#include "ExtensionDispatchers.hpp"
#include "ReduceThenScanDispatcher.hpp"

int main(int argc, char** argv)
{
    bool quick = false, ext = false;
    int algo = GPS_ALGO_CSDL;
    for (int i = 1; i < argc; ++i) {
        if (!strcmp(argv[i], "--quick")) quick = true;
        else if (!strcmp(argv[i], "--ext")) ext = true;
        else if (!strcmp(argv[i], "--algo") && i + 1 < argc) {
            ++i;
            algo = !strcmp(argv[i], "reg") ? GPS_ALGO_CSDL_REG : !strcmp(argv[i], "tma") ? GPS_ALGO_CSDL_TMA : GPS_ALGO_CSDL;
        }
    }
    bool ok = true;
    if (ext) {
        {
            StructScanDispatcher st(1 << 26);  // 1 GiB of vec4<u32>
            st.TestAll();
            st.BatchTiming(1 << 26, 50);
            ok = ok && st.ok();
        }
        {
            Uint64ScanDispatcher u64(1 << 27);
            u64.TestAll();
            u64.BatchTiming(1 << 27, 50);
            ok = ok && u64.ok();
        }
        {
            RadixSortDispatcher rs(1 << 28);
            rs.TestAll();
            rs.BatchTiming(1 << 28, 20);
            ok = ok && rs.ok();
        }
        {
            RadixSort64Dispatcher rs64(1 << 27);
            rs64.TestAll();
            rs64.BatchTiming(1 << 27, 10);
            ok = ok && rs64.ok();
        }
        return ok ? 0 : 1;
    }
    {
        ChainedScanDecoupledLookbackDispatcher* csdl = new ChainedScanDecoupledLookbackDispatcher(1 << 28, algo);
        if (!quick) {
            csdl->TestAllExclusive();
            csdl->TestAllInclusive();
        }
        csdl->BatchTimingInclusive(1 << 28, 100);
        ok = ok && csdl->ok();
        delete csdl;
    }
    {
        ReduceThenScanDispatcher* rts = new ReduceThenScanDispatcher(1 << 28);
        if (!quick) {
            rts->TestAllExclusive();
            rts->TestAllInclusive();
        }
        rts->BatchTimingInclusive(1 << 28, 100);
        ok = ok && rts->ok();
        delete rts;
    }
    {
        EmulatedDeadlockingDispatcher* ed = new EmulatedDeadlockingDispatcher(1 << 28);
        if (!quick) ed->TestAllInclusive();
        ed->BatchTimingInclusive(1 << 28, 100);
        ok = ok && ed->ok();
        delete ed;
    }
    {
        CubDispatcher* cub = new CubDispatcher(1 << 28);
        cub->BatchTimingCubChainedScan(1 << 28, 100);
        delete cub;
    }
    return ok ? 0 : 1;
}
