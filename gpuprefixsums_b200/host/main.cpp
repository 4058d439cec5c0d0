// main.cpp -- the reference's driver (GPUPrefixSumsCUDA/GPUPrefixSumsCUDA.cu:14-38) on
// the new library: constructs each dispatcher at 1<<28, runs TestAll* and
// BatchTimingInclusive(1<<28, 100).  `gps_main --quick` skips the 3 x 4097-size sweeps
// (timing only); `--algo reg|tma` pins the CSDL variant.
#include <cstring>

#include "ChainedScanDecoupledLookbackDispatcher.hpp"
#include "CubDispatcher.hpp"
#include "EmulatedDeadlockingDispatcher.hpp"
#include "ReduceThenScanDispatcher.hpp"

int main(int argc, char** argv)
{
    bool quick = false;
    int algo = GPS_ALGO_CSDL;
    for (int i = 1; i < argc; ++i) {
        if (!strcmp(argv[i], "--quick")) quick = true;
        else if (!strcmp(argv[i], "--algo") && i + 1 < argc) {
            ++i;
            algo = !strcmp(argv[i], "reg") ? GPS_ALGO_CSDL_REG : !strcmp(argv[i], "tma") ? GPS_ALGO_CSDL_TMA : GPS_ALGO_CSDL;
        }
    }
    bool ok = true;
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
