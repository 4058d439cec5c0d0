#!/bin/bash
# refresh of the launch list of a library sweep 2^10..2^26 with the final dispatch
set +e
mkdir -p gpurun_out
cat > /tmp/tf_driver.py <<'P'
import sys, torch
sys.path.insert(0, '.')
import gpuprefixsums_b200 as gps
ps = gps.PrefixSum(1 << 26)
for n in [1 << lg for lg in range(10, 27)] + [3_800_000, 5_000_000, 6_000_000]:
    x = torch.ones(n, dtype=torch.int32, device='cuda'); o = torch.empty_like(x)
    for _ in range(6): ps.scan(x, o)
    torch.cuda.synchronize()
ps.check()
P
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_sizes.csv python /tmp/tf_driver.py > gpurun_out/r2al.log 2>&1; echo "rc=$?"
