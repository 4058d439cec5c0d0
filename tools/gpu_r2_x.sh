#!/bin/bash
# ncu evidence for the ticket-free mid-size kernels: launch list of a library sweep 2^10..2^26 (which kernel runs where), full captures at 2^20 and 2^23
set +e
mkdir -p gpurun_out
cat > /tmp/tf_driver.py <<'P'
import sys, torch
sys.path.insert(0, '.')
import gpuprefixsums_b200 as gps
ps = gps.PrefixSum(1 << 26)
sizes = [int(a) for a in sys.argv[1:]] or [1 << lg for lg in range(10, 27)]
for n in sizes:
    x = torch.ones(n, dtype=torch.int32, device='cuda'); o = torch.empty_like(x)
    for _ in range(6): ps.scan(x, o)
    torch.cuda.synchronize()
ps.check()
P
echo "== launch list"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_sizes.csv python /tmp/tf_driver.py > gpurun_out/r2x.log 2>&1; echo "rc=$?"
echo "== full capture 2^20 / 2^23"; timeout 600 ncu --set full --clock-control none -k regex:'csdl_reg_kernel|csdl_smem_fallback_kernel' -s 4 -c 1 -f -o gpurun_out/r02_prof_tf_2p20 python /tmp/tf_driver.py 1048576 >> gpurun_out/r2x.log 2>&1; echo "rc=$?"
timeout 600 ncu --set full --clock-control none -k regex:'csdl_reg_kernel|csdl_smem_fallback_kernel' -s 4 -c 1 -f -o gpurun_out/r02_prof_tf_2p23 python /tmp/tf_driver.py 8388608 >> gpurun_out/r2x.log 2>&1; echo "rc=$?"
python tools/ncu_summary.py gpurun_out/r02_ncu_ticket_free.json gpurun_out/r02_prof_tf_2p20.ncu-rep gpurun_out/r02_prof_tf_2p23.ncu-rep | head -70
python - <<'P'
import csv, collections
rows = list(csv.reader(l for l in open('gpurun_out/r02_launches_sizes.csv') if l.startswith('"')))
head = rows[0]; ki = head.index('Kernel Name'); vi = head.index('Metric Value'); gi = head.index('Grid Size')
seen = collections.OrderedDict()
for r in rows[1:]:
    k = (r[ki][:70], r[gi])
    seen.setdefault(k, []).append(float(r[vi].replace(',', '')))
for (k, g), v in seen.items():
    print('%-72s grid %-12s launches %2d  min %8.2f us' % (k, g, len(v), min(v) / 1000.0))
P
