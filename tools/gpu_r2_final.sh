#!/bin/bash
# end of round 2: launch list of the bench command with the final code, clean bench lines of both arms for profiles/
set +e
mkdir -p gpurun_out
echo "== launch list"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_bench_final.csv python bench.py --steps 20 --warmup 3 --no-cpu-baseline --e2e-steps 2 > gpurun_out/r02_bench_under_ncu_final.log 2>&1; echo "rc=$?"
python - <<'P'
import csv, collections
rows = list(csv.reader(l for l in open('gpurun_out/r02_launches_bench_final.csv') if l.startswith('"')))
head = rows[0]; ki = head.index('Kernel Name'); vi = head.index('Metric Value')
agg = collections.OrderedDict()
for r in rows[1:]:
    agg.setdefault(r[ki][:90], []).append(float(r[vi].replace(',', '')) / 1000)
tot = sum(sum(v) for v in agg.values())
for k, v in agg.items():
    print('%-92s n %3d  mean %9.2f us  share %5.1f %%' % (k, len(v), sum(v) / len(v), 100 * sum(v) / tot))
P
echo "== bench ours, 200 steps"; timeout 600 python bench.py --steps 200 --warmup 10 > gpurun_out/r02_bench_1gpu_final.json 2> gpurun_out/r02_bench_1gpu_final.err; echo "rc=$?"; cut -c1-400 gpurun_out/r02_bench_1gpu_final.json
echo "== bench reference arm"; timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02_bench_1gpu_reference_arm_final.json 2>/dev/null; echo "rc=$?"; cut -c1-300 gpurun_out/r02_bench_1gpu_reference_arm_final.json
