#!/bin/bash
# round 2: ncu evidence for profiles/ -- launch list of the bench command, full captures of the default
# scan kernel and of one rank's sharded step (reduce + fold + scan), summaries as JSON
set +e
mkdir -p gpurun_out
echo "== launch list"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 20 --warmup 3 --no-cpu-baseline --e2e-steps 2 > gpurun_out/r02_bench_under_ncu.log 2>&1; echo "rc=$?"
echo "== full capture csdl_blocked (default)"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:csdl_blocked_kernel -s 5 -c 2 -f -o gpurun_out/r02_prof_csdl_blocked python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/r02_ncu_csdl_blocked.log 2>&1; echo "rc=$?"
echo "== full capture sharded pair"; timeout 900 ncu --set full --clock-control none -k regex:'rts_reduce_kernel|rts_fold_kernel|csdl_blocked_kernel' -s 9 -c 6 -f -o gpurun_out/r02_prof_sharded_pair python tools/sharded_pair_one_gpu.py 28 8 > gpurun_out/r02_ncu_sharded_pair.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/r02_ncu_sharded_pair.log
echo "== full capture wide fallback (struct scan)"; timeout 900 ncu --set full --clock-control none -k regex:wide_scan_fallback_kernel -s 3 -c 1 -f -o gpurun_out/r02_prof_wide_fallback python -c "
import torch, sys
sys.path.insert(0, '.')
import gpuprefixsums_b200 as gps
ps = gps.PrefixSum(1 << 28)
t = torch.ones((1 << 26, 4), dtype=torch.int32, device='cuda'); o = torch.empty_like(t)
for _ in range(6): ps.scan_wide(t, o, algo='csdl_fallback', struct4=True)
ps.check()
" > gpurun_out/r02_ncu_wide_fb.log 2>&1; echo "rc=$?"
python tools/ncu_summary.py gpurun_out/r02_ncu_summary.json gpurun_out/r02_prof_csdl_blocked.ncu-rep gpurun_out/r02_prof_sharded_pair.ncu-rep gpurun_out/r02_prof_wide_fallback.ncu-rep > /dev/null
ls -la gpurun_out/r02_*
echo "== clean bench (not under ncu)"; timeout 600 python bench.py --steps 200 > gpurun_out/r02_bench_1gpu.json 2> gpurun_out/r02_bench_1gpu.err; cut -c1-600 gpurun_out/r02_bench_1gpu.json
echo "== sharded tests (relaxed slot protocol)"; timeout 600 python -m pytest tests/test_sharded_gpu.py -q -x 2>&1 | tail -3
