#!/bin/bash
# ncu evidence for profiles/: launch list of the bench command + full captures of the scan kernels
set +e
mkdir -p gpurun_out
echo "== launch list"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 20 --warmup 3 --no-cpu-baseline --e2e-steps 2 > gpurun_out/bench_under_ncu.log 2>&1; echo "rc=$?"
echo "== full capture csdl_blocked (default)"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:csdl_blocked_kernel -s 5 -c 2 -f -o gpurun_out/prof_csdl_blocked python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_csdl_blocked.log 2>&1; echo "rc=$?"
echo "== full capture csdl_smem"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:csdl_smem_kernel -s 5 -c 1 -f -o gpurun_out/prof_csdl_smem python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 1 --algo csdl_smem > gpurun_out/ncu_csdl_smem.log 2>&1; echo "rc=$?"
echo "== full capture smem fallback"; timeout 900 ncu --set full --clock-control none -k regex:csdl_smem_fallback_kernel -s 5 -c 1 -f -o gpurun_out/prof_csdl_smem_fallback python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 1 --algo csdl_fallback > gpurun_out/ncu_fb.log 2>&1; echo "rc=$?"
echo "== clean bench (not under ncu) for the share comparison"; timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_clean.log 2>&1; cat gpurun_out/bench_clean.log | cut -c1-400
ls -la gpurun_out/*.ncu-rep gpurun_out/launches.csv
