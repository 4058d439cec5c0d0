#!/bin/bash
set +e
mkdir -p gpurun_out
T=gpuprefixsums_b200/bin/gps_tune
echo "== experiments"; timeout 300 $T --log2n 28 --iters 30 --filter exp > gpurun_out/tune_exp.jsonl 2> gpurun_out/tune_exp.err; echo "rc=$?"
grep -h gkeys gpurun_out/tune_exp.jsonl | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l)
    print('%-44s %8.2f Gk/s  inplace %8.2f  errs %d wd %d' % (d['variant'], d['gkeys_s'], d['gkeys_s_inplace'], d['errors'], d['watchdog']))
"
echo "== reference kernels on random data"; timeout 300 python tools/run_ref_kernel.py 28 20 3 > gpurun_out/ref_kernels.json 2> gpurun_out/ref_kernels.err; echo "rc=$?"; cat gpurun_out/ref_kernels.json
echo "== ncu mine"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:csdl_reg_kernel -s 4 -c 1 -f -o gpurun_out/prof_reg $T --log2n 28 --iters 2 --filter "csdl_reg<256,4>" > gpurun_out/ncu_reg.log 2>&1; echo "rc=$?"
echo "== ncu ref"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:CSDLInclusive -s 2 -c 1 -f -o gpurun_out/prof_refcsdl python tools/run_ref_kernel.py 28 3 3 > gpurun_out/ncu_ref.log 2>&1; echo "rc=$?"
ls -la gpurun_out/*.ncu-rep
