#!/bin/bash
# First GPU pass: sanity at small size, tuning sweep at 2^28, parity tests, reference
# baselines, golden fixtures, bench.  Every step has its own timeout; nothing aborts the rest.
set +e
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks.mem,power.draw,memory.total --format=csv > gpurun_out/gpu_info.csv 2>&1
T=gpuprefixsums_b200/bin/gps_tune
echo "== tune 2^20 (sanity)"; timeout 120 $T --log2n 20 --iters 5 > gpurun_out/tune_2p20.jsonl 2> gpurun_out/tune_2p20.err; echo "rc=$?"
tail -3 gpurun_out/tune_2p20.err
echo "== tune ragged (sanity)"; timeout 120 $T --n 1234567 --iters 3 --filter csdl > gpurun_out/tune_ragged.jsonl 2> gpurun_out/tune_ragged.err; echo "rc=$?"
echo "== tune 2^28"; timeout 900 $T --log2n 28 --iters 30 > gpurun_out/tune_2p28.jsonl 2> gpurun_out/tune_2p28.err; echo "rc=$?"
tail -3 gpurun_out/tune_2p28.err
echo "== golden"; timeout 300 python oracle/gen_golden.py gpurun_out/golden > gpurun_out/golden.log 2>&1; echo "rc=$?"
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"
tail -5 gpurun_out/pytest_gpu.log
echo "== reference main"; timeout 600 oracle/_ref/ref_scan > gpurun_out/ref_scan.log 2>&1; echo "rc=$?"
grep -E "passed|failed|speed" gpurun_out/ref_scan.log
echo "== gps_main --quick"; timeout 300 gpuprefixsums_b200/bin/gps_main --quick > gpurun_out/gps_main_quick.log 2>&1; echo "rc=$?"
grep -E "passed|failed|speed|stats" gpurun_out/gps_main_quick.log
echo "== bench"; timeout 600 python bench.py > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "rc=$?"
cat gpurun_out/bench.log; tail -3 gpurun_out/bench.err
grep -h gkeys gpurun_out/tune_2p28.jsonl | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l)
    print('%-36s %8.2f Gk/s  inplace %8.2f  errs %d wd %d' % (d['variant'], d['gkeys_s'], d['gkeys_s_inplace'], d['errors'], d['watchdog']))
"
