#!/bin/bash
# usage: tools/gpu_sweep.sh <tag> [gps_tune args...]  -> gpurun_out/sweep_<tag>.jsonl + table
set +e
mkdir -p gpurun_out
tag=$1; shift
T=gpuprefixsums_b200/bin/gps_tune
timeout 900 $T "$@" > gpurun_out/sweep_$tag.jsonl 2> gpurun_out/sweep_$tag.err; echo "rc=$?"
tail -2 gpurun_out/sweep_$tag.err
python - gpurun_out/sweep_$tag.jsonl <<'PY'
import sys, json
for l in open(sys.argv[1]):
    d = json.loads(l)
    if 'error' in d: print(d); continue
    if 'profile' in d:
        print("   profile: claim %.2f load %.2f reduce %.2f lookback %.2f finish %.2f us | windows %.2f polls %.1f" % (d['us_claim'], d['us_load'], d['us_reduce'], d['us_lookback'], d['us_finish'], d['windows'], d['polls'])); continue
    print("%-44s n=2^%-4.1f %9.2f us %8.2f Gk/s  inplace %8.2f  errs %d wd %d" % (d["variant"], __import__("math").log2(d["n"]), d["ms"]*1e3, d["gkeys_s"], d["gkeys_s_inplace"], d["errors"], d["watchdog"]))
PY
