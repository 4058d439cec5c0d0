#!/bin/bash
set +e
mkdir -p gpurun_out
timeout 900 python tools/config3_bench.py > gpurun_out/r02_config3_float32.json 2> gpurun_out/r02_config3.err; echo "rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/r02_config3_float32.json'))
for k,v in d['rows'].items(): print('%-32s'%k, v)
"; tail -3 gpurun_out/r02_config3.err
