#!/bin/bash
set -u
mkdir -p gpurun_out
run() {
  timeout 300 python bench.py --workload cfg4 --steps 20 --warmup 3 --no-e2e --no-precision-study --no-cpu-baseline > gpurun_out/tmp.json 2> gpurun_out/tmp.err
  python - "$1" <<PY
import json,sys
try:
    d=json.load(open('gpurun_out/tmp.json')); r=d['roofline']
    print(sys.argv[1], '%.4f ms' % d['ms_per_step'], {k:round(v,3) for k,v in r['phases_ms'].items()})
except Exception as e:
    print(sys.argv[1], 'FAILED', e); print(open('gpurun_out/tmp.err').read()[-800:])
PY
}
run "default"
TTB_MAX_GROUP=64 run "MG=64"
TTB_MMA_NW_PRE=8 run "pre8"
TTB_MMA_NW_POST=16 run "post16"
TTB_DBG=3 run "DBG=3"
TTB_DBG=2 run "DBG=2"
TTB_DBG=1 run "DBG=1"
