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
for TB in 148 222 296 444 888; do for MG in 32 96; do
  TTB_TARGET_BLOCKS=$TB TTB_MAX_GROUP=$MG run "post8 pre16 TB=$TB MG=$MG"
done; done
TTB_MMA_NW_POST=16 TTB_TARGET_BLOCKS=296 run "post16 pre16 TB=296"
TTB_MMA_NW_PRE=8 TTB_TARGET_BLOCKS=296 run "post8 pre8 TB=296"
