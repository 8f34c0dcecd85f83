#!/bin/bash
set -u
mkdir -p gpurun_out
run() {
  timeout 300 python bench.py --workload cfg2 --steps 30 --warmup 5 --no-e2e --no-precision-study --no-cpu-baseline > gpurun_out/tmp.json 2> gpurun_out/tmp.err
  python - "$1" <<PY
import json,sys
try:
    d=json.load(open('gpurun_out/tmp.json')); r=d['roofline']
    print(sys.argv[1], '%.4f ms' % d['ms_per_step'], {k:round(v,3) for k,v in r['phases_ms'].items()})
except Exception as e:
    print(sys.argv[1], 'FAILED', e); print(open('gpurun_out/tmp.err').read()[-800:])
PY
}
for TB in 296 444 592 888 1332; do for MG in 32 64; do
  TTB_TARGET_BLOCKS=$TB TTB_MAX_GROUP=$MG run "cfg2 TB=$TB MG=$MG"
done; done
