#!/bin/bash
set -u
mkdir -p gpurun_out
run() {
  timeout 300 python bench.py --workload $1 --steps 20 --warmup 5 --no-e2e --no-precision-study --cpu-patterns 200 > gpurun_out/tmp.json 2> gpurun_out/tmp.err
  python - "$2" <<PY
import json,sys
try:
    d=json.load(open('gpurun_out/tmp.json')); r=d['roofline']
    print(sys.argv[1], '%.4f ms whole %.3f' % (d['ms_per_step'], r['whole_pass']['frac']), {k:round(v,3) for k,v in r['phases_ms'].items()}, d.get('parity',{}).get('log_lh_rel_err'))
except Exception as e:
    print(sys.argv[1], 'FAILED', e); print(open('gpurun_out/tmp.err').read()[-800:])
PY
}
run cfg2 "cfg2 new default"
TTB_TARGET_BLOCKS=888 run cfg2 "cfg2 TB=888 (no-overflow rule)"
run cfg3 "cfg3 new default"
TTB_TARGET_BLOCKS=888 run cfg3 "cfg3 TB=888 (no-overflow rule)"
run cfg4 "cfg4"
