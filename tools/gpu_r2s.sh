#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "aa or cfg4 or golden or 22 or tips or alphabet" 2>&1 | tail -5
run() {
  timeout 300 python bench.py --workload cfg4 --steps 20 --warmup 3 --no-e2e --no-precision-study --cpu-patterns 128 > gpurun_out/tmp.json 2> gpurun_out/tmp.err
  python - "$1" <<PY
import json,sys
try:
    d=json.load(open('gpurun_out/tmp.json')); r=d['roofline']
    print(sys.argv[1], '%.4f ms' % d['ms_per_step'], {k:round(v,3) for k,v in r['phases_ms'].items()}, d.get('parity',{}).get('log_lh_rel_err'), d.get('parity',{}).get('max_profile_abs_err'), d.get('parity',{}).get('argmax_mismatch_off_ties'))
except Exception as e:
    print(sys.argv[1], 'FAILED', e); print(open('gpurun_out/tmp.err').read()[-800:])
PY
}
run "default"
TTB_MMA_NW_PRE=8 TTB_TARGET_BLOCKS=296 run "pre8 TB=296"
TTB_MMA_NW_PRE=8 TTB_TARGET_BLOCKS=888 run "pre8 TB=888"
TTB_MMA_NW_PRE=4 TTB_TARGET_BLOCKS=296 run "pre4 post8 TB=296"
TTB_DBG=2 run "DBG=2"
