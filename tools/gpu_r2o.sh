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
for NW in 4 8 16; do
  export TTB_MMA_NW=$NW
  run "NW=$NW"
  TTB_TARGET_BLOCKS=296 run "NW=$NW TARGET_BLOCKS=296"
  TTB_TARGET_BLOCKS=592 run "NW=$NW TARGET_BLOCKS=592"
done
export TTB_MMA_NW=8
TTB_TRACE_GRID=736 TTB_TRACE=gpurun_out/trace.bin timeout 300 python bench.py --workload cfg4 --steps 3 --warmup 3 --no-e2e --no-precision-study --no-cpu-baseline > gpurun_out/tmp.json 2> gpurun_out/tmp.err || tail -5 gpurun_out/tmp.err
python tools/trace_view.py gpurun_out/trace.bin 736 8; rm -f gpurun_out/trace.bin
