#!/bin/bash
# q >= 20 level kernels on the fp64 tensor pipe: parity tests + A/B bench at cfg4
set -u
mkdir -p gpurun_out
TAG=${1:-R2h}
timeout 900 python -m pytest tests -m gpu -x -q -k "aa or cfg4 or golden or 22 or tips or alphabet" 2>&1 | tail -40
for NW in 0 4 8 16; do
  if [ $NW = 0 ]; then export TTB_NO_MMA=1; else unset TTB_NO_MMA; export TTB_MMA_NW=$NW; fi
  timeout 300 python bench.py --workload cfg4 --steps 20 --warmup 3 --no-e2e --no-precision-study --cpu-patterns 128 > gpurun_out/bench_cfg4_${TAG}_nw$NW.json 2> gpurun_out/bench_cfg4_${TAG}_nw$NW.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/bench_cfg4_${TAG}_nw$NW.json')); r=d['roofline']
    print('NW=$NW %.4f ms  %.3e upd/s  whole %.3f' % (d['ms_per_step'], d['value'], r['whole_pass']['frac']), {k:round(v,3) for k,v in r['phases_ms'].items()}, d.get('parity',{}).get('log_lh_rel_err'), d.get('parity',{}).get('max_profile_abs_err'), d.get('parity',{}).get('argmax_mismatch_off_ties'))
except Exception as e:
    print('NW=$NW FAILED', e); print(open('gpurun_out/bench_cfg4_${TAG}_nw$NW.err').read()[-1500:])
PY
done
