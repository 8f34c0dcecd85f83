#!/bin/bash
# cfg3 resident-pass bench only (no e2e, no CPU baseline), phases printed
set -u
mkdir -p gpurun_out
TAG=${1:-c3q}
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_cfg3_$TAG.json 2> gpurun_out/bench_cfg3_$TAG.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_cfg3_$TAG.json')); r=d['roofline']
print('cfg3 $TAG: %.3e upd/s %.2f ms' % (d['value'], d['ms_per_step']), {k:round(v,2) for k,v in r['phases_ms'].items()}, 'whole %.3f' % r['whole_pass']['frac'], 'LH', d['log_lh_rank0'])
PY
tail -2 gpurun_out/bench_cfg3_$TAG.err
