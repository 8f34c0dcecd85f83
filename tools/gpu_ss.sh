#!/bin/bash
# site-specific iteration: parity tests + cfg5 shard bench line
set -u
mkdir -p gpurun_out
TAG=${1:-ss}
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python bench.py --workload cfg5 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_cfg5_$TAG.json 2> gpurun_out/bench_cfg5_$TAG.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_cfg5_$TAG.json')); r=d['roofline']
print('cfg5: %.3e upd/s %.2f ms' % (d['value'], d['ms_per_step']), {k:round(v,2) for k,v in r['phases_ms'].items()}, 'whole %.3f' % r['whole_pass']['frac'])
PY
tail -2 gpurun_out/bench_cfg5_$TAG.err
