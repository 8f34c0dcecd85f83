#!/bin/bash
# symmetric site-specific kernels: full parity suite + cfg5 shard A/B (symmetric vs general kernels)
set -u
mkdir -p gpurun_out
TAG=${1:-sym}
python -m pytest tests -m gpu -x -q 2>&1 | tail -6
for sym in 1 0; do
  TTB_SS_SYM=$sym python bench.py --workload cfg5 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_cfg5_${TAG}_sym$sym.json 2> gpurun_out/bench_cfg5_${TAG}_sym$sym.err
  python - <<PY
import json
d=json.load(open('gpurun_out/bench_cfg5_${TAG}_sym$sym.json')); r=d['roofline']
print('cfg5 sym=$sym: %.3e upd/s %.2f ms' % (d['value'], d['ms_per_step']), {k:round(v,2) for k,v in r['phases_ms'].items()}, 'whole %.3f' % r['whole_pass']['frac'], 'LH', d['log_lh_rank0'])
PY
  tail -2 gpurun_out/bench_cfg5_${TAG}_sym$sym.err
done
