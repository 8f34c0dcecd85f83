#!/bin/bash
# Round 2: merged-level launches -- full GPU suite, then A/B of the merge threshold on cfg2 / cfg3 / cfg4 / cfg5.
set -u
mkdir -p gpurun_out
TAG=${1:-R2b}
python -m pytest tests -m gpu -x -q 2>&1 | tail -8
for M in 0 12 24 48 96; do
  for W in cfg2 cfg4; do
    TTB_MERGE_NODES=$M python bench.py --workload $W --steps 30 --warmup 5 --no-e2e --no-cpu-baseline > gpurun_out/bench_${W}_${TAG}_m$M.json 2> gpurun_out/bench_${W}_${TAG}_m$M.err
  done
done
for M in 0 24 48; do
  TTB_MERGE_NODES=$M python bench.py --workload cfg3 --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_cfg3_${TAG}_m$M.json 2> gpurun_out/bench_cfg3_${TAG}_m$M.err
  TTB_MERGE_NODES=$M python bench.py --workload cfg5 --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_cfg5_${TAG}_m$M.json 2> gpurun_out/bench_cfg5_${TAG}_m$M.err
done
python - <<PY
import json, glob
for f in sorted(glob.glob('gpurun_out/bench_*_${TAG}_m*.json')):
    try:
        d=json.load(open(f)); r=d['roofline']
        print(f.split('/')[-1], '%.4f ms  %.3e upd/s  launches %d  whole %.3f' % (d['ms_per_step'], d['value'], d['gpu_launches'], r['whole_pass']['frac']), {k:round(v,3) for k,v in r['phases_ms'].items()})
    except Exception as e:
        print(f, 'FAILED', e); print(open(f.replace('.json','.err')).read()[-800:])
PY
