#!/bin/bash
# quick GPU iteration: parity tests + cfg3/cfg2 bench lines (no CPU baseline)
set -u
mkdir -p gpurun_out
TAG=${1:-q}
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cfg3_$TAG.json 2> gpurun_out/bench_cfg3_$TAG.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_cfg3_$TAG.json'))
r=d['roofline']
print('cfg3: %.3e upd/s  %.2f ms/step  phases %s  dom %s %.0f GB/s frac %.3f  whole %.3f  e2e %.2f ms' % (d['value'], d['ms_per_step'], {k:round(v,2) for k,v in r['phases_ms'].items()}, r['kernel'][:12], r['achieved'], r['frac'], r['whole_pass']['frac'], d.get('e2e',{}).get('ms_per_step',0)), 'sparse-io e2e %.2f ms' % d.get('e2e_sparse_io',{}).get('ms_per_step',0))
PY
tail -3 gpurun_out/bench_cfg3_$TAG.err
python bench.py --workload cfg2 --steps 20 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_cfg2_$TAG.json 2> gpurun_out/bench_cfg2_$TAG.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_cfg2_$TAG.json'))
r=d['roofline']
print('cfg2: %.3e upd/s  %.3f ms/step  phases %s  whole %.3f' % (d['value'], d['ms_per_step'], {k:round(v,3) for k,v in r['phases_ms'].items()}, r['whole_pass']['frac']))
PY
