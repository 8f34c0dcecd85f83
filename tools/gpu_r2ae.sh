#!/bin/bash
set -u
mkdir -p gpurun_out
TAG=${1:-R2ae}
timeout 1700 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
for W in cfg3 cfg5 cfg4 cfg2; do
  python bench.py --workload $W --steps 10 --warmup 3 > gpurun_out/bench_${W}_$TAG.json 2> gpurun_out/bench_${W}_$TAG.err || tail -20 gpurun_out/bench_${W}_$TAG.err
  python - <<PY
import json
d=json.load(open('gpurun_out/bench_${W}_$TAG.json')); r=d['roofline']; e=d.get('e2e') or {}
print('$W %.3e upd/s %.3f ms dom %.3f whole %.3f e2e %s ms' % (d['value'], d['ms_per_step'], r['frac'], r['whole_pass']['frac'], e.get('ms_per_step')), {k:round(v,3) for k,v in r['phases_ms'].items()}, (d.get('parity') or {}).get('log_lh_rel_err'), (d.get('parity') or {}).get('max_profile_abs_err'), (d.get('parity') or {}).get('argmax_mismatch_off_ties'))
PY
done
