#!/bin/bash
set -u
mkdir -p gpurun_out
TAG=${1:-R2x}
python tools/e2e_profile.py 2>&1 | head -24
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_cfg3_$TAG.json 2> gpurun_out/bench_cfg3_$TAG.err || tail -20 gpurun_out/bench_cfg3_$TAG.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_cfg3_$TAG.json')); r=d['roofline']
print('cfg3 %.3e upd/s %.3f ms dom %.3f whole %.3f e2e %.2f ms (%.3e)' % (d['value'], d['ms_per_step'], r['frac'], r['whole_pass']['frac'], d['e2e']['ms_per_step'], d['e2e']['value']), {k:round(v,3) for k,v in r['phases_ms'].items()}, d['parity']['log_lh_rel_err'], d['parity']['max_profile_abs_err'], d['parity']['argmax_mismatch_off_ties'])
PY
