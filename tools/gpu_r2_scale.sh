#!/bin/bash
# Round 2: the driver's SCALE step in miniature -- bench.py at N ranks (strong scaling by default, weak + north-star blocks).
set -u
mkdir -p gpurun_out
TAG=${1:-R2s}
N=${2:-2}
EXTRA=${3:-}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 10 --warmup 3 $EXTRA > gpurun_out/scale_${TAG}_n$N.json 2> gpurun_out/scale_${TAG}_n$N.err
python - <<PY
import json
try:
    d=json.load(open('gpurun_out/scale_${TAG}_n$N.json'))
    print('N=$N %s value=%.4e ms/step=%.3f whole %.3f  e2e=%.4e (%.2f ms) lhdiff %.1e' % (d['scaling'], d['value'], d['ms_per_step'], d['roofline']['whole_pass']['frac'], d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e']['rel_lh_diff_vs_resident_pass']))
    print('  parity', {k:d['parity'][k] for k in ('log_lh_rel_err','max_profile_abs_err','argmax_mismatch_off_ties','patterns_compared')})
    w=d.get('weak_scaling')
    if w: print('  weak: value=%.4e ms=%.3f e2e %.2f ms' % (w['value'], w['ms_per_step'], w['e2e']['ms_per_step']))
    n=d.get('north_star_config')
    if n: print('  cfg5: value=%.4e ms=%.3f whole frac %.3f e2e %.2f ms parity %s' % (n['value'], n['ms_per_step'], n['whole_pass_frac_of_hbm_roofline'], n['e2e']['ms_per_step'], {k:n['parity'][k] for k in ('log_lh_rel_err','max_profile_abs_err','argmax_mismatch_off_ties')}))
except Exception as e:
    print('N=$N failed', e); print(open('gpurun_out/scale_${TAG}_n$N.err').read()[-3000:])
PY
