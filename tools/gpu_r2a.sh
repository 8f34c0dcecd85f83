#!/bin/bash
# Round 2, first GPU call: the new tests (drop-in against the staged reference, named-config slices), then the new bench line.
set -u
mkdir -p gpurun_out
TAG=${1:-R2a}
nvidia-smi --query-gpu=name,memory.total --format=csv,noheader | head -2
python -m pytest tests/test_gpu_dropin.py tests/test_gpu_named_configs.py -m gpu -x -q -s 2>&1 | tail -25
python bench.py --workload cfg2 --steps 20 --warmup 3 > gpurun_out/bench_cfg2_$TAG.json 2> gpurun_out/bench_cfg2_$TAG.err || tail -30 gpurun_out/bench_cfg2_$TAG.err
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_cfg3_$TAG.json 2> gpurun_out/bench_cfg3_$TAG.err || tail -30 gpurun_out/bench_cfg3_$TAG.err
python - <<PY
import json
for c in ('cfg2','cfg3'):
    try:
        d=json.load(open('gpurun_out/bench_%s_$TAG.json' % c)); r=d['roofline']
        print(c, '%.3e upd/s %.2f ms' % (d['value'], d['ms_per_step']), {k:round(v,2) for k,v in r['phases_ms'].items()}, 'dom frac %.3f whole %.3f' % (r['frac'], r['whole_pass']['frac']))
        print('   e2e(TreeAnc) %.2f ms  dense %.2f ms  cpu %.2e' % (d['e2e']['ms_per_step'], d.get('e2e_cabi_dense',{}).get('ms_per_step',-1), d['cpu_baseline']['value']), d['e2e']['rel_lh_diff_vs_resident_pass'])
        print('   parity', {k:d['parity'][k] for k in ('log_lh_rel_err','max_profile_abs_err','argmax_mismatch','argmax_mismatch_off_ties','patterns_compared')})
    except Exception as e:
        print(c, 'FAILED', e)
PY
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_cfg3_reference_$TAG.json 2> gpurun_out/bench_cfg3_reference_$TAG.err
tail -c 700 gpurun_out/bench_cfg3_reference_$TAG.json
