#!/bin/bash
set -u
mkdir -p gpurun_out
TAG=${1:-R2c}
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -s -k "float_message or nuc_binary or site_specific_gtr" 2>&1 | tail -12
python bench.py --workload cfg2 --steps 20 --warmup 3 --no-e2e > gpurun_out/bench_cfg2_$TAG.json 2> gpurun_out/bench_cfg2_$TAG.err || tail -30 gpurun_out/bench_cfg2_$TAG.err
python bench.py --steps 10 --warmup 3 --no-e2e > gpurun_out/bench_cfg3_$TAG.json 2> gpurun_out/bench_cfg3_$TAG.err || tail -30 gpurun_out/bench_cfg3_$TAG.err
python bench.py --workload cfg5 --steps 10 --warmup 3 --no-e2e > gpurun_out/bench_cfg5_$TAG.json 2> gpurun_out/bench_cfg5_$TAG.err || tail -30 gpurun_out/bench_cfg5_$TAG.err
python - <<PY
import json
for c in ('cfg2','cfg3','cfg5'):
    try:
        d=json.load(open('gpurun_out/bench_%s_$TAG.json' % c)); r=d['roofline']
        print(c, '%.3e upd/s %.3f ms whole %.3f' % (d['value'], d['ms_per_step'], r['whole_pass']['frac']))
        print('   ', json.dumps(d.get('precision_study')))
    except Exception as e:
        print(c, 'FAILED', e)
PY
