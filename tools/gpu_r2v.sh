#!/bin/bash
set -u
mkdir -p gpurun_out
TAG=${1:-R2v}
timeout 1700 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for W in cfg3 cfg2; do
  python bench.py --workload $W --steps 20 --warmup 3 --no-e2e --no-precision-study --cpu-patterns 200 > gpurun_out/bench_${W}_$TAG.json 2> gpurun_out/bench_${W}_$TAG.err || tail -5 gpurun_out/bench_${W}_$TAG.err
  python - <<PY
import json
d=json.load(open('gpurun_out/bench_${W}_$TAG.json')); r=d['roofline']
print('$W %.4f ms  %.3e upd/s dom %.3f whole %.3f' % (d['ms_per_step'], d['value'], r['frac'], r['whole_pass']['frac']), {k:round(v,3) for k,v in r['phases_ms'].items()}, d.get('parity',{}).get('log_lh_rel_err'), d.get('parity',{}).get('max_profile_abs_err'), d.get('parity',{}).get('argmax_mismatch_off_ties'))
PY
done
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_write.sum --clock-control none -k regex:post_leaf -c 3 --csv \
    --log-file gpurun_out/leaf_cfg3_$TAG.csv python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-precision-study > gpurun_out/ncu_leaf_$TAG.log 2>&1
grep -E "gpu__time" gpurun_out/leaf_cfg3_$TAG.csv | cut -d, -f5,9,13-15 | head -4
