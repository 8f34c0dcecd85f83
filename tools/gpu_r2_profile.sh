#!/bin/bash
# Round 2 evidence: bench lines of every configuration, the ncu launch list of the default bench command and full
# captures of the dominant kernels.  Everything lands in gpurun_out/ (copied to profiles/ by hand).
set -u
mkdir -p gpurun_out
TAG=${1:-R2f}
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_cfg3_$TAG.json 2> gpurun_out/bench_cfg3_$TAG.err || tail -20 gpurun_out/bench_cfg3_$TAG.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_cfg3_reference_$TAG.json 2> gpurun_out/bench_cfg3_reference_$TAG.err
for W in cfg2 cfg4 cfg5; do
  python bench.py --workload $W --steps 20 --warmup 3 > gpurun_out/bench_${W}_$TAG.json 2> gpurun_out/bench_${W}_$TAG.err || tail -20 gpurun_out/bench_${W}_$TAG.err
done
python - <<PY
import json
for c in ('cfg3','cfg2','cfg4','cfg5'):
    try:
        d=json.load(open('gpurun_out/bench_%s_$TAG.json' % c)); r=d['roofline']
        print(c, '%.3e upd/s %.3f ms dom %.3f whole %.3f e2e %.2f ms' % (d['value'], d['ms_per_step'], r['frac'], r['whole_pass']['frac'], d['e2e']['ms_per_step']), {k:round(v,3) for k,v in r['phases_ms'].items()}, d['parity']['log_lh_rel_err'], d['parity']['max_profile_abs_err'], d['parity']['argmax_mismatch_off_ties'])
    except Exception as e:
        print(c, 'FAILED', e)
print(open('gpurun_out/bench_cfg3_reference_$TAG.json').read()[:300])
PY
# launch list of the bench command (cold-cache, serialised: compare shares only) + DRAM bytes of every launch
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 76 -c 160 --csv \
    --log-file gpurun_out/launches_cfg3_$TAG.csv python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_launch_$TAG.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pre_level -s 18 -c 1 -o gpurun_out/pre_cfg3_$TAG -f \
    python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_pre_$TAG.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:post_leaf_level -s 1 -c 1 -o gpurun_out/leaf_cfg3_$TAG -f \
    python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_leaf_$TAG.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:post_level_kernel -s 2 -c 1 -o gpurun_out/post_cfg3_$TAG -f \
    python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_post_$TAG.log 2>&1
ls -la gpurun_out | grep $TAG
