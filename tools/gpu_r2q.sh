#!/bin/bash
# full GPU test suite + cfg4 evidence for the tensor-pipe kernels (bench line, launch list, captures)
set -u
mkdir -p gpurun_out
TAG=${1:-R2q}
timeout 1700 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python bench.py --workload cfg4 --steps 20 --warmup 3 > gpurun_out/bench_cfg4_$TAG.json 2> gpurun_out/bench_cfg4_$TAG.err || tail -20 gpurun_out/bench_cfg4_$TAG.err
TTB_NO_MMA=1 python bench.py --workload cfg4 --steps 20 --warmup 3 --no-e2e --no-cpu-baseline --no-precision-study > gpurun_out/bench_cfg4_${TAG}_no_mma.json 2> gpurun_out/bench_cfg4_${TAG}_no_mma.err
python - <<PY
import json
for f in ('gpurun_out/bench_cfg4_$TAG.json','gpurun_out/bench_cfg4_${TAG}_no_mma.json'):
    try:
        d=json.load(open(f)); r=d['roofline']
        print(f, '%.4f ms  %.3e upd/s dom %.3f whole %.3f' % (d['ms_per_step'], d['value'], r['frac'], r['whole_pass']['frac']), {k:round(v,3) for k,v in r['phases_ms'].items()}, d.get('parity',{}).get('log_lh_rel_err'), d.get('parity',{}).get('max_profile_abs_err'), d.get('parity',{}).get('argmax_mismatch_off_ties'))
    except Exception as e:
        print(f, 'FAILED', e)
PY
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 76 -c 80 --csv \
    --log-file gpurun_out/launches_cfg4_$TAG.csv python bench.py --workload cfg4 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-precision-study > gpurun_out/ncu_launch_$TAG.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pre_level_mma -s 16 -c 1 -o gpurun_out/pre_cfg4_$TAG -f \
    python bench.py --workload cfg4 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-precision-study > gpurun_out/ncu_pre_$TAG.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:post_level_mma -s 2 -c 1 -o gpurun_out/post_cfg4_$TAG -f \
    python bench.py --workload cfg4 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-precision-study > gpurun_out/ncu_post_$TAG.log 2>&1
ls -la gpurun_out | grep $TAG
