#!/bin/bash
# Runs on the GPU box under gpurun: tests, bench lines, ncu launch list and full captures.
# Everything lands in gpurun_out/ (merged back by gpurun).
set -u
mkdir -p gpurun_out
TAG=${1:-r1}
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --workload cfg2 --steps 20 --warmup 3 > gpurun_out/bench_cfg2_$TAG.json 2> gpurun_out/bench_cfg2_$TAG.err
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_cfg3_$TAG.json 2> gpurun_out/bench_cfg3_$TAG.err
python - <<PY
import json
for c in ('cfg2','cfg3'):
    d=json.load(open('gpurun_out/bench_%s_$TAG.json' % c)); r=d['roofline']
    print(c, '%.3e upd/s %.2f ms' % (d['value'], d['ms_per_step']), {k:round(v,2) for k,v in r['phases_ms'].items()}, 'dom frac %.3f whole %.3f e2e %.1f ms cpu %.2e' % (r['frac'], r['whole_pass']['frac'], d['e2e']['ms_per_step'], d['cpu_baseline']['value']))
PY
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_cfg3_reference_$TAG.json 2> gpurun_out/bench_cfg3_reference_$TAG.err
tail -c 400 gpurun_out/bench_cfg3_reference_$TAG.json
# launch list of one bench command (cold-cache, serialised: compare shares only) + DRAM bytes of every launch
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 75 -c 160 --csv \
    --log-file gpurun_out/launches_cfg3_$TAG.csv python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_launch_$TAG.log 2>&1
# full capture of the two level kernels at the bench configuration (one mid-tree level each)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pre_level -s 18 -c 1 -o gpurun_out/pre_cfg3_$TAG -f \
    python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_pre_$TAG.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:post_level_kernel -s 2 -c 1 -o gpurun_out/post_cfg3_$TAG -f \
    python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_post_$TAG.log 2>&1
ls -la gpurun_out | tail -14
