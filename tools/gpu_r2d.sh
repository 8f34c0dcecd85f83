#!/bin/bash
set -u
mkdir -p gpurun_out
TAG=${1:-R2d}
python -m pytest tests -m gpu -x -q 2>&1 | tail -6
for W in cfg2 cfg3; do
  TTB_NO_PAIR_TABLES=1 python bench.py --workload $W --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_${W}_${TAG}_nopairs.json 2> gpurun_out/bench_${W}_${TAG}_nopairs.err
  python bench.py --workload $W --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_${W}_${TAG}_pairs.json 2> gpurun_out/bench_${W}_${TAG}_pairs.err
done
python - <<PY
import json, glob
for f in sorted(glob.glob('gpurun_out/bench_*_${TAG}_*pairs.json')):
    try:
        d=json.load(open(f)); r=d['roofline']
        print(f.split('/')[-1], '%.4f ms  %.3e upd/s  whole %.3f' % (d['ms_per_step'], d['value'], r['whole_pass']['frac']), {k:round(v,3) for k,v in r['phases_ms'].items()})
    except Exception as e:
        print(f, 'FAILED', e); print(open(f.replace('.json','.err')).read()[-800:])
PY
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 75 -c 80 --csv \
    --log-file gpurun_out/launches_cfg3_$TAG.csv python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_launch_$TAG.log 2>&1
grep -E "leaf" gpurun_out/launches_cfg3_$TAG.csv | grep gpu__time | head -4
