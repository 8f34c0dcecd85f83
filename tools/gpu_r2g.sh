#!/bin/bash
# round 2, session 3: DMMA probe + cfg4 baseline on the same box
set -u
mkdir -p gpurun_out
TAG=${1:-R2g}
./tools/probe/dmma_probe 2>&1 | tee gpurun_out/dmma_probe_$TAG.txt
python bench.py --workload cfg4 --steps 20 --warmup 3 --no-e2e --no-cpu-baseline --no-precision-study > gpurun_out/bench_cfg4_$TAG.json 2> gpurun_out/bench_cfg4_$TAG.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_cfg4_$TAG.json')); r=d['roofline']
print('%.4f ms  %.3e upd/s  whole %.3f' % (d['ms_per_step'], d['value'], r['whole_pass']['frac']), {k:round(v,3) for k,v in r['phases_ms'].items()})
PY
