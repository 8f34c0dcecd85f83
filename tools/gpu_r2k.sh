#!/bin/bash
# where does the time of the q = 20 tensor-pipe level kernels go: arithmetic skipped (TTB_DBG=1), copies skipped (2), both (3),
# and the run length per block (TTB_TARGET_BLOCKS)
set -u
mkdir -p gpurun_out
TAG=${1:-R2k}
run() {
  timeout 300 python bench.py --workload cfg4 --steps 20 --warmup 3 --no-e2e --no-precision-study --no-cpu-baseline > gpurun_out/tmp.json 2> gpurun_out/tmp.err
  python - "$1" <<PY
import json,sys
try:
    d=json.load(open('gpurun_out/tmp.json')); r=d['roofline']
    print(sys.argv[1], '%.4f ms' % d['ms_per_step'], {k:round(v,3) for k,v in r['phases_ms'].items()})
except Exception as e:
    print(sys.argv[1], 'FAILED', e); print(open('gpurun_out/tmp.err').read()[-800:])
PY
}
for NW in 8 16; do
  export TTB_MMA_NW=$NW
  for DBG in 0 1 2 3; do TTB_DBG=$DBG run "NW=$NW DBG=$DBG"; done
  for TB in 296 592 1776 3552; do TTB_TARGET_BLOCKS=$TB run "NW=$NW TARGET_BLOCKS=$TB"; done
done
