#!/bin/bash
set -u
mkdir -p gpurun_out
./tools/probe/dmma_probe 2>&1 | grep latency
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
for NW in 16 8; do
  export TTB_MMA_NW=$NW
  for DBG in 2 6 10 14 3; do TTB_DBG=$DBG run "NW=$NW DBG=$DBG"; done
done
