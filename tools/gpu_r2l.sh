#!/bin/bash
set -u
mkdir -p gpurun_out
TAG=${1:-R2l}
./tools/probe/dmma_probe 2>&1 | grep -E "MIXED|warps/SM 16" | tee gpurun_out/dmma_probe_$TAG.txt
timeout 900 python -m pytest tests -m gpu -x -q -k "aa or cfg4 or golden or 22 or tips or alphabet" 2>&1 | tail -5
run() {
  timeout 300 python bench.py --workload cfg4 --steps 20 --warmup 3 --no-e2e --no-precision-study --cpu-patterns 128 > gpurun_out/tmp.json 2> gpurun_out/tmp.err
  python - "$1" <<PY
import json,sys
try:
    d=json.load(open('gpurun_out/tmp.json')); r=d['roofline']
    print(sys.argv[1], '%.4f ms' % d['ms_per_step'], {k:round(v,3) for k,v in r['phases_ms'].items()}, d.get('parity',{}).get('log_lh_rel_err'), d.get('parity',{}).get('max_profile_abs_err'), d.get('parity',{}).get('argmax_mismatch_off_ties'))
except Exception as e:
    print(sys.argv[1], 'FAILED', e); print(open('gpurun_out/tmp.err').read()[-800:])
PY
}
for NW in 4 8 16; do
  export TTB_MMA_NW=$NW
  for DBG in 0 2; do TTB_DBG=$DBG run "NW=$NW DBG=$DBG"; done
done
for NW in 8 16; do
export TTB_MMA_NW=$NW
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pre_level_mma -s 16 -c 1 -o gpurun_out/pre_cfg4_${TAG}_nw$NW -f \
    python bench.py --workload cfg4 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-precision-study > gpurun_out/ncu_pre_$TAG.log 2>&1
done
