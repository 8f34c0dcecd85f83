#!/bin/bash
set -u
mkdir -p gpurun_out
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
run "default"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 76 -c 75 --csv \
    --log-file gpurun_out/launches_cfg4_R2t.csv python bench.py --workload cfg4 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-precision-study > gpurun_out/ncu_launch_R2t.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/launches_cfg4_R2t.csv')) if len(r)>5]
hdr=rows[0]
ik=hdr.index('Kernel Name'); iv=hdr.index('Metric Value'); ig=hdr.index('Grid Size'); im=hdr.index('Metric Name'); iid=hdr.index('ID')
out=[(int(r[iid]), r[ik][:30], r[ig], float(r[iv])/1000) for r in rows[1:] if r[im]=='gpu__time_duration.sum']
for o in out[:72]: print(o)
PY
