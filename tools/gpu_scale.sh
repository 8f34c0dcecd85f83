#!/bin/bash
# N = 1, 2, 4, 8 weak-scaling runs of bench.py on one box (like the driver's SCALE step)
set -u
mkdir -p gpurun_out
TAG=${1:-scale}
NMAX=${2:-8}
for N in 1 2 4 8; do
  if [ $N -gt $NMAX ]; then break; fi
  if [ $N -eq 1 ]; then
    python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/scale_${TAG}_n$N.json 2> gpurun_out/scale_${TAG}_n$N.err
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/scale_${TAG}_n$N.json 2> gpurun_out/scale_${TAG}_n$N.err
  fi
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/scale_${TAG}_n$N.json'))
    print('N=$N value=%.4e ms/step=%.2f e2e=%.4e (%.1f ms)' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step']))
except Exception as e:
    print('N=$N failed', e); print(open('gpurun_out/scale_${TAG}_n$N.err').read()[-1500:])
PY
done
# BASELINE.json configs[4] in full: 8 shards of 3,750 sites of the 100k-tip site-specific workload
if [ $NMAX -ge 8 ]; then
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29529 bench.py --workload cfg5 --gpus 8 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/scale_${TAG}_cfg5_n8.json 2> gpurun_out/scale_${TAG}_cfg5_n8.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/scale_${TAG}_cfg5_n8.json'))
    print('cfg5 N=8 value=%.4e ms/step=%.2f whole-pass frac %.3f' % (d['value'], d['ms_per_step'], d['roofline']['whole_pass']['frac']))
except Exception as e:
    print('cfg5 N=8 failed', e); print(open('gpurun_out/scale_${TAG}_cfg5_n8.err').read()[-1500:])
PY
fi
