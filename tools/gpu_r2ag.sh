#!/bin/bash
set -u
mkdir -p gpurun_out
for MODE in new old new old; do
  if [ $MODE = old ]; then export TTB_TWO_COLLECTIVES=1; else unset TTB_TWO_COLLECTIVES; fi
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 --no-secondary --no-cpu-baseline > gpurun_out/tmp_$MODE.json 2> gpurun_out/tmp_$MODE.err
  python -c "
import json
d=json.load(open('gpurun_out/tmp_$MODE.json')); print('$MODE', 'pass %.3f ms  e2e %.2f ms' % (d['ms_per_step'], d['e2e']['ms_per_step']))"
done
