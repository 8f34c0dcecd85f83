#!/bin/bash
for wv in 2 4 8; do
  for mg in 48 128 256; do
    TTB_SS_WAVES=$wv TTB_MAX_GROUP=$mg python bench.py --workload cfg5 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read());print('waves=$wv mg=$mg cfg5 %.3f ms' % d['ms_per_step'])"
  done
done
