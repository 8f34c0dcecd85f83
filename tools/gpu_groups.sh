#!/bin/bash
# sweep of the level-grouping knobs (blocks per level launch / nodes per block)
for tb in 4736 1776 888 444; do
  for mg in 32 64; do
    for w in cfg2 cfg3 cfg4; do
      TTB_TARGET_BLOCKS=$tb TTB_MAX_GROUP=$mg python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read());print('tb=$tb mg=$mg $w %.3f ms' % d['ms_per_step'])"
    done
  done
done
