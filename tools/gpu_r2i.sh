#!/bin/bash
set -u
mkdir -p gpurun_out
TAG=${1:-R2i}
timeout 900 python -m pytest tests -m gpu -x -q -k "aa or cfg4 or golden or 22 or tips or alphabet" 2>&1 | tail -4
export TTB_MMA_NW=${NW:-16}
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 76 -c 80 --csv \
    --log-file gpurun_out/launches_cfg4_$TAG.csv python bench.py --workload cfg4 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-precision-study > gpurun_out/ncu_launch_$TAG.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pre_level_mma -s 16 -c 1 -o gpurun_out/pre_cfg4_$TAG -f \
    python bench.py --workload cfg4 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-precision-study > gpurun_out/ncu_pre_$TAG.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:post_level_mma -s 2 -c 1 -o gpurun_out/post_cfg4_$TAG -f \
    python bench.py --workload cfg4 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-precision-study > gpurun_out/ncu_post_$TAG.log 2>&1
ls -la gpurun_out | grep $TAG
