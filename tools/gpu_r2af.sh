#!/bin/bash
set -u
mkdir -p gpurun_out
TAG=R2af
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pre_level_kernel -s 20 -c 1 -o gpurun_out/pre_cfg5_$TAG -f \
    python bench.py --workload cfg5 --steps 1 --warmup 2 --no-e2e --no-cpu-baseline --no-precision-study > gpurun_out/ncu_pre_$TAG.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:post_level_kernel -s 3 -c 1 -o gpurun_out/post_cfg5_$TAG -f \
    python bench.py --workload cfg5 --steps 1 --warmup 2 --no-e2e --no-cpu-baseline --no-precision-study > gpurun_out/ncu_post_$TAG.log 2>&1
ls -la gpurun_out | grep $TAG
