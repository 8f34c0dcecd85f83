#!/bin/bash
# ncu full captures of the site-specific level kernels at the cfg5 shard (one postorder, one preorder level)
set -u
mkdir -p gpurun_out
TAG=${1:-ss}
timeout 900 ncu --set full --clock-control none --import-source on -k regex:post_level_kernel -s 3 -c 1 -o gpurun_out/post_cfg5_$TAG -f \
    python bench.py --workload cfg5 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_post5_$TAG.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pre_level -s 25 -c 1 -o gpurun_out/pre_cfg5_$TAG -f \
    python bench.py --workload cfg5 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_pre5_$TAG.log 2>&1
ls -la gpurun_out | grep cfg5_$TAG
