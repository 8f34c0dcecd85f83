#!/bin/bash
# Runs on the GPU box under gpurun: tests, bench lines, ncu launch list and full captures.
# Everything lands in gpurun_out/ (merged back by gpurun).
set -u
mkdir -p gpurun_out
TAG=${1:-r1}
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --workload cfg2 --steps 20 --warmup 3 > gpurun_out/bench_cfg2_$TAG.json 2> gpurun_out/bench_cfg2_$TAG.err
tail -c 1500 gpurun_out/bench_cfg2_$TAG.json
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_cfg3_$TAG.json 2> gpurun_out/bench_cfg3_$TAG.err
tail -c 3000 gpurun_out/bench_cfg3_$TAG.json; tail -5 gpurun_out/bench_cfg3_$TAG.err
# launch list (cold-cache, serialised: compare shares only)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 80 -c 160 --csv --log-file gpurun_out/launches_cfg3_$TAG.csv \
    python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_launch_$TAG.log 2>&1
# full capture of the two level kernels on cfg2 (small footprint => fast replays)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pre_level -s 8 -c 2 -o gpurun_out/pre_cfg2_$TAG -f \
    python bench.py --workload cfg2 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_pre_$TAG.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:post_level -s 0 -c 2 -o gpurun_out/post_cfg2_$TAG -f \
    python bench.py --workload cfg2 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_post_$TAG.log 2>&1
ls -la gpurun_out | tail -12
