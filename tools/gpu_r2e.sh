#!/bin/bash
set -u
mkdir -p gpurun_out
TAG=${1:-R2e}
python -m pytest tests -m gpu -x -q -k "brent or optimize or dropin" 2>&1 | tail -8
python tools/opt_bench.py --cpu-sites 0 --repeat 3 > gpurun_out/opt_bench_$TAG.json 2> gpurun_out/opt_bench_$TAG.err || tail -20 gpurun_out/opt_bench_$TAG.err
cat gpurun_out/opt_bench_$TAG.json
