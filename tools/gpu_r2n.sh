#!/bin/bash
set -u
mkdir -p gpurun_out
for NW in 16 8; do
for DBG in 0 2 14; do
TTB_MMA_NW=$NW TTB_DBG=$DBG TTB_TRACE_GRID=736 TTB_TRACE=gpurun_out/trace_nw${NW}_dbg$DBG.bin timeout 300 python bench.py --workload cfg4 --steps 3 --warmup 3 --no-e2e --no-precision-study --no-cpu-baseline > gpurun_out/tmp.json 2> gpurun_out/tmp.err || tail -5 gpurun_out/tmp.err
echo "== NW=$NW DBG=$DBG"; python tools/trace_view.py gpurun_out/trace_nw${NW}_dbg$DBG.bin 736 $NW; rm -f gpurun_out/trace_nw${NW}_dbg$DBG.bin
done
done
