#!/bin/bash
# Round-2 closing evidence on one B200: full GPU test suite, smoke, bench lines of every configuration (+ the q >= 20 A/B),
# the reference arm, the fp64 probe, ncu launch lists (cfg3, cfg4) and full captures of the tensor-pipe kernels.
set -u
mkdir -p gpurun_out
TAG=${1:-R2y}
(cd tools/probe && nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_probe dmma_probe.cu) && ./tools/probe/dmma_probe > gpurun_out/dmma_probe_$TAG.txt 2>&1
timeout 1700 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_cfg3_$TAG.json 2> gpurun_out/bench_cfg3_$TAG.err || tail -20 gpurun_out/bench_cfg3_$TAG.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_cfg3_reference_$TAG.json 2> gpurun_out/bench_cfg3_reference_$TAG.err
for W in cfg2 cfg4 cfg5; do
  python bench.py --workload $W --steps 20 --warmup 3 > gpurun_out/bench_${W}_$TAG.json 2> gpurun_out/bench_${W}_$TAG.err || tail -20 gpurun_out/bench_${W}_$TAG.err
done
TTB_NO_MMA=1 python bench.py --workload cfg4 --steps 20 --warmup 3 --no-e2e --no-cpu-baseline --no-precision-study > gpurun_out/bench_cfg4_${TAG}_no_mma.json 2> gpurun_out/bench_cfg4_${TAG}_no_mma.err
python - <<PY
import json
for c in ('cfg3','cfg2','cfg4','cfg5','cfg4_${TAG}_no_mma'):
    f = 'gpurun_out/bench_%s_$TAG.json' % c if 'no_mma' not in c else 'gpurun_out/bench_%s.json' % c
    try:
        d=json.load(open(f)); r=d['roofline']; e=d.get('e2e') or {}
        print(c, '%.3e upd/s %.3f ms dom %.3f whole %.3f e2e %s ms' % (d['value'], d['ms_per_step'], r['frac'], r['whole_pass']['frac'], e.get('ms_per_step')), {k:round(v,3) for k,v in r['phases_ms'].items()}, (d.get('parity') or {}).get('log_lh_rel_err'), (d.get('parity') or {}).get('max_profile_abs_err'), (d.get('parity') or {}).get('argmax_mismatch_off_ties'))
    except Exception as ex:
        print(c, 'FAILED', ex)
print(open('gpurun_out/bench_cfg3_reference_$TAG.json').read()[:400])
PY
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 76 -c 160 --csv \
    --log-file gpurun_out/launches_cfg3_$TAG.csv python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-precision-study > gpurun_out/ncu_launch_cfg3_$TAG.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 76 -c 80 --csv \
    --log-file gpurun_out/launches_cfg4_$TAG.csv python bench.py --workload cfg4 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-precision-study > gpurun_out/ncu_launch_cfg4_$TAG.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pre_level_mma -s 16 -c 1 -o gpurun_out/pre_cfg4_$TAG -f \
    python bench.py --workload cfg4 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-precision-study > gpurun_out/ncu_pre_$TAG.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:post_level_mma -s 2 -c 1 -o gpurun_out/post_cfg4_$TAG -f \
    python bench.py --workload cfg4 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-precision-study > gpurun_out/ncu_post_$TAG.log 2>&1
ls -la gpurun_out | grep $TAG
