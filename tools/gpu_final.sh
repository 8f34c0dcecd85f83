#!/bin/bash
# Round-end measurement on one B200: parity suite, smoke, bench lines of every BASELINE config that fits one GPU,
# the CPU (reference) arm, ncu launch lists.  Everything lands in gpurun_out/.
set -u
mkdir -p gpurun_out
TAG=${1:-final}
python -m pytest tests -m gpu -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_cfg3_$TAG.json 2> gpurun_out/bench_cfg3_$TAG.err
python bench.py --workload cfg2 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cfg2_$TAG.json 2> gpurun_out/bench_cfg2_$TAG.err
python bench.py --workload cfg4 --steps 20 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_cfg4_$TAG.json 2> gpurun_out/bench_cfg4_$TAG.err
python bench.py --workload cfg5 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_cfg5_$TAG.json 2> gpurun_out/bench_cfg5_$TAG.err
python - <<PY
import json
for c in ('cfg3','cfg2','cfg4','cfg5'):
    try:
        d=json.load(open('gpurun_out/bench_%s_$TAG.json' % c)); r=d['roofline']
        print(c, '%.3e upd/s %.3f ms' % (d['value'], d['ms_per_step']), {k:round(v,2) for k,v in r['phases_ms'].items()}, 'dom %.3f whole %.3f' % (r['frac'], r['whole_pass']['frac']),
              'e2e %.2f ms' % d['e2e']['ms_per_step'] if 'e2e' in d else '', 'cpu %.2e' % d['cpu_baseline']['value'] if d.get('cpu_baseline') else '', d['clocks'])
    except Exception as e:
        print(c, 'FAILED', e)
PY
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_cfg3_reference_$TAG.json 2> gpurun_out/bench_cfg3_reference_$TAG.err
cut -c1-200 gpurun_out/bench_cfg3_reference_$TAG.json; python -c "
import json; d=json.load(open('gpurun_out/bench_cfg3_reference_$TAG.json')); print(d['value'], d['cpu_baseline'])"
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 75 -c 160 --csv \
    --log-file gpurun_out/launches_cfg3_$TAG.csv python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_launch_$TAG.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 90 -c 200 --csv \
    --log-file gpurun_out/launches_cfg5_$TAG.csv python bench.py --workload cfg5 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_launch5_$TAG.log 2>&1
ls -la gpurun_out | grep $TAG
