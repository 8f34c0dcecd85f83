#!/usr/bin/env python
"""Per-pass DRAM traffic and time of every kernel from an ncu --csv launch list
(tools/gpu_round.sh) -> profiles/traffic.json (read by bench.py's roofline.traffic)."""
import csv
import json
import sys
from collections import OrderedDict


def main():
    src, workload, out = sys.argv[1], sys.argv[2], sys.argv[3]
    rows = list(csv.reader(open(src)))
    h = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
    H = rows[h]
    ki, mi, vi, ui, idi = H.index('Kernel Name'), H.index('Metric Name'), H.index('Metric Value'), H.index('Metric Unit'), H.index('ID')
    launches = OrderedDict()
    for r in rows[h + 1:]:
        if len(r) <= vi:
            continue
        d = launches.setdefault(r[idi], {'name': r[ki].split('(')[0].replace('void ', '')})
        v = float(r[vi].replace(',', ''))
        u = r[ui]
        if r[mi].startswith('gpu__time'):
            v *= {'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 's': 1e6}.get(u.replace('second', 's').replace('usecond', 'us'), 1.0) if u in ('ns', 'us', 'ms', 's') else (1e-3 if u.startswith('n') else 1.0 if u.startswith('u') else 1e3)
            d['us'] = v
        else:
            scale = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(u, 1.0)
            d[r[mi]] = v * scale
    seq = list(launches.values())
    # one pass = the launches after a finish_kernel up to and including the next one (site-specific passes have no
    # expqt launch to split at)
    ends = [i for i, d in enumerate(seq) if d['name'].startswith('finish_kernel')]
    if len(ends) < 2:
        print('no complete pass in the capture'); return
    one = seq[ends[0] + 1:ends[1] + 1]
    agg = OrderedDict()
    for d in one:
        a = agg.setdefault(d['name'], {'launches': 0, 'us': 0.0, 'dram_bytes': 0.0})
        a['launches'] += 1
        a['us'] += d.get('us', 0.0)
        a['dram_bytes'] += d.get('dram__bytes_read.sum', 0.0) + d.get('dram__bytes_write.sum', 0.0)
    tot = sum(a['us'] for a in agg.values())
    for k, a in agg.items():
        a['share_of_pass'] = a['us'] / tot
    try:
        cur = json.load(open(out))
    except Exception:
        cur = {}
    cur[workload] = {'source': src, 'pass_us_under_ncu': tot, 'kernels': agg}
    json.dump(cur, open(out, 'w'), indent=1)
    for k, a in agg.items():
        print('%-34s launches=%3d  %9.1f us  %6.2f GB  share %.3f' % (k, a['launches'], a['us'], a['dram_bytes'] / 1e9, a['share_of_pass']))


if __name__ == '__main__':
    main()
