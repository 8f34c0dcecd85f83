#!/usr/bin/env python
"""Instruction mix and hottest stall sites from `ncu -i rep --page source --csv` output (file argument)."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, data = rows[1], rows[2:]
iS, iN, iE = hdr.index('Source'), hdr.index('# Samples'), hdr.index('Instructions Executed')
cnt = collections.Counter(int(r[iE]) for r in data if 'DFMA' in r[iS] and int(r[iE]))
unit = cnt.most_common(1)[0][0]
c, st = collections.Counter(), collections.Counter()
for r in data:
    t = r[iS].split()
    op = (t[1] if t[0].startswith('@') else t[0]).split('.')[0]
    c[op] += int(r[iE]); st[op] += int(r[iN])
tot, tots = sum(c.values()), sum(st.values())
print('warp instructions %d, per inner iteration (%d) %.1f' % (tot, unit, tot / unit))
for op, n in c.most_common(int(sys.argv[2]) if len(sys.argv) > 2 else 18):
    print('%-10s %5.1f%% exec %6.1f /iter  stall %5.1f%%' % (op, 100 * n / tot, n / unit, 100 * st[op] / tots))
idx = sorted(range(len(data)), key=lambda i: -int(data[i][iN]))[:22]
for i in sorted(idx):
    print(i, '%.1f%%' % (100 * int(data[i][iN]) / tots), data[i][iE], data[i][iS][:90])
