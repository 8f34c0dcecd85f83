#!/usr/bin/env python
"""Instruction mix and stall samples per SASS opcode from an .ncu-rep captured with --import-source on
(ncu -i REP --page source --csv).  Usage: ncu_opmix.py REP [top]"""
import collections
import csv
import subprocess
import sys


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
    out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == 'Address')
    h = rows[hi]
    isrc, iex, ismp = h.index('Source'), h.index('Instructions Executed'), h.index('# Samples')
    ex, smp = collections.Counter(), collections.Counter()
    for r in rows[hi + 1:]:
        if len(r) <= iex:
            continue
        toks = r[isrc].split()
        if not toks:
            continue
        op = toks[1] if toks[0].startswith('@') and len(toks) > 1 else toks[0]
        op = op.rstrip(';')
        ex[op] += int(r[iex] or 0)
        smp[op] += int(r[ismp] or 0)
    tot, tots = sum(ex.values()), sum(smp.values())
    print('# %s: %d warp instructions, %d stall samples' % (rep, tot, tots))
    for op, n in ex.most_common(top):
        print('%-28s %12d  %5.1f %%   samples %5.1f %%' % (op, n, 100.0 * n / tot, 100.0 * smp[op] / max(1, tots)))


if __name__ == '__main__':
    main()
