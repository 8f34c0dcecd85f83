#!/usr/bin/env python
"""Block timelines written by TTB_TRACE=<file> (ttb_profile_marginal, tensor-pipe preorder kernel): per chunk, how long the
pattern warps waited for the stage, computed while holding it, and finished after releasing it.  Measurement tool only.
Usage: trace_view.py FILE [grid] [NW]"""
import sys
import numpy as np

SLOT = 8 + 17 * 32


def main():
    raw = np.fromfile(sys.argv[1], dtype=np.uint64)
    n = int(min(raw[0], 4096))
    slots = raw[16:16 + n * SLOT].reshape(n, SLOT)
    grids = np.unique(slots[:, 0], return_counts=True)
    print('slots', n, 'grids', dict(zip(grids[0].tolist(), grids[1].tolist())))
    grid = int(sys.argv[2]) if len(sys.argv) > 2 else int(grids[0][np.argmax(grids[1])])
    nw = int(sys.argv[3]) if len(sys.argv) > 3 else 16
    sel = slots[slots[:, 0] == grid]
    ghz = 1.92
    t0g = sel[:, 3].astype(np.int64)
    print('grid %d: %d blocks, start spread %.1f us' % (grid, len(sel), (t0g.max() - t0g.min()) / 1e3))
    waits, holds, tails, totals, startup = [], [], [], [], []
    for b in sel:
        ev = b[8:].reshape(17, 32).astype(np.int64)
        nch = int(b[5])
        c0 = int(b[6])
        w = ev[:nw]
        startup.append((w[:, 2].min() - c0) / ghz / 1e3)          # block start -> first stage ready
        for u in range(min(nch, 9)):
            ready, rel, end = w[:, 2 + 3 * u], w[:, 3 + 3 * u], w[:, 4 + 3 * u]
            prev_end = w[:, 4 + 3 * (u - 1)] if u else w[:, 1]
            waits.append(((ready - prev_end).mean()) / ghz / 1e3)
            holds.append(((rel - ready).mean()) / ghz / 1e3)
            tails.append(((end - rel).mean()) / ghz / 1e3)
        last = min(nch, 9) - 1
        totals.append((w[:, 4 + 3 * last].max() - c0) / ghz / 1e3)
    f = lambda x: 'mean %.2f  p10 %.2f  p50 %.2f  p90 %.2f' % (np.mean(x), np.percentile(x, 10), np.percentile(x, 50), np.percentile(x, 90))
    print('chunks per block: mean %.2f' % sel[:, 5].astype(float).mean())
    print('startup (block start -> first stage ready), us:', f(startup))
    print('per chunk: wait for the stage, us:            ', f(waits))
    print('per chunk: stage held (P1 .. P2), us:         ', f(holds))
    print('per chunk: after release (epilogue), us:      ', f(tails))
    print('block lifetime (first %d chunks), us:           ' % 9, f(totals))
    b = sel[len(sel) // 2]
    ev = b[8:].reshape(17, 32).astype(np.int64)
    c0 = int(b[6])
    print('example block %d on SM %d, %d chunks; rows = warps 0, 1, %d (producer); events in us from block start' % (b[1], b[2], int(b[5]), nw))
    for wi in (0, 1, nw):
        print('  w%-2d' % wi, ' '.join('%6.2f' % ((x - c0) / ghz / 1e3) if x else '     -' for x in ev[wi][:20]))


if __name__ == '__main__':
    main()
