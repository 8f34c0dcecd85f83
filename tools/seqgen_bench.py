#!/usr/bin/env python
"""N4 measurement: ttb_seqgen throughput (site x branch draws per second) at the bench shapes.
Reference: treetime.seqgen.SeqGen evolves ~1.3e6 site-branches/s."""
import json
import sys
import time

import numpy as np

sys.path.insert(0, '.')
from treetime_b200 import synth                                         # noqa: E402
from treetime_b200.gtr import GTR, GTRSiteSpecific                      # noqa: E402
from treetime_b200.seqgen import SeqGen                                 # noqa: E402

out = []
for name, n_tips, L, ss in (('cfg2', 2000, 10000, False), ('cfg3', 20000, 29903, False), ('cfg5 shard', 100000, 3750, True)):
    tree = synth.random_tree(n_tips, seed=1, mean_bl=1.0 / L)
    if ss:
        gtr = GTRSiteSpecific.random(L=L, alphabet='nuc', rng=np.random.default_rng(1))
    else:
        gtr = GTR.custom(pi=np.array([0.3, 0.2, 0.2, 0.29, 0.01]), W=np.ones((5, 5)), alphabet='nuc')
    sg = SeqGen(L, tree=tree, gtr=gtr, rng_seed=1)
    eng, s2c = sg._prepare()
    eng.seqgen(1, s2c, return_states=False)                              # warm-up
    t0 = time.time()
    reps = 3
    for r in range(reps):
        eng.seqgen(2 + r, s2c, return_states=False)                      # synchronous call
    dt = (time.time() - t0) / reps
    draws = (2 * n_tips - 2) * L
    out.append({'workload': name, 'n_tips': n_tips, 'sites': L, 'site_specific': ss, 'ms': dt * 1e3, 'site_branches_per_s': draws / dt})
    print('%-10s %7d tips x %6d sites: %8.2f ms  %.3e site-branches/s' % (name, n_tips, L, dt * 1e3, draws / dt), file=sys.stderr)
print(json.dumps({'metric': 'seqgen site x branch draws/s', 'reference_cpu': 1.3e6, 'runs': out}))
