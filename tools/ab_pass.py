#!/usr/bin/env python
"""A/B timing of the resident pass through the C-ABI (TTB_LIB selects the library): python tools/ab_pass.py cfg5 [steps]"""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import bench
from treetime_b200.engine import Engine
name = sys.argv[1] if len(sys.argv) > 1 else 'cfg5'
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
topo, flat, g = bench.make_workload(name, 1)
eng = Engine(g['Pi'].shape[0])
eng.set_tree(flat['parent'], flat['child_ptr'], flat['child_idx'], flat['tip_row'])
eng.set_patterns(flat['tip_codes'], flat['code_profiles'], flat['multiplicity'])
eng.set_gtr(g)
eng.set_branch_lengths(flat['t'])
for _ in range(5):
    eng.marginal()
eng.sync()
best = []
for rep in range(3):
    t0 = time.perf_counter()
    for _ in range(steps):
        eng.marginal()
    eng.sync()
    best.append(round((time.perf_counter() - t0) / steps * 1e3, 4))
ph = eng.profile_marginal()
print(json.dumps({'lib': os.environ.get('TTB_LIB', 'libttb.so'), 'workload': name, 'ms_per_pass': best, 'phases': {k: round(v[0], 3) for k, v in ph.items()}, 'lh': eng.results()[0]}))
