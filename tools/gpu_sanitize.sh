#!/bin/bash
# compute-sanitizer over a small end-to-end run (memcheck + racecheck + synccheck)
set -u
mkdir -p gpurun_out
cat > /tmp/san.py <<'PY'
import sys; sys.path[:0]=['.','oracle','tests']
import numpy as np, util, flat_numpy as O
from treetime_b200 import synth
from treetime_b200.gtr import GTRSiteSpecific
def run(gtr, n, L, seed, mean_bl=0.01, **kw):
    tree = synth.random_tree(n, seed=seed, mean_bl=mean_bl, polytomy_frac=0.3)
    topo, flat, g = util.make_flat(tree, gtr, L, seed, amb_frac=0.02, amb_chars=('X' if gtr.n_states > 8 else 'N-RY'), **kw)
    eng = util.engine_for(flat, g)
    eng.marginal(); tot,_ = eng.results()
    eng.marginal(reconstruct_tips=True); eng.results()
    eng.marginal(lh_only=True); eng.results()
    nodes=np.arange(1,flat['parent'].shape[0],dtype=np.int32)
    eng.branch_objective(nodes, np.full(nodes.shape[0],0.01)); eng.branch_hamming(nodes)
    eng.node_array(3,1); eng.all_seq_idx(); eng.mutations()
    if not g['site_specific']:
        eng.mutation_counts()
        eng.joint(); eng.results(); eng.joint(reconstruct_tips=True); eng.results(); eng.all_seq_idx()
        eng.marginal(); eng.results()
    # sampled sequences, per-pattern statistics, per-branch masks (MASK kernels), back to the plain kernels
    L_ = flat['multiplicity'].shape[0]; n_nodes = flat['parent'].shape[0]
    eng.marginal(reconstruct_tips=True, keep_prev=True); eng.results()
    eng.sample_states(np.arange(1, n_nodes, dtype=np.int32), np.random.default_rng(1).random((n_nodes - 1, L_)))
    eng.mutation_counts_per_site()
    M = np.ones((2, L_), dtype=np.uint8); M[0, L_ // 2:] = 0
    eng.set_branch_masks(M, (np.arange(n_nodes) % 3 - 1).astype(np.int32))
    eng.marginal(reconstruct_tips=True); eng.results()
    eng.branch_objective(nodes, np.full(nodes.shape[0],0.01)); eng.mutation_counts_per_site()
    eng.set_branch_masks(None, None)
    eng.marginal(); eng.results()
    res = O.marginal(flat, g)
    assert abs(tot-res.total_LH) < 1e-9*abs(res.total_LH)
    print('ok', g['Pi'].shape, tot)
run(util.nuc_gtr(), 40, 300, 1)
run(util.random_gtr('aa_nogap', 3), 20, 150, 2)
run(GTRSiteSpecific.random(L=200, alphabet='nuc', rng=np.random.default_rng(5)), 30, 200, 3, compress=False)
run(GTRSiteSpecific.random(L=130, alphabet='nuc', rng=np.random.default_rng(6)), 24, 130, 4, mean_bl=3.0, compress=False)   # long branches: exact exp path
PY
for tool in memcheck racecheck synccheck; do
  echo "== $tool"
  EXTRA=""; if [ $tool = racecheck ]; then EXTRA="--racecheck-report all"; fi
  timeout 900 compute-sanitizer --tool $tool $EXTRA --error-exitcode 3 python /tmp/san.py > gpurun_out/sanitizer_$tool.log 2>&1
  echo "exit $?"; tail -4 gpurun_out/sanitizer_$tool.log
done
