"""Shared helpers for the test-suite: seeded problems in flat-array form."""
import numpy as np
from treetime_b200 import synth
from treetime_b200.gtr import GTR
from treetime_b200.sequence_data import SequenceData
from treetime_b200.flatten import FlatTopology, code_table, gtr_arrays
from treetime_b200 import config as ttconf

NUC_PI = np.array([0.3, 0.2, 0.2, 0.29, 0.01])


def nuc_gtr(mu=1.0):
    return GTR.custom(mu=mu, pi=NUC_PI.copy(), W=np.ones((5, 5)), alphabet='nuc')


def random_gtr(alphabet, seed):
    return GTR.random(alphabet=alphabet, rng=np.random.default_rng(seed))


def make_flat(tree, gtr, L, seed, amb_frac=0.0, amb_chars='N-RY', compress=True, mu_sim=1.0):
    """tree + simulated alignment -> (topo, flat dict, gtr dict)."""
    return synth.make_flat_problem(tree, gtr, L, seed, amb_frac=amb_frac, amb_chars=amb_chars, compress=compress,
                                   mu_sim=mu_sim)


flat_from = synth.flat_problem


def engine_for(flat, g, device=0):
    from treetime_b200.engine import Engine
    eng = Engine(g['Pi'].shape[0], device=device)
    eng.set_tree(flat['parent'], flat['child_ptr'], flat['child_idx'], flat['tip_row'])
    eng.set_patterns(flat['tip_codes'], flat['code_profiles'], flat['multiplicity'])
    eng.set_gtr(g)
    eng.set_branch_lengths(flat['t'])
    return eng


def tie_mask(profile, rel=1e-12):
    """True where the top-2 entries of a profile row are closer than `rel` (argmax ties)."""
    s = np.sort(profile, axis=1)
    return (s[:, -1] - s[:, -2]) <= rel * s[:, -1]


def joint_assignment_lh(flat, g, seq_idx):
    """Per-pattern log-likelihood of a full assignment of internal states under the joint model
    (tips contribute their best compatible state): an assignment is a joint-ML optimum iff this
    equals tree.sequence_LH of the joint pass.  seq_idx[n] for internal nodes, None for tips."""
    import flat_numpy as O
    G = O.make_gtr(g)
    parent = flat['parent']
    L = flat['multiplicity'].shape[0]
    lh = np.log(G.Pi)[seq_idx[0]].astype(float)
    ar = np.arange(L)
    for n in range(1, parent.shape[0]):
        logP = np.log(np.maximum(O.TINY_NUMBER, G.expQt(flat['t'][n])))
        sp = seq_idx[parent[n]]
        if flat['tip_row'][n] >= 0:
            prof = flat['code_profiles'][flat['tip_codes'][flat['tip_row'][n]]]
            msg = np.log(np.maximum(prof, O.TINY_NUMBER))                 # (L, q) over child states i
            lh += (logP[:, sp].T + msg).max(axis=1)
        else:
            lh += logP[seq_idx[n], sp]
    return lh
