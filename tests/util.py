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
