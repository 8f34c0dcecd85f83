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
    idx = synth.evolve_alignment(tree, L, gtr.Pi if np.ndim(gtr.Pi) == 1 else gtr.Pi.mean(axis=1),
                                 gtr.W, mu=mu_sim, seed=seed)
    aln = {k: gtr.alphabet[v] for k, v in idx.items()}
    if amb_frac:
        aln = synth.sprinkle_ambiguous(aln, amb_frac, amb_chars, seed=seed + 1)
    sd = SequenceData(aln, compress=compress, ambiguous=gtr.ambiguous)
    return flat_from(tree, sd, gtr)


def flat_from(tree, sd, gtr):
    tree.ladderize()
    topo = FlatTopology(tree.root)
    chars, lut, table = code_table(gtr.profile_map, gtr.n_states)
    lut8 = np.full(256, 255, dtype=np.uint8)
    for c, i in lut.items():
        lut8[ord(c)] = i
    codes = np.empty((topo.n_tips, sd.compressed_length), dtype=np.uint8)
    for n in topo.tip_nodes:
        node = topo.nodes[n]
        if node.name in sd.compressed_alignment:
            codes[topo.tip_row[n]] = lut8[sd.compressed_alignment.codes(node.name)]
        else:
            codes[topo.tip_row[n]] = len(chars)
    assert (codes != 255).all()
    one_mutation = 1.0 / sd.full_length
    t = np.array([max(ttconf.MIN_BRANCH_LENGTH * one_mutation, n.branch_length if n.branch_length else 0.0)
                  for n in topo.nodes])
    t[0] = max(ttconf.MIN_BRANCH_LENGTH * one_mutation, 0.001)
    flat = topo.as_dict()
    flat.update(tip_codes=codes, code_profiles=table, multiplicity=sd.multiplicity().copy(), t=t)
    return topo, flat, gtr_arrays(gtr)


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
