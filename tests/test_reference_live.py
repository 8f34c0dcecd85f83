"""Build container only (needs /root/reference): the oracle and the TreeAnc mirror live against
the UNMODIFIED reference imported through oracle/bioshim."""
import numpy as np
import pytest

import refenv

pytestmark = [pytest.mark.reference, pytest.mark.skipif(not refenv.available(), reason='reference not present')]


def test_oracle_bit_identical_to_reference():
    import validate_against_reference as V
    V.main()


def test_mirror_live_against_reference_including_optimize_and_gtr():
    refenv.activate()
    import oracle_engine
    from treetime import GTR as RG
    from treetime_b200 import synth
    from treetime_b200.gtr import GTR
    from treetime_b200.treeanc import TreeAnc
    pi = np.array([.3, .2, .2, .29, .01])
    T = synth.random_tree(30, seed=21, mean_bl=0.01)
    nwk = T.to_newick()
    g = GTR.custom(pi=pi.copy(), W=np.ones((5, 5)), alphabet='nuc')
    idx = synth.evolve_alignment(T, 300, g.Pi, g.W, seed=21)
    aln = synth.sprinkle_ambiguous({k: g.alphabet[v] for k, v in idx.items()}, 0.02, 'N-R', seed=3)
    rt = refenv.reference_treeanc(nwk, aln, RG.custom(pi=pi.copy(), W=np.ones((5, 5)), alphabet='nuc'), rng_seed=1)
    mt = TreeAnc(tree=nwk, aln=aln, gtr=g, rng_seed=1, engine_factory=oracle_engine.factory)
    assert rt.infer_ancestral_sequences(marginal=True) == mt.infer_ancestral_sequences(marginal=True)
    assert rt.sequence_LH() == mt.sequence_LH()
    rn = {n.name: n for n in rt.tree.find_clades()}
    for n in mt.tree.find_clades():
        r = rn[n.name]
        if not n.is_terminal():
            assert np.array_equal(n.marginal_profile, r.marginal_profile) and (n.cseq == r.cseq).all()
            assert n.mutations == [(a, int(p), d) for a, p, d in r.mutations]
    rt.optimize_tree(branch_length_mode='marginal', max_iter=2, infer_gtr=True, prune_short=True)
    mt.optimize_tree(branch_length_mode='marginal', max_iter=2, infer_gtr=True, prune_short=True)
    assert [n.name for n in rt.tree.find_clades()] == [n.name for n in mt.tree.find_clades()]
    a = np.array([n.branch_length for n in rt.tree.find_clades()]); b = np.array([n.branch_length for n in mt.tree.find_clades()])
    assert np.allclose(a[1:], b[1:], rtol=1e-9, atol=1e-14)
    assert np.isclose(rt.sequence_LH(), mt.sequence_LH(), rtol=1e-12)
    rt.optimize_gtr_rate(); mt.optimize_gtr_rate()
    assert np.isclose(rt.gtr.mu, mt.gtr.mu, rtol=1e-9)
