"""TEST-ONLY: live comparison oracle/flat_numpy.py  <->  unmodified reference.

Run in the build container (needs /root/reference):
    python oracle/validate_against_reference.py
Builds seeded inputs, runs reference TreeAnc.infer_ancestral_sequences(marginal=True)
and the flat restatement on the flattened problem, and asserts bit-level
agreement of every per-node array, the per-pattern LH, total LH, N_diff,
optimal branch lengths and the GTR-inference counts."""
import os
import sys
import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import refenv  # noqa: E402

refenv.activate()
import flat_numpy as O  # noqa: E402
from treetime_b200 import synth  # noqa: E402
from treetime_b200.flatten import flatten_treeanc  # noqa: E402


def compare(tt, reconstruct_tip_states=False, check_bl=3, check_counts=False, label=''):
    topo, flat, g = flatten_treeanc(tt)
    n1 = tt.infer_ancestral_sequences(marginal=True, reconstruct_tip_states=reconstruct_tip_states, debug=True)
    res = O.marginal(flat, g, reconstruct_tip_states=reconstruct_tip_states)
    assert res.N_diff == n1, (res.N_diff, n1)
    worst = 0.0
    for i, node in enumerate(topo.nodes):
        for ours, name in ((res.subtree_LH[i], 'marginal_subtree_LH'), (res.outgroup_LH[i], 'marginal_outgroup_LH'),
                           (res.profile[i], 'marginal_profile')):
            ref = getattr(node, name, None)
            if ref is None:
                assert ours is None or name == 'marginal_profile', (i, name)
                continue
            d = np.abs(ours - ref).max()
            worst = max(worst, d)
        if res.seq_idx[i] is not None:
            assert (tt.gtr.alphabet[res.seq_idx[i]] == node._cseq).all()
    if g['site_specific']:
        assert np.allclose(res.sequence_LH, tt.tree.sequence_LH, rtol=1e-13, atol=0)
    else:
        assert np.array_equal(res.sequence_LH, tt.tree.sequence_LH)
        assert res.total_LH == tt.tree.total_sequence_LH
    # second call: N_diff vs previous reconstruction
    n2 = tt.infer_ancestral_sequences(marginal=True, reconstruct_tip_states=reconstruct_tip_states)
    res2 = O.marginal(flat, g, reconstruct_tip_states=reconstruct_tip_states, prev_seq_idx=res.seq_idx)
    assert res2.N_diff == n2 == 0, (res2.N_diff, n2)
    for n in list(range(1, topo.n_nodes))[:check_bl]:
        a = tt.optimal_marginal_branch_length(topo.nodes[n])
        b = O.optimal_marginal_branch_length(flat, g, res, n)
        # bit-equal inputs; the only slack is numpy's einsum path on (non-)contiguous
        # per-site eigen-systems, which moves Brent's flat minimum by ~1e-9 relative
        assert a == b or (flat['t'].ndim and g['site_specific'] and abs(a - b) < 1e-7 * max(a, b)), (n, a, b)
    if check_counts:
        n_ija, T_ia = O.mutation_counts(flat, g, res)
        # reference accumulation (treeanc.py:1556-1572)
        q = tt.gtr.n_states
        L = flat['multiplicity'].shape[0]
        rn = np.zeros((q, q, L)); rT = np.zeros((q, L))
        for node in tt.tree.get_nonterminals():
            for c in node:
                ms = np.transpose(tt.get_branch_mutation_matrix(c, full_sequence=False), (1, 2, 0))
                rT += 0.5 * tt._branch_length_to_gtr(c) * ms.sum(axis=0) * tt.data.multiplicity(mask=c.mask)
                rT += 0.5 * tt._branch_length_to_gtr(c) * ms.sum(axis=1) * tt.data.multiplicity(mask=c.mask)
                rn += ms * tt.data.multiplicity(mask=c.mask)
        assert np.array_equal(rn, n_ija) and np.array_equal(rT, T_ia)
    print('%-28s nodes=%5d L\'=%5d q=%2d  max|d(profiles)|=%.1e  total_LH=%.10f  OK' % (
        label, topo.n_nodes, flat['multiplicity'].shape[0], g['Pi'].shape[0], worst, res.total_LH))
    # single-site models: bit-identical.  Site-specific: einsum on contiguous copies of
    # the (q,q,L) eigen-systems rounds differently from the reference's swapaxes views.
    assert worst == 0.0 or (g['site_specific'] and worst < 1e-13), worst


def main():
    from treetime import GTR
    from treetime.gtr_site_specific import GTR_site_specific
    from treetime.seqgen import SeqGen
    from Bio import Phylo
    from io import StringIO

    def seqgen_aln(newick, L, gtr, seed):
        sg = SeqGen(L, tree=Phylo.read(StringIO(newick), 'newick'), gtr=gtr, rng_seed=seed, verbose=0)
        sg.evolve()
        return {r.id: np.array(list(str(r.seq))) for r in sg.get_aln()}

    # 1. nuc, q=5, binary tree, with ambiguity codes
    nwk = synth.random_tree(60, seed=1, mean_bl=0.02).to_newick()
    gtr = GTR.custom(pi=np.array([.3, .2, .2, .29, .01]), W=np.ones((5, 5)), alphabet='nuc')
    aln = synth.sprinkle_ambiguous(seqgen_aln(nwk, 400, gtr, 1), 0.02, 'N-RY', seed=2)
    compare(refenv.reference_treeanc(nwk, aln, gtr, rng_seed=1), check_counts=True, label='nuc q=5 binary')
    compare(refenv.reference_treeanc(nwk, aln, gtr, rng_seed=1), reconstruct_tip_states=True, label='nuc q=5 reconstruct tips')
    # 2. polytomies + zero-length branches
    nwk = synth.random_tree(80, seed=2, mean_bl=0.01, polytomy_frac=0.4, zero_frac=0.2).to_newick()
    aln = seqgen_aln(nwk, 300, gtr, 2)
    compare(refenv.reference_treeanc(nwk, aln, gtr, rng_seed=1), label='nuc polytomies/zero bl')
    # 3. aa, JTT92 (q=20)
    g20 = GTR.standard('JTT92')
    nwk = synth.random_tree(30, seed=3, mean_bl=0.05).to_newick()
    aln = seqgen_aln(nwk, 120, g20, 3)
    compare(refenv.reference_treeanc(nwk, aln, g20, rng_seed=1), label='aa JTT92 q=20')
    # 4. site-specific, interpolated and exact
    for approx in (True, False):
        gs = GTR_site_specific.random(L=150, alphabet='nuc', rng=np.random.default_rng(4))
        gs.approximate = approx
        nwk = synth.random_tree(25, seed=4, mean_bl=0.05).to_newick()
        aln = seqgen_aln(nwk, 150, gs, 4)
        compare(refenv.reference_treeanc(nwk, aln, gs, rng_seed=1, compress=False), check_bl=2,
                label='site-specific approx=%s' % approx)
    print('oracle == reference on all cases')


if __name__ == '__main__':
    main()
