"""CPU: the TreeAnc mirror's host logic (flattening, N_diff rules, lazy node views, lock-step
Brent, damping, GTR inference) with the oracle-backed test engine, checked against golden
vectors from the unmodified reference -- and live against the reference when it is present."""
import numpy as np
import pytest

import golden_util as G
import oracle_engine
from treetime_b200.treeanc import TreeAnc, MissingDataError, UnknownMethodError


def mirror_from_golden(z, **kw):
    return TreeAnc(tree=str(z['newick']), aln=G.alignment(z), gtr=G.model(z), rng_seed=1,
                   engine_factory=oracle_engine.factory, **kw)


def check_against_golden(tt, z, exact=True, prof_tol=0.0, lh_rtol=0.0):
    tips = bool(z['reconstruct_tips'])
    n1 = tt.infer_ancestral_sequences(marginal=True, reconstruct_tip_states=tips)
    assert n1 == int(z['N_diff_first'])
    assert np.array_equal(tt.data.multiplicity(), z['multiplicity'])
    assert np.allclose(tt.tree.sequence_LH, z['sequence_LH'], rtol=max(lh_rtol, 1e-15), atol=1e-12 if not exact else 0)
    assert abs(tt.sequence_LH() - float(z['total_LH'])) <= max(lh_rtol, 1e-15) * abs(float(z['total_LH']))
    nodes = list(tt.tree.find_clades())
    assert [n.name for n in nodes] == [str(s) for s in z['node_names']]
    for i in z['stored_nodes']:
        n = nodes[i]
        assert np.abs(n.marginal_subtree_LH - z['subtree_%d' % i]).max() <= prof_tol
        if i > 0:
            assert np.abs(n.marginal_outgroup_LH - z['outgroup_%d' % i]).max() <= prof_tol
        if 'profile_%d' % i in z.files:
            assert np.abs(n.marginal_profile - z['profile_%d' % i]).max() <= prof_tol
    for n, s in zip(nodes, z['cseq']):
        c = n.cseq
        if c is None:
            assert str(s) == ''
        elif exact:
            assert ''.join(c) == str(s)
        else:
            bad = np.array(list(str(s))) != c
            if bad.any():
                import util
                assert util.tie_mask(n.marginal_profile)[bad].all()
    assert tt.infer_ancestral_sequences(marginal=True, reconstruct_tip_states=tips) == int(z['N_diff_second'])
    return nodes


@pytest.mark.parametrize('name', ['kat3', 'nuc40', 'nuc40_tips', 'poly70', 'aa16_jtt92', 'aa16_q22'])
def test_mirror_reconstruction_matches_reference_golden(name):
    z = G.load(name)
    tt = mirror_from_golden(z)
    nodes = check_against_golden(tt, z)
    for k, i in enumerate(z['bl_nodes']):
        bl = tt.optimal_marginal_branch_length(nodes[i])
        assert abs(bl - z['bl_opt'][k]) <= 1e-9 * z['bl_opt'][k]
        bl2 = tt.optimal_marginal_branch_length(nodes[i], tol=1e-2)
        assert abs(bl2 - z['bl_opt_tol1e-2'][k]) <= 1e-9 * z['bl_opt_tol1e-2'][k]


@pytest.mark.parametrize('name', ['nuc40', 'poly70'])
def test_mirror_optimize_tree_marginal_matches_reference(name):
    """optimize_tree(branch_length_mode='marginal', max_iter=2): batched lock-step Brent + damping +
    the double update across a bifurcating root reproduce the reference's branch lengths."""
    z = G.load(name)
    tt = mirror_from_golden(z)
    tt.optimize_tree(branch_length_mode='marginal', max_iter=2, infer_gtr=False, prune_short=False)
    bl = np.array([n.branch_length for n in tt.tree.find_clades()])
    assert np.allclose(bl[1:], z['opt_branch_length'][1:], rtol=1e-9, atol=1e-12)
    assert abs(tt.sequence_LH() - float(z['opt_total_LH'])) <= 1e-11 * abs(float(z['opt_total_LH']))


def test_mirror_infer_gtr_matches_reference():
    z = G.load('nuc40')
    tt = mirror_from_golden(z)
    tt.optimize_tree(branch_length_mode='marginal', max_iter=2, infer_gtr=False, prune_short=False)
    tt.infer_gtr(marginal=True)
    assert np.allclose(tt.gtr.W, z['inferred_W'], rtol=1e-9) and np.allclose(tt.gtr.Pi, z['inferred_Pi'], rtol=1e-9)
    assert np.isclose(tt.gtr.mu, float(z['inferred_mu']), rtol=1e-12)


def test_mirror_big_config1_regenerated_inputs():
    """BASELINE.json configs[0] at full size: inputs regenerated from seeds (checksum-pinned)."""
    z = G.load('cfg1_200x1400')
    tree, aln, g, sha = G.regenerate_big('cfg1')
    assert sha == str(z['input_sha'])
    tt = TreeAnc(tree=tree, aln=aln, gtr=g, rng_seed=1, engine_factory=oracle_engine.factory)
    assert tt.infer_ancestral_sequences(marginal=True) == int(z['N_diff_first'])
    assert np.allclose(tt.tree.sequence_LH, z['sequence_LH'], rtol=1e-13)
    assert abs(tt.sequence_LH() - float(z['total_LH'])) < 1e-12 * abs(float(z['total_LH']))
    for n, s in zip(tt.tree.find_clades(), z['cseq']):
        if not n.is_terminal():
            assert ''.join(n.cseq) == str(s)


def test_error_behaviour_mirrors_reference():
    z = G.load('nuc40')
    tt = mirror_from_golden(z)
    with pytest.raises(UnknownMethodError):
        tt.infer_ancestral_sequences(method='nonsense', marginal=True)
    with pytest.raises(UnknownMethodError):
        tt.optimize_tree(branch_length_mode='nonsense')
    with pytest.raises(Exception):
        tt.marginal_branch_profile(tt.tree.root)
    with pytest.raises(TypeError):
        TreeAnc(tree=None, aln=G.alignment(z))
    with pytest.raises(MissingDataError):
        TreeAnc(tree='(A:0.1,B:0.2);', aln={'A': 'ACGT', 'B': 'ACGT'}, engine_factory=oracle_engine.factory)
    # a tree whose tips are not in the alignment (treeanc.py:419-430)
    with pytest.raises(MissingDataError):
        TreeAnc(tree='((x:0.1,y:0.1):0.1,(z:0.1,w:0.2):0.1);', aln=G.alignment(z), engine_factory=oracle_engine.factory)
    # sample_from_profile='root' consumes the host RNG like the reference (one uniform per pattern)
    tt.infer_ancestral_sequences(marginal=True, sample_from_profile='root')
    assert tt.tree.root.cseq.shape[0] == tt.data.compressed_length


def test_reconstructed_alignment_and_mutations():
    z = G.load('nuc40')
    tt = mirror_from_golden(z)
    aln = tt.get_reconstructed_alignment()
    assert set(aln) == {str(s) for s in z['node_names']}
    L = tt.data.full_length
    assert all(len(s) == L for s in aln.values())
    for n in tt.tree.find_clades():
        if n.up is not None and not n.is_terminal():
            for a, pos, d in n.mutations:
                assert aln[n.up.name][pos] == a and aln[n.name][pos] == d and a != d


def test_device_compress_path_equals_host_compress():
    """N3 host logic: patterns numbered from device-style column statistics give the same TreeAnc results."""
    z = G.load('nuc40')
    a = mirror_from_golden(z)
    b = mirror_from_golden(z, device_compress=True)
    assert b.data.device_resident and not a.data.device_resident
    assert a.infer_ancestral_sequences(marginal=True) == b.infer_ancestral_sequences(marginal=True) == int(z['N_diff_first'])
    assert a.sequence_LH() == b.sequence_LH() == float(z['total_LH'])
    assert np.array_equal(a.tree.sequence_LH, b.tree.sequence_LH)
    for x, y in zip(a.tree.find_clades(), b.tree.find_clades()):
        if not x.is_terminal():
            assert (x.cseq == y.cseq).all() and x.mutations == y.mutations


@pytest.mark.parametrize('name', G.JOINT)
def test_mirror_joint_matches_reference_golden(name):
    """N2 host logic (N_diff bookkeeping incl. tips, attributes) against the reference's golden output."""
    zj = G.load(name)
    z = G.load(str(zj['source']))
    tips = bool(zj['reconstruct_tips'])
    tt = mirror_from_golden(z)
    assert tt.infer_ancestral_sequences(marginal=False, reconstruct_tip_states=tips, debug=True) == int(zj['N_diff_first'])
    assert tt.tree.sequence_joint_LH == float(zj['sequence_joint_LH']) and np.array_equal(tt.tree.root.joint_Lx, zj['root_joint_Lx'])
    assert np.array_equal(tt.tree.sequence_LH, zj['sequence_LH'])
    for n, s in zip(tt.tree.find_clades(), zj['cseq']):
        if tips or not n.is_terminal():
            assert ''.join(n.cseq) == str(s)
    assert tt.infer_ancestral_sequences(marginal=False, reconstruct_tip_states=tips) == int(zj['N_diff_second'])
    assert tt.infer_ancestral_sequences(marginal=True, reconstruct_tip_states=tips) == int(zj['N_diff_marginal_after'])


def test_mirror_joint_branch_lengths():
    """N2: optimize_tree(branch_length_mode='joint') -- pair counts from the engine, all branches in one
    lock-step Brent -- against the oracle's per-branch restatement of GTR.optimal_t_compressed."""
    import flat_numpy as O
    from treetime_b200.flatten import flatten_treeanc
    z = G.load('nuc40')
    tt = mirror_from_golden(z)
    tt.infer_ancestral_sequences(marginal=False)
    topo, flat, g = flatten_treeanc(tt)
    gt = O.make_gtr(g)
    res = O.joint(flat, g)
    ab = [str(c) for c in tt.gtr.alphabet]
    chars = sorted(tt.gtr.profile_map.keys())
    want = []
    for n in range(1, flat['parent'].shape[0]):
        sp = np.array(ab)[res.seq_idx[flat['parent'][n]]]
        sc = np.array(chars + ['?'])[flat['tip_codes'][flat['tip_row'][n]]] if flat['tip_row'][n] >= 0 else np.array(ab)[res.seq_idx[n]]
        pairs, mult = O.state_pair(ab, tt.gtr.gap_index, sp, sc, flat['multiplicity'], ignore_gaps=tt.ignore_gaps)
        bs = tt._branch_state(topo.nodes[n])
        assert np.array_equal(bs['pair'], pairs) and np.array_equal(bs['multiplicity'], mult)
        want.append(O.optimal_t_compressed(gt, pairs, mult))
    one = tt.optimal_branch_length(topo.nodes[5])
    assert abs(one - want[4]) <= 5e-6 * max(want[4], 1e-6)       # minima from function values: ~sqrt(eps) relative
    tt.optimize_branch_lengths_joint()
    got = np.array([n.branch_length for n in topo.nodes[1:]])
    assert np.allclose(got, np.maximum(0, want), rtol=5e-6, atol=1e-10)
    # the full loop terminates and leaves a consistent tree
    tt2 = mirror_from_golden(z)
    tt2.optimize_tree(branch_length_mode='joint', max_iter=3, prune_short=True)
    assert np.isfinite(tt2.tree.unconstrained_sequence_LH) and tt2.sequence_reconstruction == 'joint'
    # tip without sequence: the reference's error
    z2 = G.load('nuc40')
    aln = G.alignment(z2)
    del aln[sorted(aln)[0]]
    tt3 = TreeAnc(tree=str(z2['newick']), aln=aln, gtr=G.model(z2), engine_factory=oracle_engine.factory)
    tt3.infer_ancestral_sequences(marginal=False)
    from treetime_b200.treeanc import MissingDataError
    with pytest.raises(MissingDataError):
        tt3.optimize_branch_lengths_joint()
    tt3.infer_ancestral_sequences(marginal=False, reconstruct_tip_states=True)
    tt3.optimize_branch_lengths_joint()


@pytest.mark.parametrize('name', ['joint_nuc40', 'joint_poly70'])
def test_mirror_joint_branch_length_optimisation_golden(name):
    zj = G.load(name)
    z = G.load(str(zj['source']))
    tt = mirror_from_golden(z)
    tt.optimize_tree(branch_length_mode='joint', max_iter=2, prune_short=False)
    got = np.array([n.branch_length for n in tt.tree.find_clades()])
    assert np.allclose(got[1:], zj['opt_joint_branch_length'][1:], rtol=5e-6, atol=1e-10)
    tot = float(zj['opt_joint_sequence_LH'])
    assert abs(tt.tree.unconstrained_sequence_LH - tot) <= 1e-7 * abs(tot)


def test_mirror_joint_reconstruction():
    """N2 through the mirror API: infer_ancestral_sequences(marginal=False)."""
    z = G.load('nuc40')
    tt = mirror_from_golden(z)
    n1 = tt.infer_ancestral_sequences(marginal=False)
    assert n1 == (len(z['node_names']) - 40 - 1) * tt.data.compressed_length
    assert tt.sequence_reconstruction == 'joint' and np.isfinite(tt.tree.sequence_joint_LH)
    assert tt.infer_ancestral_sequences(marginal=False) == 0
    with pytest.raises(AttributeError):
        tt.tree.root.marginal_profile          # marginal state does not exist after a joint pass
    assert tt.infer_ancestral_sequences(marginal=True) >= 0 and tt.tree.root.marginal_profile.shape[1] == 5


@pytest.mark.parametrize('name', G.EXTRAS)
def test_mirror_sampling_masks_site_specific_inference_vs_reference_golden(name):
    """Golden vectors of the unmodified reference for sample_from_profile=True, per-branch masks and
    infer_gtr(site_specific=True) (oracle/make_golden.py --only-extras) through the mirror on the CPU oracle engine."""
    import oracle_engine
    zx = G.load(name)
    z = G.load(str(zx['source']))
    G.check_extras(lambda **kw: TreeAnc(tree=str(z['newick']), aln=G.alignment(z), gtr=G.model(z),
                                        engine_factory=oracle_engine.factory, **kw), zx, exact=True)


def _small_problem(newick, L=120, seed=5):
    from treetime_b200 import synth
    from treetime_b200.gtr import GTR
    from treetime_b200.tree import read_newick
    g = GTR.custom(pi=np.array([.3, .2, .2, .29, .01]), W=np.ones((5, 5)), alphabet='nuc')
    T = read_newick(newick)
    idx = synth.evolve_alignment(T, L, g.Pi, g.W, seed=seed, mu=30.0)
    aln = {k: g.alphabet[v] for k, v in idx.items()}
    return g, aln


def test_same_shape_tip_permutation_reuploads_the_tip_codes():
    """Advisor finding (round 1): swapping two leaves keeps parent / child_idx identical; the device must still
    get the new row -> tip assignment instead of returning the LH of the old one."""
    nwk = '(A:0.02,(B:0.03,(C:0.01,D:0.04):0.02):0.01);'
    g, aln = _small_problem(nwk)
    tt = TreeAnc(tree=nwk, aln=aln, gtr=g, rng_seed=1, engine_factory=oracle_engine.factory)
    tt.infer_ancestral_sequences(marginal=True)
    lh0 = tt.sequence_LH()
    tips = {n.name: n for n in tt.tree.get_terminals()}
    a, d = tips['A'], tips['D']
    pa, pd = a.up, d.up
    ia, id_ = pa.clades.index(a), pd.clades.index(d)
    pa.clades[ia], pd.clades[id_] = d, a
    a.branch_length, d.branch_length = d.branch_length, a.branch_length
    a.mutation_length, d.mutation_length = a.branch_length, d.branch_length
    tt._prepare_nodes()
    tt.infer_ancestral_sequences(marginal=True)
    swapped = '(D:0.02,(B:0.03,(C:0.01,A:0.04):0.02):0.01);'
    fresh = TreeAnc(tree=swapped, aln=aln, gtr=g, rng_seed=1, engine_factory=oracle_engine.factory)
    fresh.infer_ancestral_sequences(marginal=True)
    assert abs(lh0 - fresh.sequence_LH()) > 1e-6          # the permutation matters for this alignment
    assert tt.sequence_LH() == fresh.sequence_LH()


def test_n_diff_after_a_topology_change_compares_surviving_nodes():
    """Advisor finding (round 1): after collapsing an internal node the next pass must compare every surviving node
    with its own previous sequence (treeanc.py:925-926), not count every position as changed."""
    nwk = '((A:0.02,B:0.03):0.01,((C:0.01,D:0.04):0.00001,E:0.02):0.02,F:0.03);'
    g, aln = _small_problem(nwk, L=150, seed=9)
    tt = TreeAnc(tree=nwk, aln=aln, gtr=g, rng_seed=1, engine_factory=oracle_engine.factory)
    tt.infer_ancestral_sequences(marginal=True)
    before = {id(n): n.cseq.copy() for n in tt.tree.get_nonterminals()}
    cd = [n for n in tt.tree.get_nonterminals() if sorted(c.name or '' for c in n.clades) == ['C', 'D']][0]
    up = cd.up
    k = up.clades.index(cd)
    for c in cd.clades:
        c.branch_length += cd.branch_length
        c.mutation_length = c.branch_length
    up.clades[k:k + 1] = cd.clades
    tt._prepare_nodes()
    n_diff = tt.infer_ancestral_sequences(marginal=True)
    expect = sum(int((n.cseq != before[id(n)]).sum()) for n in tt.tree.get_nonterminals() if n.up is not None)
    assert n_diff == expect
    assert n_diff < tt.data.compressed_length            # far from "every position of every node"
    assert tt.infer_ancestral_sequences(marginal=True) == 0


def test_sparse_host_interface_gives_the_same_reconstruction():
    """TreeAnc(sparse_io=True): tip codes uploaded as reference row + differences, sequences read back as root row +
    differences from the parent (sequence_differences); same numbers as the dense interface."""
    from treetime_b200.sparse import expand_mutations
    z = G.load('nuc40')
    dense = mirror_from_golden(z)
    sparse = mirror_from_golden(z, sparse_io=True)
    assert dense.infer_ancestral_sequences(marginal=True) == sparse.infer_ancestral_sequences(marginal=True)
    assert dense.sequence_LH() == sparse.sequence_LH()
    root, node, pos, state = sparse.sequence_differences()
    topo = sparse._flat()
    idx = expand_mutations(topo.parent, topo.tip_row, root, node, pos, state)
    for k, n in enumerate(topo.internal_nodes):
        assert (sparse.gtr.alphabet[idx[k]] == topo.nodes[n].cseq).all()
        assert (dense._flat().nodes[n].cseq == topo.nodes[n].cseq).all()
