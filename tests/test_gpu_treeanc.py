"""GPU: the TreeAnc mirror on the real CUDA engine against golden vectors of the unmodified
reference (tests/golden) -- the tests read like the reference's own (test/test_treetime.py)."""
import numpy as np
import pytest

import golden_util as G
from treetime_b200.treeanc import TreeAnc

pytestmark = pytest.mark.gpu

LH_RTOL = 1e-9       # BASELINE.json: total log-LH within 1e-9 relative (fp64)
PROF_ATOL = 1e-6     # BASELINE.json: marginal profiles within 1e-6


def gpu_from_golden(z, **kw):
    return TreeAnc(tree=str(z['newick']), aln=G.alignment(z), gtr=G.model(z), rng_seed=1, **kw)


def check(tt, z):
    import util
    tips = bool(z['reconstruct_tips'])
    n1 = tt.infer_ancestral_sequences(marginal=True, reconstruct_tip_states=tips)
    assert n1 == int(z['N_diff_first'])
    tot = float(z['total_LH'])
    assert abs(tt.sequence_LH() - tot) <= LH_RTOL * abs(tot)
    assert np.allclose(tt.tree.sequence_LH, z['sequence_LH'], rtol=1e-11, atol=1e-11)
    nodes = list(tt.tree.find_clades())
    worst = 0.0
    for i in z['stored_nodes']:
        n = nodes[i]
        worst = max(worst, np.abs(n.marginal_subtree_LH - z['subtree_%d' % i]).max())
        if i > 0:
            worst = max(worst, np.abs(n.marginal_outgroup_LH - z['outgroup_%d' % i]).max())
        if 'profile_%d' % i in z.files:
            worst = max(worst, np.abs(n.marginal_profile - z['profile_%d' % i]).max())
    assert worst < PROF_ATOL
    for n, s in zip(nodes, z['cseq']):
        c = n.cseq
        if c is None:
            assert str(s) == ''
            continue
        bad = np.array(list(str(s))) != c
        if bad.any():       # identical except at exact ties
            assert util.tie_mask(n.marginal_profile)[bad].all(), n.name
    assert tt.infer_ancestral_sequences(marginal=True, reconstruct_tip_states=tips) == int(z['N_diff_second'])
    return nodes, worst


@pytest.mark.parametrize('name', G.SMALL)
def test_gpu_reconstruction_matches_reference_golden(name):
    z = G.load(name)
    tt = gpu_from_golden(z)
    nodes, worst = check(tt, z)
    for k, i in enumerate(z['bl_nodes']):
        bl = tt.optimal_marginal_branch_length(nodes[i])
        assert abs(bl - z['bl_opt'][k]) <= 1e-6 * z['bl_opt'][k] + 1e-12
    print('%s: rel dLH=%.1e  max|dprofile|=%.1e' % (name, abs(tt.sequence_LH() - float(z['total_LH'])) / abs(float(z['total_LH'])), worst))


@pytest.mark.parametrize('name', G.JOINT)
def test_gpu_joint_reconstruction_matches_reference_golden(name):
    """N2: infer_ancestral_sequences(marginal=False) on the device vs the reference's _ml_anc_joint."""
    import util
    zj = G.load(name)
    z = G.load(str(zj['source']))
    tips = bool(zj['reconstruct_tips'])
    tt = gpu_from_golden(z)
    assert tt.infer_ancestral_sequences(marginal=False, reconstruct_tip_states=tips, debug=True) == int(zj['N_diff_first'])
    tot = float(zj['sequence_joint_LH'])
    assert abs(tt.tree.sequence_joint_LH - tot) <= LH_RTOL * abs(tot)
    assert np.allclose(tt.tree.sequence_LH, zj['sequence_LH'], rtol=1e-11, atol=1e-9)
    assert np.allclose(tt.tree.root.joint_Lx, zj['root_joint_Lx'], rtol=1e-11, atol=1e-9)
    nodes = list(tt.tree.find_clades())
    flat, g = G.flat_and_gtr(z)
    seqs, n_bad = [None] * len(nodes), 0
    lut = {c: i for i, c in enumerate(tt.gtr.alphabet)}
    for i, (n, s) in enumerate(zip(nodes, zj['cseq'])):
        if n.is_terminal() and not tips:
            continue
        c = n.cseq
        n_bad += int((np.array(list(str(s))) != c).sum())
        seqs[i] = np.array([lut[x] for x in c])
    if n_bad:        # states may differ only where two assignments are equally likely to rounding
        assert n_bad < 2e-3 * len(nodes) * tt.data.compressed_length
        assert np.allclose(util.joint_assignment_lh(flat, g, seqs), zj['sequence_LH'], rtol=1e-10, atol=1e-8)
    assert tt.infer_ancestral_sequences(marginal=False, reconstruct_tip_states=tips) == int(zj['N_diff_second'])
    assert tt.infer_ancestral_sequences(marginal=True, reconstruct_tip_states=tips) == int(zj['N_diff_marginal_after']) or n_bad


@pytest.mark.parametrize('name', ['joint_nuc40', 'joint_poly70'])
def test_gpu_joint_branch_length_optimisation_golden(name):
    """N2: optimize_tree(branch_length_mode='joint') -- device pair counts + lock-step Brent -- vs the reference."""
    zj = G.load(name)
    z = G.load(str(zj['source']))
    tt = gpu_from_golden(z)
    tt.optimize_tree(branch_length_mode='joint', max_iter=2, prune_short=False)
    got = np.array([n.branch_length for n in tt.tree.find_clades()])
    want = zj['opt_joint_branch_length']
    # minima located from function values agree to ~sqrt(eps); zero-length branches exactly
    assert np.allclose(got[1:], want[1:], rtol=5e-6, atol=1e-10)
    tot = float(zj['opt_joint_sequence_LH'])
    assert abs(tt.tree.unconstrained_sequence_LH - tot) <= 1e-7 * abs(tot)


def test_gpu_site_specific_golden():
    """Site-specific model (reference default: interpolated expQt) against the reference's output."""
    from treetime_b200.gtr import GTRSiteSpecific
    z = G.load('sitespec20')
    ab = [str(c) for c in z['gtr_alphabet']]
    gtr = GTRSiteSpecific(seq_len=int(z['gtr_mu'].shape[0]), approximate=bool(z['gtr_approximate']), alphabet='nuc')
    assert list(gtr.alphabet) == ab
    gtr._W, gtr._Pi, gtr._mu = z['gtr_W'].copy(), z['gtr_Pi'].copy(), z['gtr_mu'].copy()
    gtr.eigenvals, gtr.v, gtr.v_inv = z['gtr_eigenvals'].copy(), z['gtr_v'].copy(), z['gtr_v_inv'].copy()
    gtr.rate_scale = float(z['gtr_rate_scale'])
    tt = TreeAnc(tree=str(z['newick']), aln=G.alignment(z), gtr=gtr, rng_seed=1, compress=False)
    nodes, worst = check(tt, z)
    for k, i in enumerate(z['bl_nodes']):
        bl = tt.optimal_marginal_branch_length(nodes[i])
        assert abs(bl - z['bl_opt'][k]) <= 1e-6 * z['bl_opt'][k] + 1e-12
    with pytest.raises(TypeError):
        TreeAnc(tree=str(z['newick']), aln=G.alignment(z), gtr=gtr, compress=True)     # treeanc.py:186-187


def test_gpu_known_answer_lh_normalisation():
    """The reference's own KAT (test/test_treetime.py:137-155): sum over all 4^3 patterns of exp(LH) = 1."""
    z = G.load('kat3')
    tt = gpu_from_golden(z)
    tt.reconstruct_anc('ml', marginal=True, debug=True)
    assert abs(np.exp(tt.tree.sequence_LH).sum() - 1.0) < 1e-6
    assert abs(tt.tree.total_sequence_LH - (-495.7525153086474)) < 1e-9 * 495.75
    tt.optimize_branch_len()


@pytest.mark.parametrize('name', ['nuc40', 'poly70'])
def test_gpu_optimize_tree_marginal(name):
    z = G.load(name)
    tt = gpu_from_golden(z)
    tt.optimize_tree(branch_length_mode='marginal', max_iter=2, infer_gtr=False, prune_short=False)
    bl = np.array([n.branch_length for n in tt.tree.find_clades()])
    ref = z['opt_branch_length']
    # Brent tolerance of the sweeps is 1e-2 / 1e-4 relative in s = sqrt(t): the iterates agree far tighter
    assert np.allclose(bl[1:], ref[1:], rtol=1e-6, atol=1e-10), np.abs(bl[1:] - ref[1:]).max()
    assert abs(tt.sequence_LH() - float(z['opt_total_LH'])) <= 1e-9 * abs(float(z['opt_total_LH']))


def test_gpu_infer_gtr_and_rate():
    z = G.load('nuc40')
    tt = gpu_from_golden(z)
    tt.optimize_tree(branch_length_mode='marginal', max_iter=2, infer_gtr=False, prune_short=False)
    tt.infer_gtr(marginal=True)
    assert np.allclose(tt.gtr.W, z['inferred_W'], rtol=1e-6) and np.allclose(tt.gtr.Pi, z['inferred_Pi'], rtol=1e-6)
    tt.optimize_gtr_rate()
    assert tt.gtr.mu > 0


@pytest.mark.parametrize('name', G.BIG)
def test_gpu_full_size_configs(name):
    """BASELINE.json configs[0] and configs[1] at full size against the reference's output."""
    z = G.load(name)
    tree, aln, g, sha = G.regenerate_big(name)
    assert sha == str(z['input_sha'])
    tt = TreeAnc(tree=tree, aln=aln, gtr=g, rng_seed=1)
    assert tt.infer_ancestral_sequences(marginal=True) == int(z['N_diff_first'])
    tot = float(z['total_LH'])
    assert abs(tt.sequence_LH() - tot) <= LH_RTOL * abs(tot)
    assert np.allclose(tt.tree.sequence_LH, z['sequence_LH'], rtol=1e-11, atol=1e-10)
    nodes = list(tt.tree.find_clades())
    if 'cseq_idx' in z.files:
        internal = [i for i, n in enumerate(nodes) if not n.is_terminal()]
        idx = tt._engine.all_seq_idx()
        ref = z['cseq_idx'][internal]
        mism = int((idx != ref).sum())
        assert mism <= 5, mism                 # only exact ties may differ
        if mism:
            import util
            for r, c in zip(*np.nonzero(idx != ref)):
                assert util.tie_mask(nodes[internal[r]].marginal_profile)[c]
    else:
        for n, s in zip(nodes, z['cseq']):
            if not n.is_terminal():
                assert ''.join(n.cseq) == str(s)
    print('%s: rel dLH = %.2e' % (name, abs(tt.sequence_LH() - tot) / abs(tot)))


def test_gpu_one_vs_two_engine_shards_agree():
    """Pattern sharding on one GPU: two engines, half the patterns each, partial sums add up."""
    import util
    from treetime_b200 import synth
    tree = synth.random_tree(80, seed=31, mean_bl=0.01)
    topo, flat, g = util.make_flat(tree, util.nuc_gtr(), 700, 31, amb_frac=0.01)
    full = util.engine_for(flat, g)
    full.marginal()
    tot, nd = full.results()
    L = flat['multiplicity'].shape[0]
    parts = []
    for lo, hi in ((0, L // 2), (L // 2, L)):
        f = dict(flat); f['tip_codes'] = np.ascontiguousarray(flat['tip_codes'][:, lo:hi]); f['multiplicity'] = flat['multiplicity'][lo:hi]
        e = util.engine_for(f, g)
        e.marginal()
        parts.append((e, e.results()))
    assert abs(sum(p[1][0] for p in parts) - tot) <= 1e-12 * abs(tot)
    assert sum(p[1][1] for p in parts) == nd
    n = int(flat['child_idx'][0])
    whole = full.node_array(n, 1)
    halves = np.vstack([p[0].node_array(n, 1) for p in parts])
    assert np.array_equal(whole, halves)


def test_gpu_device_pattern_compression():
    """N3: column statistics + pattern gather on the device == host SequenceData compression."""
    from treetime_b200.sequence_data import SequenceData
    z = G.load('nuc40')
    a = gpu_from_golden(z)
    b = gpu_from_golden(z, device_compress=True)
    assert b.data.device_resident
    assert np.array_equal(a.data.multiplicity(), b.data.multiplicity())
    assert np.array_equal(a.data.full_to_compressed_sequence_map, b.data.full_to_compressed_sequence_map)
    assert np.array_equal(a.data.compressed_matrix, b.data.compressed_matrix)
    check(b, z)
    assert a.infer_ancestral_sequences(marginal=True) == int(z['N_diff_first'])
    assert a.sequence_LH() == b.sequence_LH()
    # larger, regenerated input incl. overhang gaps and an unknown character
    tree, aln, g, sha = G.regenerate_big('cfg1')
    names = sorted(aln)
    aln[names[0]][:9] = '-'; aln[names[1]][-5:] = '-'; aln[names[2]][100:130] = 'N'
    h = TreeAnc(tree=tree.to_newick(), aln=aln, gtr=g)
    d = TreeAnc(tree=tree.to_newick(), aln=aln, gtr=g, device_compress=True)
    assert np.array_equal(h.data.multiplicity(), d.data.multiplicity())
    assert h.infer_ancestral_sequences(marginal=True) == d.infer_ancestral_sequences(marginal=True)
    assert h.sequence_LH() == d.sequence_LH()
    assert np.array_equal(h._engine.all_seq_idx(), d._engine.all_seq_idx())


def test_gpu_sample_from_profile_all_nodes():
    """infer_ancestral_sequences(marginal=True, sample_from_profile=True) on the device: same sampled sequences and
    N_diff as the same TreeAnc on the CPU oracle engine (which test_reference_live pins to the reference), the
    host generator consumed identically."""
    import oracle_engine
    z = G.load('nuc40')
    a = gpu_from_golden(z)
    b = gpu_from_golden(z, engine_factory=oracle_engine.factory)
    for kw in (dict(), dict(reconstruct_tip_states=True), dict(reconstruct_tip_states=True), dict()):
        na = a.infer_ancestral_sequences(marginal=True, sample_from_profile=True, **kw)
        nb = b.infer_ancestral_sequences(marginal=True, sample_from_profile=True, **kw)
        diff = 0
        for x, y in zip(a.tree.find_clades(), b.tree.find_clades()):
            if kw or not x.is_terminal():
                diff += int((x.cseq != y.cseq).sum())
        assert diff <= 1            # a draw within rounding of a cumulative sum may fall either way
        assert abs(na - nb) <= 2 * diff and na > 0
    assert a.rng.random() == b.rng.random()
    assert a._engine.launch_count() > 0


def test_gpu_infer_site_specific_gtr():
    """infer_gtr(marginal=True, site_specific=True): the device's per-pattern statistics give the same
    GTRSiteSpecific as the CPU oracle engine (pinned to the reference in test_reference_live), and the following
    reconstruction under that model agrees."""
    import oracle_engine
    z = G.load('nuc40')
    a = gpu_from_golden(z, compress=False)
    b = gpu_from_golden(z, compress=False, engine_factory=oracle_engine.factory)
    for tt in (a, b):
        tt.infer_ancestral_sequences(marginal=True)
    for it in range(2):
        ga = a.infer_gtr(marginal=True, site_specific=True, pc=1.0)
        gb = b.infer_gtr(marginal=True, site_specific=True, pc=1.0)
        assert ga.is_site_specific
        assert np.allclose(ga.Pi, gb.Pi, rtol=1e-8, atol=1e-12) and np.allclose(ga.mu, gb.mu, rtol=1e-8) and np.allclose(ga.W, gb.W, rtol=1e-8)
        a.infer_ancestral_sequences(marginal=True); b.infer_ancestral_sequences(marginal=True)
        assert abs(a.sequence_LH() - b.sequence_LH()) <= LH_RTOL * abs(b.sequence_LH())


@pytest.mark.parametrize('name', G.EXTRAS)
def test_gpu_sampling_masks_site_specific_inference_vs_reference_golden(name):
    """The same golden vectors of the unmodified reference (sampled sequences, per-branch masks, site-specific GTR
    inference) against the CUDA engine."""
    zx = G.load(name)
    z = G.load(str(zx['source']))
    G.check_extras(lambda **kw: TreeAnc(tree=str(z['newick']), aln=G.alignment(z), gtr=G.model(z), **kw), zx, exact=False)


def test_gpu_infer_gtr_from_reconstructed_sequences():
    """infer_gtr(marginal=False): counts from the device's pair tables equal those of the CPU oracle engine (pinned to
    the reference in test_reference_live), after joint and marginal reconstructions, with and without tips."""
    import oracle_engine
    z = G.load('nuc40')
    for recon in (dict(marginal=False), dict(marginal=True), dict(marginal=False, reconstruct_tip_states=True)):
        a = gpu_from_golden(z)
        b = gpu_from_golden(z, engine_factory=oracle_engine.factory)
        assert a.infer_ancestral_sequences(**recon) == b.infer_ancestral_sequences(**recon)
        ga, gb = a.infer_gtr(marginal=False, pc=2.0), b.infer_gtr(marginal=False, pc=2.0)
        assert np.allclose(ga.W, gb.W, rtol=1e-10) and np.allclose(ga.Pi, gb.Pi, rtol=1e-10)
