"""Helpers to load tests/golden/*.npz (made by oracle/make_golden.py from the unmodified reference)."""
import os
import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
SMALL = ['kat3', 'nuc40', 'nuc40_tips', 'poly70', 'aa16_jtt92', 'aa16_q22']
SITE_SPECIFIC = ['sitespec20']
BIG = ['cfg1_200x1400', 'cfg2_2000x10000']
JOINT = ['joint_nuc40', 'joint_nuc40_tips', 'joint_poly70', 'joint_aa16_jtt92_tips']      # N2; inputs live in z['source']


def load(name):
    return np.load(os.path.join(GOLDEN, name + '.npz'), allow_pickle=False)


def flat_and_gtr(z):
    """Flat problem + model dict exactly as flattened from the reference's own objects."""
    flat = {k[5:]: z[k] for k in z.files if k.startswith('flat_')}
    g = dict(eigenvals=z['gtr_eigenvals'], v=z['gtr_v'], v_inv=z['gtr_v_inv'], Pi=z['gtr_Pi'], gap_index=None)
    ab = [str(c) for c in z['gtr_alphabet']]
    if '-' in ab:
        g['gap_index'] = ab.index('-')
    if z['gtr_Pi'].ndim == 2:
        g.update(site_specific=True, mu=z['gtr_mu'], rate_scale=float(z['gtr_rate_scale']), approximate=bool(z['gtr_approximate']))
    else:
        g.update(site_specific=False, mu=float(z['gtr_mu']))
    return flat, g


def model(z):
    """treetime_b200 GTR carrying the reference model's numbers verbatim: alphabet, profile map and
    ambiguous character as the user built it, (W, Pi, mu) and the reference's eigen-system."""
    from treetime_b200.gtr import GTR
    ab = np.array([str(c) for c in z['gtr_alphabet']])
    prof_map = {str(c): row.copy() for c, row in zip(z['gtr_prof_chars'], z['gtr_prof_table'])}
    g = GTR(alphabet=ab, prof_map=prof_map)
    g._W, g._Pi, g._mu = z['gtr_W'].copy(), z['gtr_Pi'].copy(), float(z['gtr_mu'])
    g.eigenvals, g.v, g.v_inv = z['gtr_eigenvals'].copy(), z['gtr_v'].copy(), z['gtr_v_inv'].copy()
    amb = str(z['gtr_ambiguous'])
    g.ambiguous = amb if amb else None
    return g


def alignment(z):
    return {str(k): np.array(list(str(s))) for k, s in zip(z['aln_names'], z['aln_seqs'])}


def regenerate_big(name):
    """Inputs of the big cases are regenerated from seeds; a checksum pins the generator."""
    import hashlib
    from treetime_b200 import synth
    from treetime_b200.gtr import GTR
    g = GTR.custom(pi=np.array([.3, .2, .2, .29, .01]), W=np.ones((5, 5)), alphabet='nuc')
    if name.startswith('cfg1'):
        tree = synth.random_tree(200, seed=1, mean_bl=2e-3)
        idx = synth.evolve_alignment(tree, 1400, g.Pi, g.W, seed=1)
    else:
        tree = synth.random_tree(2000, seed=1, mean_bl=5e-4)
        idx = synth.evolve_alignment(tree, 10000, g.Pi, g.W, seed=1)
    sha = hashlib.sha256(np.ascontiguousarray(np.vstack([idx[k] for k in sorted(idx)])).tobytes()).hexdigest()
    aln = {k: g.alphabet[v] for k, v in idx.items()}
    return tree, aln, g, sha


EXTRAS = ['extras_nuc40']       # sampling / masks / site-specific inference on the inputs of z['source']


def check_extras(make_tt, zx, exact):
    """The TreeAnc built by make_tt(**kw) against the reference's outputs in an extras_* fixture.
    exact: bit-level agreement expected (CPU oracle engine) or only to the stated tolerances (CUDA engine)."""
    rt = 0.0 if exact else 1e-9
    # 1. sampled sequences: same generator, same draws
    tt = make_tt(rng_seed=7)
    nodes = list(tt.tree.find_clades())
    for k, tips in enumerate((False, True)):
        nd = tt.infer_ancestral_sequences(marginal=True, sample_from_profile=True, reconstruct_tip_states=tips)
        got = [''.join(n.cseq) if (tips or not n.is_terminal()) else None for n in nodes]
        diff = sum(sum(a != b for a, b in zip(g, str(r))) for g, r in zip(got, zx['sample_cseq_%d' % k]) if g is not None)
        assert diff <= (0 if exact else 1), diff        # a uniform within rounding of a cumulative sum may fall either way
        assert abs(nd - int(zx['sample_N_diff_%d' % k])) <= 2 * diff
    assert tt.rng.random() == float(zx['sample_next_uniform'])
    # 2. per-branch masks
    tt = make_tt(rng_seed=1)
    nodes = list(tt.tree.find_clades())
    seg, L = zx['mask_segment'], tt.data.compressed_length
    for k, n in enumerate(nodes):
        n.mask = seg if k % 3 == 0 else np.ones(L)
    assert tt.infer_ancestral_sequences(marginal=True) == int(zx['mask_N_diff'])
    tot = float(zx['mask_total_LH'])
    assert abs(tt.sequence_LH() - tot) <= max(rt, 1e-13) * abs(tot)
    assert np.allclose(tt.tree.sequence_LH, zx['mask_sequence_LH'], rtol=1e-11, atol=1e-11)
    for n, s in zip(nodes, zx['mask_cseq']):
        if not n.is_terminal():
            bad = n.cseq != np.array(list(str(s)))
            if bad.any():       # identical except at exact ties (e.g. a pattern masked on every branch below a node)
                import util
                assert not exact and util.tie_mask(n.marginal_profile)[bad].all(), n.name
    for i in zx['mask_profile_nodes']:
        assert np.allclose(nodes[i].marginal_profile, zx['mask_profile_%d' % i], rtol=0, atol=1e-12)
        assert np.allclose(nodes[i].marginal_outgroup_LH, zx['mask_outgroup_%d' % i], rtol=0, atol=1e-12)
    for i, bl in zip(zx['mask_bl_nodes'], zx['mask_bl_opt']):
        got = tt.optimal_marginal_branch_length(nodes[i])
        assert abs(got - bl) <= 1e-6 * bl + 1e-12, (i, got, bl)
    g = tt.infer_gtr(marginal=True, pc=1.0)
    assert np.allclose(g.W, zx['mask_inferred_W'], rtol=1e-8) and np.allclose(g.Pi, zx['mask_inferred_Pi'], rtol=1e-8)
    # 3. site-specific GTR inference
    tt = make_tt(rng_seed=1, compress=False)
    tt.infer_ancestral_sequences(marginal=True)
    g = tt.infer_gtr(marginal=True, site_specific=True, pc=1.0)
    assert np.allclose(g.Pi, zx['ss_Pi'], rtol=1e-8, atol=1e-12) and np.allclose(g.mu, zx['ss_mu'], rtol=1e-8)
    assert np.allclose(g.W, zx['ss_W'], rtol=1e-8)
    assert tt.infer_ancestral_sequences(marginal=True) == int(zx['ss_N_diff'])
    tot = float(zx['ss_total_LH'])
    assert abs(tt.sequence_LH() - tot) <= 1e-9 * abs(tot)
