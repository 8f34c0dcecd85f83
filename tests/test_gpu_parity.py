"""GPU parity: CUDA engine (through the C-ABI) vs the CPU oracle on seeded inputs."""
import numpy as np
import pytest

import flat_numpy as O
import util
from treetime_b200 import synth

pytestmark = pytest.mark.gpu

LH_RTOL = 1e-9      # BASELINE.json: total log-LH within 1e-9 relative (fp64)
PROF_ATOL = 1e-6    # BASELINE.json: marginal profiles within 1e-6


def compare_all(flat, g, eng, res, reconstruct_tips=False, every=1):
    n_nodes = flat['parent'].shape[0]
    site = eng.site_lh()
    assert np.allclose(site, res.sequence_LH, rtol=1e-11, atol=1e-11)
    worst = 0.0
    for n in range(0, n_nodes, every):
        tip = flat['tip_row'][n] >= 0
        s = eng.node_array(n, 0)
        worst = max(worst, np.abs(s - res.subtree_LH[n]).max())
        o = eng.node_array(n, 1)
        worst = max(worst, np.abs(o - res.outgroup_LH[n]).max())
        if not tip or reconstruct_tips:
            p = eng.node_array(n, 2)
            worst = max(worst, np.abs(p - res.profile[n]).max())
            idx = eng.seq_idx([n])[0]
            bad = idx != res.seq_idx[n]
            if bad.any():
                assert util.tie_mask(res.profile[n])[bad].all(), 'argmax differs off ties at node %d' % n
    assert worst < PROF_ATOL, worst
    return worst


@pytest.mark.parametrize('n_tips,L,seed', [(3, 64, 1), (50, 300, 2), (200, 1400, 3)])
def test_nuc_binary(n_tips, L, seed):
    tree = synth.random_tree(n_tips, seed=seed, mean_bl=0.01)
    topo, flat, g = util.make_flat(tree, util.nuc_gtr(), L, seed, amb_frac=0.02)
    eng = util.engine_for(flat, g)
    eng.marginal()
    tot, nd = eng.results()
    res = O.marginal(flat, g)
    assert abs(tot - res.total_LH) <= LH_RTOL * abs(res.total_LH)
    assert nd == res.N_diff
    w = compare_all(flat, g, eng, res, every=1 if n_tips <= 50 else 7)
    # second pass: nothing changes
    eng.marginal()
    tot2, nd2 = eng.results()
    assert tot2 == tot and nd2 == 0
    print('nuc n=%d L\'=%d  relLH=%.2e  max|dprof|=%.2e' % (n_tips, flat['multiplicity'].shape[0],
                                                         abs(tot - res.total_LH) / abs(res.total_LH), w))


def test_polytomies_and_zero_branches():
    tree = synth.random_tree(300, seed=4, mean_bl=0.005, polytomy_frac=0.5, zero_frac=0.3)
    topo, flat, g = util.make_flat(tree, util.nuc_gtr(), 600, 4, amb_frac=0.01)
    assert np.diff(flat['child_ptr']).max() > 4
    eng = util.engine_for(flat, g)
    eng.marginal()
    tot, nd = eng.results()
    res = O.marginal(flat, g)
    assert abs(tot - res.total_LH) <= LH_RTOL * abs(res.total_LH)
    compare_all(flat, g, eng, res, every=5)


def test_star_tree_huge_polytomy():
    """One node with 3000 children: exercises the power-of-two rescaling."""
    from treetime_b200.tree import Node, Tree
    rng = np.random.default_rng(7)
    root = Node(clades=[Node(name='t%06d' % i, branch_length=float(b)) for i, b in enumerate(rng.exponential(0.3, 3000))])
    tree = Tree(root)
    topo, flat, g = util.make_flat(tree, util.nuc_gtr(), 200, 7)
    eng = util.engine_for(flat, g)
    eng.marginal()
    tot, _ = eng.results()
    res = O.marginal(flat, g)
    assert np.isfinite(tot)
    assert abs(tot - res.total_LH) <= LH_RTOL * abs(res.total_LH)
    assert np.allclose(eng.site_lh(), res.sequence_LH, rtol=1e-11)


def test_reconstruct_tip_states():
    tree = synth.random_tree(40, seed=5, mean_bl=0.02)
    topo, flat, g = util.make_flat(tree, util.nuc_gtr(), 250, 5, amb_frac=0.05)
    eng = util.engine_for(flat, g)
    eng.marginal(reconstruct_tips=True)
    tot, nd = eng.results()
    res = O.marginal(flat, g, reconstruct_tip_states=True)
    assert abs(tot - res.total_LH) <= LH_RTOL * abs(res.total_LH)
    assert nd == res.N_diff
    compare_all(flat, g, eng, res, reconstruct_tips=True)


@pytest.mark.parametrize('alphabet,q', [('nuc_nogap', 4), ('aa_nogap', 20), ('aa', 22)])
def test_other_alphabets(alphabet, q):
    gtr = util.random_gtr(alphabet, 11)
    assert gtr.n_states == q
    tree = synth.random_tree(60, seed=6, mean_bl=0.05)
    topo, flat, g = util.make_flat(tree, gtr, 200, 6)
    eng = util.engine_for(flat, g)
    eng.marginal()
    tot, nd = eng.results()
    res = O.marginal(flat, g)
    assert abs(tot - res.total_LH) <= LH_RTOL * abs(res.total_LH)
    compare_all(flat, g, eng, res, every=3)


@pytest.mark.parametrize('alphabet,q,nw', [('aa_nogap', 20, None), ('aa', 22, None), ('aa_nogap', 20, '4'), ('aa_nogap', 20, '8'),
                                           ('aa', 22, '16')])
def test_large_alphabet_tensor_pipe_kernels(alphabet, q, nw, monkeypatch):
    """q >= 20 level kernels on the fp64 tensor pipe (csrc/ttb_mma.cuh): polytomies up to a 40-child star (power-of-two
    rescaling across the four lanes of a pattern), ambiguous characters, a ragged third tile, reconstructed tips, N_diff,
    every compiled warp count, and agreement with the one-thread-per-pattern kernels (TTB_NO_MMA=1)."""
    if nw:
        monkeypatch.setenv('TTB_MMA_NW', nw)
    gtr = util.random_gtr(alphabet, 21)
    assert gtr.n_states == q
    tree = synth.random_tree(90, seed=31, mean_bl=0.08, polytomy_frac=0.3)
    topo, flat, g = util.make_flat(tree, gtr, 300, 31, amb_frac=0.03, amb_chars='X' if alphabet == 'aa_nogap' else 'X-')
    Lp = flat['multiplicity'].shape[0]
    assert 256 < Lp and Lp % 128 != 0
    eng = util.engine_for(flat, g)
    eng.marginal()
    tot, nd = eng.results()
    res = O.marginal(flat, g)
    assert abs(tot - res.total_LH) <= LH_RTOL * abs(res.total_LH) and nd == res.N_diff
    compare_all(flat, g, eng, res, every=2)
    eng.marginal()
    assert eng.results() == (tot, 0)
    eng.marginal(reconstruct_tips=True)
    res_t = O.marginal(flat, g, reconstruct_tip_states=True)
    compare_all(flat, g, eng, res_t, reconstruct_tips=True, every=3)
    seq_mma, prof_mma = eng.all_seq_idx(), [eng.node_array(n, 2) for n in range(0, flat['parent'].shape[0], 7)]
    monkeypatch.setenv('TTB_NO_MMA', '1')
    old = util.engine_for(flat, g)
    old.marginal(reconstruct_tips=True)
    assert abs(old.results()[0] - tot) <= 1e-12 * abs(tot)
    prof_old = [old.node_array(n, 2) for n in range(0, flat['parent'].shape[0], 7)]
    assert max(np.abs(a - b).max() for a, b in zip(prof_mma, prof_old)) < 1e-10
    differ = seq_mma != old.all_seq_idx()
    assert differ.mean() < 1e-4          # exact ties aside, the two kernel families reconstruct the same sequences


def test_large_alphabet_star_tree_rescaling():
    """A 3 000-child star at q = 20: the running product of the up-messages leaves the double range many times; the
    tensor-pipe postorder rescales by exact powers of two per pattern (max over the four lanes that share it)."""
    from treetime_b200.tree import Node, Tree
    gtr = util.random_gtr('aa_nogap', 5)
    rng = np.random.default_rng(4)
    tree = Tree(Node(clades=[Node(name='t%06d' % i, branch_length=float(b)) for i, b in enumerate(rng.exponential(0.5, 3000))]))
    topo, flat, g = util.make_flat(tree, gtr, 140, 4)
    eng = util.engine_for(flat, g)
    eng.marginal()
    tot, nd = eng.results()
    res = O.marginal(flat, g)
    assert np.isfinite(tot) and abs(tot - res.total_LH) <= LH_RTOL * abs(res.total_LH)
    compare_all(flat, g, eng, res, every=500)


@pytest.mark.parametrize('approximate', [True, False])
def test_site_specific_gtr(approximate):
    """gtr_site_specific.py: per-site Pi/mu, interpolated (default) and exact exp(Qt)."""
    from treetime_b200.gtr import GTRSiteSpecific
    L = 300
    gtr = GTRSiteSpecific.random(L=L, alphabet='nuc', rng=np.random.default_rng(17))
    gtr.approximate = approximate
    tree = synth.random_tree(40, seed=12, mean_bl=0.05)
    topo, flat, g = util.make_flat(tree, gtr, L, 12, amb_frac=0.02, compress=False)
    assert flat['multiplicity'].shape[0] == L and g['site_specific']
    eng = util.engine_for(flat, g)
    eng.marginal()
    tot, nd = eng.results()
    res = O.marginal(flat, g)
    assert abs(tot - res.total_LH) <= LH_RTOL * abs(res.total_LH)
    assert nd == res.N_diff
    compare_all(flat, g, eng, res, every=2)
    eng.marginal(reconstruct_tips=True)
    res_t = O.marginal(flat, g, reconstruct_tip_states=True)
    compare_all(flat, g, eng, res_t, reconstruct_tips=True, every=3)
    # branch objective with the per-pattern matrices, below and above the interpolation range
    nodes = np.arange(1, flat['parent'].shape[0], 3, dtype=np.int32)
    for tval in (1e-3, 0.05, 0.7, 12.0 / g['rate_scale']):
        f = eng.branch_objective(nodes, np.full(nodes.shape[0], tval))
        ref = np.array([O.branch_objective(flat, g, res_t, n, tval) for n in nodes])
        assert np.allclose(f, ref, rtol=1e-10, atol=1e-9), (tval, np.abs(f - ref).max())
    print('site-specific approx=%s relLH=%.1e' % (approximate, abs(tot - res.total_LH) / abs(res.total_LH)))


def test_lh_only_and_new_rate():
    """optimize_gtr_rate's cost function: postorder + root only, for several mu."""
    tree = synth.random_tree(100, seed=8, mean_bl=0.01)
    topo, flat, g = util.make_flat(tree, util.nuc_gtr(), 400, 8)
    eng = util.engine_for(flat, g)
    for mu in (0.5, 1.0, 2.0):
        gg = dict(g); gg['mu'] = g['mu'] * mu
        eng.set_gtr(gg)
        eng.marginal(lh_only=True)
        tot, _ = eng.results()
        ref = O.sequence_LH_only(flat, gg)
        assert abs(tot - ref.total_LH) <= LH_RTOL * abs(ref.total_LH)


def test_branch_objective_hamming_counts():
    tree = synth.random_tree(60, seed=9, mean_bl=0.02)
    topo, flat, g = util.make_flat(tree, util.nuc_gtr(), 500, 9, amb_frac=0.02)
    eng = util.engine_for(flat, g)
    eng.marginal()
    eng.results()
    res = O.marginal(flat, g)
    G = O.make_gtr(g)
    nodes = np.arange(1, flat['parent'].shape[0], dtype=np.int32)
    for tval in (1e-4, 0.01, 0.3):
        f = eng.branch_objective(nodes, np.full(nodes.shape[0], tval))
        ref = np.array([O.branch_objective(flat, g, res, n, tval) for n in nodes])
        assert np.allclose(f, ref, rtol=1e-10, atol=1e-9), np.abs(f - ref).max()
    num, den = eng.branch_hamming(nodes)
    m = flat['multiplicity']
    ref = np.array([np.sum(m * np.sum(res.outgroup_LH[n] * res.subtree_LH[n], axis=1)) for n in nodes])
    assert np.allclose(num, ref, rtol=1e-12) and den == m.sum()
    # merged root branch (treeanc.py:1317-1326)
    n1 = flat['child_idx'][0]
    pp, pc = O.root_branch_profiles(flat, g, res)
    f = eng.branch_objective([n1], [0.05], kinds=[1])[0]
    assert np.isclose(f, G.prob_t_profiles((pp, pc), m, 0.05, return_log=True), rtol=1e-10)
    # substitution statistics (treeanc.py:1556-1572)
    n_ij, T_i = eng.mutation_counts()
    rn, rT = O.mutation_counts(flat, g, res)
    assert np.allclose(n_ij, rn.sum(axis=-1), rtol=1e-10, atol=1e-12)
    assert np.allclose(T_i, rT.sum(axis=-1), rtol=1e-10, atol=1e-12)


def _check_mutation_order(flat, full, root_idx, mn, mp, ms):
    """ttb_fetch_mutations returns exactly the (node, position, state) triples where an internal node differs from its
    parent, ordered by (node, position) -- laid out that way by the device, no sort on either side."""
    internal = np.nonzero(flat['tip_row'] < 0)[0]
    slot = {int(n): k for k, n in enumerate(internal)}
    en, ep, es = [], [], []
    for k, n in enumerate(internal):
        if n == 0:
            continue
        d = np.nonzero(full[k] != full[slot[int(flat['parent'][n])]])[0]
        en.append(np.full(d.shape[0], n)); ep.append(d); es.append(full[k][d])
    assert np.array_equal(root_idx, full[0])
    assert np.array_equal(mn, np.concatenate(en)) and np.array_equal(mp, np.concatenate(ep)) and np.array_equal(ms, np.concatenate(es))


def test_sparse_input_and_sparse_result():
    """ttb_set_patterns_sparse (reference row + differences) and ttb_fetch_mutations (root row +
    states that differ from the parent) carry the same information as the dense calls."""
    from treetime_b200.sparse import sparse_from_dense, expand_mutations
    from treetime_b200.engine import Engine
    tree = synth.random_tree(120, seed=21, mean_bl=0.004)
    topo, flat, g = util.make_flat(tree, util.nuc_gtr(), 900, 21, amb_frac=0.01)
    dense = util.engine_for(flat, g)
    dense.marginal()
    tot, nd = dense.results()
    ref, row, pos, code = sparse_from_dense(flat['tip_codes'])
    assert row.shape[0] < 0.2 * flat['tip_codes'].size
    sp = Engine(5)
    sp.set_tree(flat['parent'], flat['child_ptr'], flat['child_idx'], flat['tip_row'])
    sp.set_patterns_sparse(ref, row, pos, code, flat['code_profiles'], flat['multiplicity'])
    sp.set_gtr(g)
    sp.set_branch_lengths(flat['t'])
    sp.marginal()
    tot2, nd2 = sp.results()
    assert tot2 == tot and nd2 == nd
    assert np.array_equal(sp.site_lh(), dense.site_lh())
    full = dense.all_seq_idx()
    root_idx, mn, mp, ms = sp.mutations()
    assert np.array_equal(expand_mutations(flat['parent'], flat['tip_row'], root_idx, mn, mp, ms), full)
    # a too small buffer is reported and retried
    r2, mn2, mp2, ms2 = sp.mutations(max_n=3)
    assert np.array_equal(mn2, mn) and np.array_equal(ms2, ms)
    _check_mutation_order(flat, full, root_idx, mn, mp, ms)
    # more patterns than one sweep of the ordered write covers (16 x 256 positions per block iteration), ragged tail
    tree2 = synth.random_tree(30, seed=22, mean_bl=0.05)
    topo2, flat2, g2 = util.make_flat(tree2, util.nuc_gtr(), 9000, 22, amb_frac=0.01)
    assert flat2['multiplicity'].shape[0] > 4200 and flat2['multiplicity'].shape[0] % 16 != 0
    e2 = util.engine_for(flat2, g2)
    e2.marginal()
    e2.results()
    _check_mutation_order(flat2, e2.all_seq_idx(), *e2.mutations())
    from treetime_b200._lib import TTBError
    with pytest.raises(TTBError):
        sp.set_patterns_sparse(ref, np.array([10 ** 6], dtype=np.int32), np.array([0], dtype=np.int32), np.array([0], dtype=np.uint8),
                               flat['code_profiles'], flat['multiplicity'])


@pytest.mark.parametrize('kind', ['nuc', 'poly', 'aa'])
def test_joint_reconstruction(kind):
    """N2: joint (max-product) reconstruction, TreeAnc._ml_anc_joint (treeanc.py:934-1080)."""
    if kind == 'nuc':
        tree = synth.random_tree(150, seed=41, mean_bl=0.01); gtr = util.nuc_gtr(); L = 800; amb = 0.02
    elif kind == 'poly':
        tree = synth.random_tree(200, seed=42, mean_bl=0.005, polytomy_frac=0.5, zero_frac=0.3); gtr = util.nuc_gtr(); L = 500; amb = 0.0
    else:
        tree = synth.random_tree(50, seed=43, mean_bl=0.05); gtr = util.random_gtr('aa_nogap', 5); L = 200; amb = 0.0
    topo, flat, g = util.make_flat(tree, gtr, L, 41, amb_frac=amb)
    eng = util.engine_for(flat, g)
    for tips in (False, True):
        eng.joint(reconstruct_tips=tips)
        tot, nd = eng.results()
        res = O.joint(flat, g, reconstruct_tip_states=tips)
        assert abs(tot - res.total_LH) <= LH_RTOL * abs(res.total_LH)
        assert np.allclose(eng.site_lh(), res.sequence_LH, rtol=1e-11, atol=1e-9)
        n_nodes = flat['parent'].shape[0]
        internal = [n for n in range(n_nodes) if flat['tip_row'][n] < 0]
        dev = eng.all_seq_idx()
        ref = np.array([res.seq_idx[n] for n in internal])
        mism = (dev != ref).mean()
        assert mism < 2e-3, mism                              # only rounding-level ties may differ ...
        if mism:                                              # ... and then the assignment is an equally good optimum
            seqs = [None] * n_nodes
            for k, n in enumerate(internal):
                seqs[n] = dev[k].astype(int)
            assert np.allclose(util.joint_assignment_lh(flat, g, seqs), res.sequence_LH, rtol=1e-10, atol=1e-8)
        if tips:
            tipn = [n for n in range(n_nodes) if flat['tip_row'][n] >= 0][:20]
            dt = eng.seq_idx(tipn)
            assert np.mean(dt != np.array([res.seq_idx[n] for n in tipn])) < 5e-3
    # second pass: nothing changes; marginal accessors are blocked until a marginal pass is run again
    eng.joint(reconstruct_tips=True)
    assert eng.results()[1] == 0
    from treetime_b200._lib import TTBError
    with pytest.raises(TTBError):
        eng.node_array(3, 2)
    lx = eng.node_array(0, 3)
    assert np.allclose(lx, res.joint_Lx[0], rtol=1e-11, atol=1e-9)
    eng.marginal()
    assert abs(eng.results()[0] - O.marginal(flat, g).total_LH) <= LH_RTOL * abs(O.marginal(flat, g).total_LH)
    # a fresh engine: joint first, then marginal -- the joint states are the marginal pass's "previous" states
    e2 = util.engine_for(flat, g)
    e2.joint()
    rj = O.joint(flat, g)
    e2.marginal()
    nd_ref = O.marginal(flat, g, prev_seq_idx=rj.seq_idx).N_diff
    assert abs(e2.results()[1] - nd_ref) <= 2e-3 * nd_ref + 4


@pytest.mark.parametrize('kind', ['nuc', 'aa'])
def test_branch_state_pairs(kind):
    """N2: ttb_branch_state_pairs against the oracle's restatement of the GTR.state_pair counting."""
    if kind == 'nuc':
        tree = synth.random_tree(120, seed=51, mean_bl=0.02, polytomy_frac=0.2); gtr = util.nuc_gtr(); L = 3000; amb = 0.03
    else:
        tree = synth.random_tree(40, seed=52, mean_bl=0.05); gtr = util.random_gtr('aa_nogap', 7); L = 500; amb = 0.0
    topo, flat, g = util.make_flat(tree, gtr, L, 51, amb_frac=amb)
    eng = util.engine_for(flat, g)
    nodes = np.arange(1, flat['parent'].shape[0], dtype=np.int32)
    for tips in (False, True):
        eng.joint(reconstruct_tips=tips)
        eng.results()
        n_nodes = flat['parent'].shape[0]
        seqs = [None] * n_nodes
        internal = [n for n in range(n_nodes) if flat['tip_row'][n] < 0]
        for k, row in zip(internal, eng.all_seq_idx()):
            seqs[k] = row
        if tips:
            tipn = [n for n in range(n_nodes) if flat['tip_row'][n] >= 0]
            for k, row in zip(tipn, eng.seq_idx(tipn)):
                seqs[k] = row
        C, F = eng.branch_state_pairs(nodes, tip_states=tips)
        Cr, Fr = O.branch_pair_tables(flat, seqs, tip_states=tips)       # same sequences: bit-exact integer work
        assert np.array_equal(C, Cr) and np.array_equal(F, Fr)
        assert np.allclose(C.sum(axis=(1, 2)), flat['multiplicity'].sum())
    # a subset, after a marginal pass
    eng.marginal()
    eng.results()
    sub = nodes[::7]
    C2, F2 = eng.branch_state_pairs(sub)
    seqs = [None] * n_nodes
    for k, row in zip(internal, eng.all_seq_idx()):
        seqs[k] = row
    Cr, Fr = O.branch_pair_tables(flat, seqs)
    assert np.array_equal(C2, Cr[sub - 1]) and np.array_equal(F2, Fr[sub - 1])
    from treetime_b200._lib import TTBError
    with pytest.raises(TTBError):
        eng.branch_state_pairs(nodes, tip_states=True)      # the marginal pass did not reconstruct tips
    with pytest.raises(TTBError):
        eng.branch_state_pairs([0])


@pytest.mark.parametrize('kind', ['nuc', 'aa', 'ss'])
def test_seqgen(kind):
    """N4: ttb_seqgen.  With caller-supplied uniforms the kernel must equal the oracle's restatement of
    SeqGen.evolve (seqgen.py:38-67) state for state; with the device's Philox stream the transition
    frequencies must follow exp(Qt) and the run must be reproducible from the seed."""
    from treetime_b200.flatten import code_table
    from treetime_b200.gtr import GTRSiteSpecific
    if kind == 'nuc':
        tree = synth.random_tree(60, seed=71, mean_bl=0.2, polytomy_frac=0.2); gtr = util.nuc_gtr(); L = 2000
    elif kind == 'aa':
        tree = synth.random_tree(20, seed=72, mean_bl=0.3); gtr = util.random_gtr('aa_nogap', 9); L = 700
    else:
        L = 900
        tree = synth.random_tree(30, seed=73, mean_bl=0.2); gtr = GTRSiteSpecific.random(L=L, alphabet='nuc', rng=np.random.default_rng(8))
    topo, flat, g = util.make_flat(tree, gtr, L, 71, compress=False)
    assert flat['multiplicity'].shape[0] == L
    eng = util.engine_for(flat, g)
    chars, lut, table = code_table(gtr.profile_map, gtr.n_states)
    s2c = np.array([lut[str(c)] for c in gtr.alphabet], dtype=np.uint8)
    n_nodes = flat['parent'].shape[0]
    rng = np.random.default_rng(5)
    U = rng.random((n_nodes, L))
    G_ = O.make_gtr(g)
    got = eng.seqgen(0, s2c, uniforms=U)
    want = O.seqgen(flat, G_, U)
    assert (got != want).mean() < 2e-5          # device exp(Qt) differs from numpy's by ulps: a draw on a boundary may flip
    root = rng.integers(0, gtr.n_states, size=L).astype(np.uint8)
    got = eng.seqgen(0, s2c, root_idx=root, uniforms=U)
    assert (got[0] == root).all() and (got != O.seqgen(flat, G_, U, root_idx=root)).mean() < 2e-5
    # the tips became the alignment: a reconstruction on it equals the oracle's on the same codes
    tips = flat['tip_row'] >= 0
    f2 = dict(flat, tip_codes=s2c[got[tips]][np.argsort(flat['tip_row'][tips])])
    eng.marginal()
    tot, _ = eng.results()
    ref = O.marginal(f2, g)
    assert abs(tot - ref.total_LH) <= LH_RTOL * abs(ref.total_LH)
    # Philox stream: reproducible, seed-dependent, right distribution
    a = eng.seqgen(1234, s2c)
    b = eng.seqgen(1234, s2c)
    c = eng.seqgen(1235, s2c)
    assert (a == b).all() and (a != c).mean() > 0.05
    q = gtr.n_states
    if kind != 'ss':
        n = int(np.argmax(flat['t'][1:])) + 1          # longest branch: most transitions
        P = G_.expQt(flat['t'][n])
        pa, ch = a[flat['parent'][n]].astype(int), a[n].astype(int)
        cnt = np.zeros((q, q)); np.add.at(cnt, (ch, pa), 1)
        exp = P * np.bincount(pa, minlength=q)[None, :]
        ok = exp > 5
        chi2 = ((cnt - exp)[ok] ** 2 / exp[ok]).sum()
        dof = ok.sum() - (ok.any(axis=0)).sum()
        assert chi2 < dof + 6 * np.sqrt(2 * dof) + 10, (chi2, dof)
        piroot = np.bincount(a[0], minlength=q) / L
        assert np.abs(piroot - g['Pi']).max() < 5 * np.sqrt(0.25 / L)
    else:
        # per-site models: compare the mean log-probability of the drawn transitions with its expectation
        n = int(np.argmax(flat['t'][1:])) + 1
        P = G_.expQt(flat['t'][n])                    # (q, q, L)
        pa, ch = a[flat['parent'][n]].astype(int), a[n].astype(int)
        site = np.arange(L)
        lp = np.log(np.maximum(P[ch, pa, site], 1e-300))
        col = np.maximum(P[:, pa, site], 1e-300)
        mean = (col * np.log(col)).sum(axis=0)
        var = (col * np.log(col) ** 2).sum(axis=0) - mean ** 2
        z = (lp.sum() - mean.sum()) / np.sqrt(var.sum())
        assert abs(z) < 5, z
    from treetime_b200._lib import TTBError
    with pytest.raises(TTBError):
        eng.seqgen(1, np.full(q, 255, dtype=np.uint8))


@pytest.mark.parametrize('kind', ['nuc', 'aa', 'ss'])
def test_sample_states(kind):
    """sample_from_profile=True (treeanc.py:919-923, seq_utils.py:266-269): states drawn on the device from the
    marginal profiles with the caller's uniforms equal the oracle's draws (bit-exact except where a cumulative sum
    lies within rounding of its uniform), N_diff counts against the previous pass."""
    rng = np.random.default_rng(101)
    if kind == 'ss':
        from treetime_b200.gtr import GTRSiteSpecific
        gtr = GTRSiteSpecific.random(L=333, alphabet='nuc', rng=np.random.default_rng(5))
        tree = synth.random_tree(30, seed=9, mean_bl=0.05)
        topo, flat, g = util.make_flat(tree, gtr, 333, 9, amb_frac=0.02, compress=False)
    else:
        gtr = util.nuc_gtr() if kind == 'nuc' else util.random_gtr('aa', 3)
        tree = synth.random_tree(45, seed=8, mean_bl=0.08)
        topo, flat, g = util.make_flat(tree, gtr, 500, 8, amb_frac=0.03)
    n_nodes = flat['parent'].shape[0]
    L = flat['multiplicity'].shape[0]
    eng = util.engine_for(flat, g)
    prev = None
    for tips in (False, True, True):
        nodes = [n for n in range(1, n_nodes) if tips or flat['tip_row'][n] < 0]
        U = rng.random((len(nodes), L))
        uni = {n: U[k] for k, n in enumerate(nodes)}
        res = O.marginal(flat, g, reconstruct_tip_states=tips, prev_seq_idx=prev, uniforms=uni)
        eng.marginal(reconstruct_tips=tips, keep_prev=True)
        eng.results()
        nd, nd_tips = eng.sample_states(nodes, U)
        got = eng.seq_idx(nodes)
        near = 0
        for k, n in enumerate(nodes):
            bad = got[k] != res.seq_idx[n]
            if bad.any():    # only where a running sum is within rounding of the uniform
                cum = np.cumsum(res.profile[n], axis=1)
                assert (np.abs(cum - U[k][:, None]).min(axis=1)[bad] < 1e-12).all(), n
                near += int(bad.sum())
        assert near <= 2
        if prev is not None:
            exp_int = sum(int((res.seq_idx[n] != prev[n]).sum()) for n in nodes if flat['tip_row'][n] < 0 and prev[n] is not None)
            exp_tip = sum(int((res.seq_idx[n] != prev[n]).sum()) for n in nodes if flat['tip_row'][n] >= 0 and prev[n] is not None)
            assert abs(nd - exp_int) <= near
            if all(prev[n] is not None for n in nodes):
                assert abs(nd_tips - exp_tip) <= near
        assert (got != np.array([res.profile[n].argmax(axis=1) for n in nodes])).any()   # really sampled
        # the root keeps its argmax; the profiles are those of the plain pass
        assert np.abs(eng.node_array(nodes[0], 2) - res.profile[nodes[0]]).max() < PROF_ATOL
        prev = list(res.seq_idx)
    # the uniforms of many nodes travel in several blocks: same states whatever the block size
    import os
    eng.marginal(reconstruct_tips=True, keep_prev=True)
    eng.results()
    eng.sample_states(nodes, U)
    whole = eng.seq_idx(nodes)
    os.environ['TTB_SAMPLE_BLOCK_DOUBLES'] = str(3 * L + 1)          # 3 nodes per block
    try:
        eng.marginal(reconstruct_tips=True, keep_prev=True)
        eng.results()
        nd2, ndt2 = eng.sample_states(nodes, U)
    finally:
        del os.environ['TTB_SAMPLE_BLOCK_DOUBLES']
    assert np.array_equal(eng.seq_idx(nodes), whole) and nd2 == 0 and ndt2 == 0
    from treetime_b200._lib import TTBError
    eng.marginal()
    with pytest.raises(TTBError):
        eng.sample_states(nodes[:1], U[:1])       # the last pass did not keep the previous states


@pytest.mark.parametrize('kind', ['nuc', 'aa', 'ss', 'ss_exact'])
def test_mutation_counts_per_site(kind):
    """A10 per pattern: n_ija (q,q,L') and T_ia (q,L') of infer_gtr (treeanc.py:1551-1572) before the sum over
    patterns, single and site-specific current models; the totals equal ttb_mutation_counts."""
    if kind.startswith('ss'):
        from treetime_b200.gtr import GTRSiteSpecific
        gtr = GTRSiteSpecific.random(L=260, alphabet='nuc', rng=np.random.default_rng(23))
        gtr.approximate = kind == 'ss'
        tree = synth.random_tree(35, seed=14, mean_bl=0.06)
        topo, flat, g = util.make_flat(tree, gtr, 260, 14, amb_frac=0.02, compress=False)
    else:
        gtr = util.nuc_gtr() if kind == 'nuc' else util.random_gtr('aa', 4)
        tree = synth.random_tree(50, seed=15, mean_bl=0.05)
        topo, flat, g = util.make_flat(tree, gtr, 400, 15, amb_frac=0.02)
    eng = util.engine_for(flat, g)
    eng.marginal()
    eng.results()
    res = O.marginal(flat, g)
    n_ija, T_ia = O.mutation_counts(flat, g, res)
    a, b = eng.mutation_counts_per_site()
    assert a.shape == n_ija.shape and b.shape == T_ia.shape
    assert np.allclose(a, n_ija, rtol=1e-9, atol=1e-12), np.abs(a - n_ija).max()
    assert np.allclose(b, T_ia, rtol=1e-9, atol=1e-14), np.abs(b - T_ia).max()
    if not kind.startswith('ss'):
        n_ij, T_i = eng.mutation_counts()
        assert np.allclose(n_ij, a.sum(axis=-1), rtol=1e-11) and np.allclose(T_i, b.sum(axis=-1), rtol=1e-11)


def test_api_errors():
    from treetime_b200.engine import Engine
    from treetime_b200._lib import TTBError
    with pytest.raises(TTBError):
        Engine(9)            # no kernels for 9 states
    eng = Engine(5)
    with pytest.raises(TTBError):
        eng.marginal()       # nothing set


@pytest.mark.parametrize('kind', ['nuc', 'aa', 'ss'])
def test_branch_masks(kind):
    """Per-branch masks (ARG mode, treeanc.py:862-872,909-917,1294,1326-1333,1564-1572): masked up-messages are
    dropped, masked children keep their subtree profile, masked patterns leave the branch's likelihood and
    substitution statistics -- against the oracle with the same masks."""
    if kind == 'ss':
        from treetime_b200.gtr import GTRSiteSpecific
        gtr = GTRSiteSpecific.random(L=300, alphabet='nuc', rng=np.random.default_rng(31))
        tree = synth.random_tree(33, seed=16, mean_bl=0.05)
        topo, flat, g = util.make_flat(tree, gtr, 300, 16, amb_frac=0.02, compress=False)
    else:
        gtr = util.nuc_gtr() if kind == 'nuc' else util.random_gtr('aa_nogap', 6)
        tree = synth.random_tree(48, seed=17, mean_bl=0.04)
        topo, flat, g = util.make_flat(tree, gtr, 420, 17, amb_frac=0.02 if kind == 'nuc' else 0.0)
    n_nodes = flat['parent'].shape[0]
    L = flat['multiplicity'].shape[0]
    rng = np.random.default_rng(7)
    M = np.ones((3, L), dtype=np.uint8)
    M[0, L // 2:] = 0                       # segment mask
    M[2] = rng.random(L) < 0.7              # arbitrary 0/1 pattern
    node_mask = rng.integers(-1, 3, size=n_nodes).astype(np.int32)
    masks = {n: M[k].astype(float) for n, k in enumerate(node_mask) if k >= 0}
    eng = util.engine_for(flat, g)
    eng.set_branch_masks(M, node_mask)
    for tips in (False, True):
        eng.marginal(reconstruct_tips=tips)
        tot, nd = eng.results()
        res = O.marginal(flat, g, reconstruct_tip_states=tips, masks=masks)
        assert abs(tot - res.total_LH) <= LH_RTOL * abs(res.total_LH)
        compare_all(flat, g, eng, res, reconstruct_tips=tips)
    # branch objective / hamming numerators with masked multiplicities, incl. the merged root branch
    import oracle_engine
    oe = oracle_engine.OracleEngine(g['Pi'].shape[0])
    oe.set_tree(flat['parent'], flat['child_ptr'], flat['child_idx'], flat['tip_row'])
    oe.set_patterns(flat['tip_codes'], flat['code_profiles'], flat['multiplicity'])
    oe.set_gtr(g); oe.set_branch_lengths(flat['t']); oe.set_branch_masks(M, node_mask)
    oe.marginal(reconstruct_tips=True)
    nodes = np.arange(1, n_nodes, 2, dtype=np.int32)
    kinds = np.zeros(nodes.shape[0], dtype=np.int32)
    if flat['child_ptr'][1] - flat['child_ptr'][0] == 2 and kind != 'ss':
        nodes = np.append(nodes, flat['child_idx'][flat['child_ptr'][0]]).astype(np.int32)
        kinds = np.append(kinds, 1).astype(np.int32)
    for tval in (1e-3, 0.08):
        f = eng.branch_objective(nodes, np.full(nodes.shape[0], tval), kinds)
        ref = oe.branch_objective(nodes, np.full(nodes.shape[0], tval), kinds)
        assert np.allclose(f, ref, rtol=1e-10, atol=1e-9), np.abs(f - ref).max()
    num, den = eng.branch_hamming(nodes, kinds)
    rnum, rden = oe.branch_hamming(nodes, kinds)
    assert np.allclose(num, rnum, rtol=1e-11) and np.isclose(den, rden)
    a, b = eng.mutation_counts_per_site()
    ra, rb = oe.mutation_counts_per_site()
    assert np.allclose(a, ra, rtol=1e-9, atol=1e-12) and np.allclose(b, rb, rtol=1e-9, atol=1e-14)
    if kind != 'ss':
        n_ij, T_i = eng.mutation_counts()
        assert np.allclose(n_ij, ra.sum(axis=-1), rtol=1e-10) and np.allclose(T_i, rb.sum(axis=-1), rtol=1e-10)
    # masks off again: back to the plain kernels
    eng.set_branch_masks(None, None)
    eng.marginal()
    tot, _ = eng.results()
    res = O.marginal(flat, g)
    assert abs(tot - res.total_LH) <= LH_RTOL * abs(res.total_LH)
    from treetime_b200._lib import TTBError
    eng.set_branch_masks(M, node_mask)
    if kind != 'ss':
        with pytest.raises(TTBError):
            eng.joint()
    with pytest.raises(TTBError):
        eng.set_branch_masks(M * 2, node_mask)


@pytest.mark.parametrize('kind', ['nuc', 'aa', 'site_specific', 'tips', 'joint'])
def test_merged_level_launches(kind, monkeypatch):
    """Runs of small levels are ONE launch whose blocks walk several levels (build_groups / Pipe::wait_done): every
    merge threshold -- none, a few levels, the whole tree above the leaf level -- gives the same messages bit for bit
    (the per-node arithmetic does not change; only the grouping of the log-prefactor sums does)."""
    rt = False
    if kind == 'aa':
        gtr, L, compress = util.random_gtr('aa_nogap', 5), 300, True
        monkeypatch.setenv('TTB_NO_MMA', '1')   # merged launches are a variant of the one-thread-per-pattern kernels
    elif kind == 'site_specific':
        from treetime_b200.gtr import GTRSiteSpecific
        L, compress = 700, False
        gtr = GTRSiteSpecific.random(L=L, alphabet='nuc', rng=np.random.default_rng(3))
    else:
        gtr, L, compress = util.nuc_gtr(), 1500, True
        rt = kind == 'tips'
    tree = synth.random_tree(400, seed=8, mean_bl=0.01, polytomy_frac=0.3)
    topo, flat, g = util.make_flat(tree, gtr, L, 8, amb_frac=0.01, amb_chars='N' if kind == 'aa' else 'N-RY', compress=compress)
    n_nodes = flat['parent'].shape[0]
    base = None
    for merge in ('0', '3', '24', '100000'):
        monkeypatch.setenv('TTB_MERGE_NODES', merge)
        eng = util.engine_for(flat, g)
        if kind == 'joint':
            eng.joint()
            tot, nd = eng.results()
            cur = (tot, eng.all_seq_idx(), eng.site_lh())
            if base is None:
                base = cur
                jres = O.joint(flat, g)
                assert abs(tot - jres.total_LH) <= LH_RTOL * abs(jres.total_LH)
            else:
                assert abs(cur[0] - base[0]) <= 1e-12 * abs(base[0]) and np.array_equal(cur[1], base[1]) and np.array_equal(cur[2], base[2])
            continue
        launches0 = eng.launch_count()
        eng.marginal(reconstruct_tips=rt)
        tot, nd = eng.results()
        n_launch = eng.launch_count() - launches0
        prof = [eng.node_array(n, 2) for n in range(0, n_nodes, 9) if flat['tip_row'][n] < 0 or rt]
        sub = [eng.node_array(n, 0) for n in range(0, n_nodes, 9)]
        cur = (tot, eng.all_seq_idx(), prof, sub, n_launch)
        if base is None:
            base = cur
            res = O.marginal(flat, g, reconstruct_tip_states=rt)
            assert abs(tot - res.total_LH) <= LH_RTOL * abs(res.total_LH) and nd == res.N_diff
            compare_all(flat, g, eng, res, reconstruct_tips=rt, every=11)
        else:
            assert abs(cur[0] - base[0]) <= 1e-12 * abs(base[0])
            assert np.array_equal(cur[1], base[1])
            assert all(np.array_equal(x, y) for x, y in zip(cur[2], base[2]))
            assert all(np.array_equal(x, y) for x, y in zip(cur[3], base[3]))
            # fewer launches than one per level (site-specific models, masks and float storage keep one launch per level)
            assert cur[4] < base[4] if kind != 'site_specific' else cur[4] == base[4]
        eng.marginal(reconstruct_tips=rt)
        assert eng.results()[1] == 0


@pytest.mark.parametrize('kind', ['nuc', 'site_specific', 'tips'])
def test_float_message_storage(kind):
    """ttb_set_message_storage(TTB_STORAGE_F32): S / M stored as float, arithmetic in double -- an opt-in whose error is
    MEASURED here and pinned (north_star: fp64 vs fp32 is decided from measured error): total log-LH 1e-7 relative
    (fp64 storage: 1e-9 bar, ~1e-16 measured), profiles 1e-6, sequences identical wherever the two largest profile
    entries are more than 1e-6 apart.  Switching back to double reproduces the fp64 result bit for bit."""
    rt = kind == 'tips'
    if kind == 'site_specific':
        from treetime_b200.gtr import GTRSiteSpecific
        L, compress = 900, False
        gtr = GTRSiteSpecific.random(L=L, alphabet='nuc', rng=np.random.default_rng(4))
    else:
        gtr, L, compress = util.nuc_gtr(), 3000, True
    tree = synth.random_tree(600, seed=9, mean_bl=0.004, polytomy_frac=0.1)
    topo, flat, g = util.make_flat(tree, gtr, L, 9, amb_frac=0.01, compress=compress)
    res = O.marginal(flat, g, reconstruct_tip_states=rt)
    eng = util.engine_for(flat, g)
    eng.marginal(reconstruct_tips=rt)
    tot64, _ = eng.results()
    idx64 = eng.all_seq_idx()
    prof64 = eng.node_array(3, 2) if flat['tip_row'][3] < 0 or rt else eng.node_array(0, 2)
    bytes64 = eng.device_bytes()
    eng.set_message_storage('f32')
    with pytest.raises(Exception):
        eng.node_array(0, 2)                        # the reconstruction was invalidated
    eng.marginal(reconstruct_tips=rt)
    tot32, nd = eng.results()
    assert eng.device_bytes() < 0.62 * bytes64
    rel = abs(tot32 - res.total_LH) / abs(res.total_LH)
    assert rel <= 1e-7, rel
    assert np.allclose(eng.site_lh(), res.sequence_LH, rtol=1e-5, atol=1e-5)     # per pattern; the total averages the rounding out
    worst = 0.0
    internal = np.nonzero(flat['tip_row'] < 0)[0]
    idx32 = eng.all_seq_idx()
    flips = 0
    for k, n in enumerate(internal):
        if k % 5 == 0:
            worst = max(worst, np.abs(eng.node_array(int(n), 2) - res.profile[n]).max(), np.abs(eng.node_array(int(n), 0) - res.subtree_LH[n]).max())
            if n:
                worst = max(worst, np.abs(eng.node_array(int(n), 1) - res.outgroup_LH[n]).max())
        bad = idx32[k] != res.seq_idx[n]
        if bad.any():
            flips += int(bad.sum())
            assert util.tie_mask(res.profile[n], rel=1e-6)[bad].all(), 'sequence differs where the top-2 gap exceeds 1e-6 (node %d)' % n
    assert worst < PROF_ATOL, worst
    f = eng.branch_objective(np.arange(1, 40, dtype=np.int32), np.full(39, 0.01))          # run-time typed readers
    ref = np.array([O.branch_objective(flat, g, res, n, 0.01) for n in range(1, 40)])
    assert np.allclose(f, ref, rtol=1e-5, atol=1e-5)
    with pytest.raises(Exception):
        eng.joint()
    eng.set_message_storage('f64')
    eng.marginal(reconstruct_tips=rt)
    assert eng.results()[0] == tot64 and np.array_equal(eng.all_seq_idx(), idx64)
    assert np.array_equal(prof64, eng.node_array(3, 2) if flat['tip_row'][3] < 0 or rt else eng.node_array(0, 2))
    print('float storage (%s): rel dLH %.1e, max|dprofile| %.1e, %d of %d states differ (all within 1e-6 of a tie)'
          % (kind, rel, worst, flips, idx32.size))


@pytest.mark.parametrize('kind', ['nuc', 'site_specific'])
def test_device_brent_matches_host_lockstep(kind):
    """ttb_brent_* (A8 on the device): the Brent state machine of scipy / treetime_b200.brent with its state on the
    device takes the same iterates as the host lock-step version driven by ttb_branch_objective -- all branches of the
    tree incl. the merged branch across a bifurcating root, tight and loose tolerances."""
    from treetime_b200.brent import brent_lockstep, BracketError
    from treetime_b200 import config as ttconf
    if kind == 'site_specific':
        from treetime_b200.gtr import GTRSiteSpecific
        L, compress = 400, False
        gtr = GTRSiteSpecific.random(L=L, alphabet='nuc', rng=np.random.default_rng(6))
    else:
        gtr, L, compress = util.nuc_gtr(), 1200, True
    tree = synth.random_tree(150, seed=14, mean_bl=0.01)
    topo, flat, g = util.make_flat(tree, gtr, L, 14, amb_frac=0.01, compress=compress)
    eng = util.engine_for(flat, g)
    eng.marginal()
    n_nodes = flat['parent'].shape[0]
    root_kids = flat['child_idx'][flat['child_ptr'][0]:flat['child_ptr'][1]]
    fids = np.array([n for n in range(1, n_nodes) if n not in root_kids] + [int(root_kids[0])], dtype=np.int32)
    kinds = np.zeros(fids.shape[0], dtype=np.int32); kinds[-1] = 1
    num, den = eng.branch_hamming(fids, kinds)
    xb = np.sqrt(1 - num / den)
    smax = np.sqrt(ttconf.MAX_BRANCH_LENGTH)
    n = fids.shape[0]
    for tol in (1e-10, 1e-2):
        calls = [0]

        def neg_prob(idx, s):
            calls[0] += 1
            return -1.0 * eng.branch_objective(fids[idx], s ** 2, kinds[idx]) + np.exp(s ** 4 / 10000)
        host = brent_lockstep(neg_prob, np.full(n, -smax), xb, np.full(n, smax), tol=tol)
        launches0 = eng.launch_count()
        dev = eng.brent_minimize(fids, kinds, np.full(n, -smax), xb, np.full(n, smax), tol)
        assert eng.launch_count() - launches0 <= 3 * (calls[0] + 8)
        assert dev['success'].all() and host['success'].all()
        # the objective values differ in the last bits (other block partition of the pattern sum): same minimum to
        # Brent's own resolution, same iteration counts almost everywhere
        # (a minimum located from function values is only defined to ~sqrt(eps) relative)
        assert np.allclose(dev['x'] ** 2, host['x'] ** 2, rtol=max(10 * tol, 5e-6), atol=1e-10)
        assert np.allclose(dev['fun'], host['fun'], rtol=1e-12)
        if tol >= 1e-6:          # below sqrt(eps) the last iterations are driven by the rounding of the function values
            assert np.mean(dev['nit'] == host['nit']) > 0.95 and np.mean(dev['nfev'] == host['nfev']) > 0.95
        else:
            assert abs(np.median(dev['nit'] - host['nit'])) <= 2
    with pytest.raises(BracketError):
        eng.brent_minimize(fids, kinds, np.full(n, -smax), np.full(n, 2 * smax), np.full(n, smax), 1e-8)
