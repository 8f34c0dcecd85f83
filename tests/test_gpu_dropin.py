"""GPU: the drop-in (`accelerate(treetime.TreeAnc / ClockTree / TreeTime)`) on the real CUDA engine next to the
UNMODIFIED reference classes, in one process.  The reference is /root/reference in the build container and the
staged copy oracle/_ref (oracle/stage_ref.py) on the GPU box; Biopython is replaced by oracle/bioshim.

Tolerances (BASELINE.json north_star): total log-LH 1e-9 relative, profiles 1e-6, sequences identical except at
exact ties; branch lengths located by Brent from function values: 1e-6 relative."""
import numpy as np
import pytest

import refenv

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not refenv.available(), reason='reference not staged (python oracle/stage_ref.py)')]

LH_RTOL = 1e-9
PROF_ATOL = 1e-6
PI = np.array([.3, .2, .2, .29, .01])


def _inputs(seed, n, L, mean_bl=0.01, amb=0.02, zero_frac=0.0):
    from treetime_b200 import synth
    from treetime_b200.gtr import GTR
    T = synth.random_tree(n, seed=seed, mean_bl=mean_bl, zero_frac=zero_frac)
    g = GTR.custom(pi=PI.copy(), W=np.ones((5, 5)), alphabet='nuc')
    idx = synth.evolve_alignment(T, L, g.Pi, g.W, seed=seed)
    aln = {k: g.alphabet[v] for k, v in idx.items()}
    if amb:
        aln = synth.sprinkle_ambiguous(aln, amb, 'N-R', seed=3)
    return T, aln


def _bio(T, aln):
    from io import StringIO
    from Bio import Phylo
    from Bio.Align import MultipleSeqAlignment
    from Bio.SeqRecord import SeqRecord
    from Bio.Seq import Seq
    return (Phylo.read(StringIO(T.to_newick()), 'newick'),
            MultipleSeqAlignment([SeqRecord(Seq(''.join(aln[k])), id=k, name=k, description='') for k in aln]))


def _ref_gtr():
    from treetime import GTR as RG
    return RG.custom(pi=PI.copy(), W=np.ones((5, 5)), alphabet='nuc')


def _same_sequences(a, b):
    import util
    bad = a.cseq != b.cseq
    if bad.any():
        assert util.tie_mask(a.marginal_profile)[bad].all(), a.name


def _pair(seed=33, n=60, L=700, **kw):
    refenv.activate()
    from treetime import TreeAnc as RefTreeAnc
    from treetime_b200.dropin import accelerate
    T, aln = _inputs(seed, n, L)
    t1, a1 = _bio(T, aln)
    t2, a2 = _bio(T, aln)
    rt = RefTreeAnc(tree=t1, aln=a1, gtr=_ref_gtr(), rng_seed=1, verbose=0, **kw)
    dt = accelerate(RefTreeAnc)(tree=t2, aln=a2, gtr=_ref_gtr(), rng_seed=1, verbose=0, **kw)
    return rt, dt


def _compare_reconstruction(rt, dt, tips=False):
    assert dt._b200_live and type(dt._engine).__module__ == 'treetime_b200.engine'      # the CUDA engine ran
    assert abs(rt.sequence_LH() - dt.sequence_LH()) <= LH_RTOL * abs(rt.sequence_LH())
    assert np.allclose(rt.tree.sequence_LH, dt.tree.sequence_LH, rtol=1e-11, atol=1e-11)
    worst = 0.0
    for a, b in zip(rt.tree.find_clades(), dt.tree.find_clades()):
        assert a.name == b.name
        if tips or not a.is_terminal():
            worst = max(worst, np.abs(a.marginal_profile - b.marginal_profile).max())
            _same_sequences(a, b)
        if a.up is not None:
            pa, pb = rt.marginal_branch_profile(a), dt.marginal_branch_profile(b)
            worst = max(worst, np.abs(pa[0] - pb[0]).max(), np.abs(pa[1] - pb[1]).max())
    assert worst < PROF_ATOL
    return worst


def test_gpu_dropin_treeanc_against_the_reference_class():
    """class B200TreeAnc(B200MarginalMixin, treetime.TreeAnc) vs treetime.TreeAnc, same inputs, same accessors."""
    rt, dt = _pair()
    assert rt.infer_ancestral_sequences(marginal=True) == dt.infer_ancestral_sequences(marginal=True)
    launches = dt._engine.launch_count()
    assert launches > 0
    worst = _compare_reconstruction(rt, dt)
    assert rt.infer_ancestral_sequences(marginal=True) == dt.infer_ancestral_sequences(marginal=True) == 0
    assert rt.infer_ancestral_sequences(marginal=True, reconstruct_tip_states=True) == \
        dt.infer_ancestral_sequences(marginal=True, reconstruct_tip_states=True)
    _compare_reconstruction(rt, dt, tips=True)
    rn, dn = list(rt.tree.find_clades()), list(dt.tree.find_clades())
    for k in (3, 7, 20, len(rn) - 1):
        assert np.allclose(rt.get_branch_mutation_matrix(rn[k]), dt.get_branch_mutation_matrix(dn[k]), rtol=1e-9, atol=1e-12)
        x, y = rt.optimal_marginal_branch_length(rn[k]), dt.optimal_marginal_branch_length(dn[k])
        assert abs(x - y) <= 1e-6 * x + 1e-12
    ra, da = rt.get_reconstructed_alignment(), dt.get_reconstructed_alignment()
    assert sum(str(r.seq) != str(d.seq) for r, d in zip(ra, da)) == 0
    rt.optimize_tree(branch_length_mode='marginal', max_iter=2, infer_gtr=True, prune_short=True)
    dt.optimize_tree(branch_length_mode='marginal', max_iter=2, infer_gtr=True, prune_short=True)
    assert dt._b200_live and dt._engine.launch_count() > launches
    a = np.array([c.branch_length for c in rt.tree.find_clades()]); b = np.array([c.branch_length for c in dt.tree.find_clades()])
    assert a.shape == b.shape and np.allclose(a[1:], b[1:], rtol=1e-6, atol=1e-12)
    assert abs(rt.sequence_LH() - dt.sequence_LH()) <= LH_RTOL * abs(rt.sequence_LH())
    assert np.allclose(rt.gtr.W, dt.gtr.W, rtol=1e-7) and np.allclose(rt.gtr.Pi, dt.gtr.Pi, rtol=1e-7)
    assert type(dt.gtr) is type(rt.gtr)
    rt.optimize_gtr_rate(); dt.optimize_gtr_rate()
    assert np.isclose(rt.gtr.mu, dt.gtr.mu, rtol=1e-6)
    # after optimize_gtr_rate the device holds one consistent pass (advisor finding, round 1)
    n = dn[5]
    pp, pc = dt.marginal_branch_profile(n)
    assert np.all(np.isfinite(pp)) and np.allclose(pp.sum(axis=1), 1.0, atol=1e-12)
    print('drop-in TreeAnc on the device: max|dprofile| = %.1e' % worst)


def test_gpu_dropin_n_diff_after_prune_and_same_shape_relabel():
    """N_diff across prune_short_branches (reference semantics, treeanc.py:925-926) and the re-upload of the tip rows
    when the topology arrays are unchanged but the tips moved (advisor findings, round 1)."""
    refenv.activate()
    from treetime import TreeAnc as RefTreeAnc
    from treetime_b200.dropin import accelerate
    T, aln = _inputs(37, 40, 400, mean_bl=0.004, amb=0.0, zero_frac=0.3)
    t1, a1 = _bio(T, aln); t2, a2 = _bio(T, aln)
    rt = RefTreeAnc(tree=t1, aln=a1, gtr=_ref_gtr(), rng_seed=1, verbose=0)
    dt = accelerate(RefTreeAnc)(tree=t2, aln=a2, gtr=_ref_gtr(), rng_seed=1, verbose=0)
    assert rt.infer_ancestral_sequences(marginal=True) == dt.infer_ancestral_sequences(marginal=True)
    n0 = len(list(dt.tree.find_clades()))
    rt.prune_short_branches(); dt.prune_short_branches()
    assert len(list(dt.tree.find_clades())) == len(list(rt.tree.find_clades())) < n0
    assert rt.infer_ancestral_sequences(marginal=True) == dt.infer_ancestral_sequences(marginal=True)
    assert dt._b200_live
    assert abs(rt.sequence_LH() - dt.sequence_LH()) <= LH_RTOL * abs(rt.sequence_LH())
    # swap two leaves in place: parent / child_idx stay the same
    for tt in (rt, dt):
        tips = {n.name: n for n in tt.tree.get_terminals()}
        names = sorted(tips)
        a, d = tips[names[0]], tips[names[-1]]
        pa, pd = a.up, d.up
        ia, id_ = pa.clades.index(a), pd.clades.index(d)
        pa.clades[ia], pd.clades[id_] = d, a
        a.branch_length, d.branch_length = d.branch_length, a.branch_length
        tt._prepare_nodes()
    lh_before = dt.sequence_LH()
    assert rt.infer_ancestral_sequences(marginal=True) == dt.infer_ancestral_sequences(marginal=True)
    assert abs(rt.sequence_LH() - dt.sequence_LH()) <= LH_RTOL * abs(rt.sequence_LH())
    assert abs(rt.sequence_LH() - lh_before) > 1e-6


def _dated_inputs(seed, n, L, rate=0.002, mean_bl=0.004, zero_frac=0.0):
    T, aln = _inputs(seed, n, L, mean_bl=mean_bl, amb=0.0, zero_frac=zero_frac)
    d2r, stack = {}, [(T.root, 0.0)]
    while stack:
        node, d = stack.pop()
        if not node.clades:
            d2r[node.name] = d
        for c in node.clades:
            stack.append((c, d + c.branch_length))
    return T, aln, {k: 2000.0 + v / rate for k, v in d2r.items()}


def test_gpu_dropin_treetime_run_marginal():
    """accelerate(TreeTime).run(branch_length_mode='marginal', max_iter=2) -- the production caller
    (treetime.py:235-243,345-352; clock_tree.py:355-356) -- equals the reference run."""
    refenv.activate()
    from treetime import TreeTime
    from treetime_b200.dropin import accelerate
    T, aln, dates = _dated_inputs(41, 50, 800)
    kw = dict(root=None, infer_gtr=False, max_iter=2, branch_length_mode='marginal', time_marginal=False, resolve_polytomies=False)
    t1, a1 = _bio(T, aln); t2, a2 = _bio(T, aln)
    ref = TreeTime(tree=t1, aln=a1, gtr=_ref_gtr(), dates=dates, verbose=0, rng_seed=1)
    ref.run(**kw)
    ours = accelerate(TreeTime)(tree=t2, aln=a2, gtr=_ref_gtr(), dates=dates, verbose=0, rng_seed=1)
    ours.run(**kw)
    assert type(ours._engine).__module__ == 'treetime_b200.engine' and ours._engine.launch_count() > 0 and ours._b200_live
    for a, b in zip(ref.tree.find_clades(), ours.tree.find_clades()):
        assert a.name == b.name
        assert np.isclose(a.numdate, b.numdate, rtol=0, atol=1e-4)
        assert np.isclose(a.branch_length, b.branch_length, rtol=1e-5, atol=1e-10)
    assert abs(ref.tree.sequence_marginal_LH - ours.tree.sequence_marginal_LH) <= LH_RTOL * abs(ref.tree.sequence_marginal_LH)
    assert np.isclose(ref.date2dist.clock_rate, ours.date2dist.clock_rate, rtol=1e-6)


def test_gpu_dropin_batched_branch_grids_match_reference_interpolators():
    """N1: BranchLenInterpolator tables (branch_len_interpolator.py:103-110) built from the device's
    batched prob_t_profiles grids equal the reference's, node by node."""
    refenv.activate()
    from treetime import ClockTree
    from treetime_b200.dropin import accelerate, B200ClockMixin
    T, aln, dates = _dated_inputs(51, 40, 600, zero_frac=0.2)
    kw = dict(dates=dates, verbose=0, rng_seed=1, branch_length_mode='marginal')
    t1, a1 = _bio(T, aln); t2, a2 = _bio(T, aln)
    ref = ClockTree(tree=t1, aln=a1, gtr=_ref_gtr(), **kw)
    ours = accelerate(ClockTree)(tree=t2, aln=a2, gtr=_ref_gtr(), **kw)
    assert isinstance(ours, B200ClockMixin)
    ref.init_date_constraints()
    launches0 = 0
    ours.init_date_constraints()
    assert ours._b200_live and 'prob_t_profiles' not in ours.gtr.__dict__
    assert type(ours._engine).__module__ == 'treetime_b200.engine' and ours._engine.launch_count() > launches0
    n_checked, worst = 0, 0.0
    for a, b in zip(ref.tree.find_clades(), ours.tree.find_clades()):
        if a.up is None:
            continue
        ia, ib = a.branch_length_interpolator, b.branch_length_interpolator
        assert np.array_equal(ia.x, ib.x)
        worst = max(worst, np.abs(ia.y - ib.y).max())
        # y = -log LH(t) on the grid, which starts at t ~ 1e-25: there the off-diagonal entries of V exp(lambda t) V^-1 are
        # differences of O(1) products -- uncertain at ~1e-16 absolute, i.e. ~1e-8 relative at t ~ 1e-8 -- in the reference
        # as much as on the device, and every mismatching pattern multiplies that into the sum
        assert np.allclose(ia.y, ib.y, rtol=1e-7, atol=1e-4)
        assert np.isclose(ia.peak_pos, ib.peak_pos)
        n_checked += 1
    assert n_checked == 78
    n, m = list(ours.tree.find_clades())[3], list(ref.tree.find_clades())[3]
    assert np.abs(n.profile_pair[0] - m.profile_pair[0]).max() < PROF_ATOL
    print('N1 on the device: %d interpolators, max|dy| = %.1e' % (n_checked, worst))


def test_gpu_dropin_joint_and_sampling():
    """N2 and sample_from_profile through the drop-in on the device: joint ML sequences and LH, root sampling and
    all-node sampling consume the caller's RNG like the reference."""
    rt, dt = _pair(seed=36, n=40, L=500)
    assert rt.infer_ancestral_sequences(marginal=False) == dt.infer_ancestral_sequences(marginal=False)
    assert dt._b200_live
    assert abs(rt.tree.sequence_joint_LH - dt.tree.sequence_joint_LH) <= LH_RTOL * abs(rt.tree.sequence_joint_LH)
    diff = sum(int((a.cseq != b.cseq).sum()) for a, b in zip(rt.tree.find_clades(), dt.tree.find_clades()) if not a.is_terminal())
    assert diff == 0
    n1 = rt.infer_ancestral_sequences(marginal=True, sample_from_profile=True)
    n2 = dt.infer_ancestral_sequences(marginal=True, sample_from_profile=True)
    assert n1 == n2 and n1 > 0
    for a, b in zip(rt.tree.find_clades(), dt.tree.find_clades()):
        if not a.is_terminal():
            assert (a.cseq == b.cseq).all()
    assert rt.rng.random() == dt.rng.random()
