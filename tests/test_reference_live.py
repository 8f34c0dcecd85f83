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


def _pair(seed=33, n=30, L=300, **kw):
    refenv.activate()
    import oracle_engine
    from treetime import GTR as RG, TreeAnc as RefTreeAnc
    from treetime_b200 import synth
    from treetime_b200.dropin import accelerate
    from treetime_b200.gtr import GTR
    pi = np.array([.3, .2, .2, .29, .01])
    T = synth.random_tree(n, seed=seed, mean_bl=0.01)
    g = GTR.custom(pi=pi.copy(), W=np.ones((5, 5)), alphabet='nuc')
    idx = synth.evolve_alignment(T, L, g.Pi, g.W, seed=seed)
    aln = synth.sprinkle_ambiguous({k: g.alphabet[v] for k, v in idx.items()}, 0.02, 'N-R', seed=3)
    mk = lambda: RG.custom(pi=pi.copy(), W=np.ones((5, 5)), alphabet='nuc')  # noqa: E731
    rt = refenv.reference_treeanc(T.to_newick(), aln, mk(), rng_seed=1, **kw)
    # the drop-in: same constructor as the reference's TreeAnc, engine injected for the CPU test
    from io import StringIO
    from Bio import Phylo
    from Bio.Align import MultipleSeqAlignment
    from Bio.SeqRecord import SeqRecord
    from Bio.Seq import Seq
    B200TreeAnc = accelerate(RefTreeAnc)
    dt = B200TreeAnc(tree=Phylo.read(StringIO(T.to_newick()), 'newick'),
                     aln=MultipleSeqAlignment([SeqRecord(Seq(''.join(aln[k])), id=k, name=k, description='') for k in aln]),
                     gtr=mk(), rng_seed=1, verbose=0, engine_factory=oracle_engine.factory, **kw)
    return rt, dt


def test_dropin_mixin_on_real_treeanc():
    """class B200TreeAnc(B200MarginalMixin, treetime.TreeAnc): same answers as treetime.TreeAnc
    through the reference's own accessors (lazy clade attributes)."""
    rt, dt = _pair()
    assert rt.infer_ancestral_sequences(marginal=True) == dt.infer_ancestral_sequences(marginal=True)
    assert dt._engine is not None and dt._b200_live              # the device path ran, not the fallback
    assert rt.sequence_LH() == dt.sequence_LH() and np.array_equal(rt.tree.sequence_LH, dt.tree.sequence_LH)
    assert rt.tree.sequence_marginal_LH == dt.tree.sequence_marginal_LH
    for a, b in zip(rt.tree.find_clades(), dt.tree.find_clades()):
        assert a.name == b.name
        if not a.is_terminal():
            assert np.array_equal(a.marginal_profile, b.marginal_profile)
            assert (a.cseq == b.cseq).all() and a.mutations == b.mutations
        if a.up is not None:
            pa, pb = rt.marginal_branch_profile(a), dt.marginal_branch_profile(b)
            assert np.array_equal(pa[0], pb[0]) and np.array_equal(pa[1], pb[1])
    assert rt.infer_ancestral_sequences(marginal=True) == dt.infer_ancestral_sequences(marginal=True) == 0
    n = [c for c in dt.tree.find_clades()][7]
    m = [c for c in rt.tree.find_clades()][7]
    assert np.array_equal(rt.get_branch_mutation_matrix(m), dt.get_branch_mutation_matrix(n))
    assert rt.optimal_marginal_branch_length(m) == dt.optimal_marginal_branch_length(n)
    ra = rt.get_reconstructed_alignment(); da = dt.get_reconstructed_alignment()
    assert [str(r.seq) for r in ra] == [str(r.seq) for r in da]
    rt.optimize_tree(branch_length_mode='marginal', max_iter=2, infer_gtr=True, prune_short=True)
    dt.optimize_tree(branch_length_mode='marginal', max_iter=2, infer_gtr=True, prune_short=True)
    a = np.array([c.branch_length for c in rt.tree.find_clades()]); b = np.array([c.branch_length for c in dt.tree.find_clades()])
    assert a.shape == b.shape and np.allclose(a[1:], b[1:], rtol=1e-9, atol=1e-14)
    assert np.isclose(rt.sequence_LH(), dt.sequence_LH(), rtol=1e-12)
    assert np.allclose(rt.gtr.W, dt.gtr.W, rtol=1e-9) and type(dt.gtr) is type(rt.gtr)
    rt.optimize_gtr_rate(); dt.optimize_gtr_rate()
    assert np.isclose(rt.gtr.mu, dt.gtr.mu, rtol=1e-9)


def test_dropin_per_branch_masks_on_the_device_path():
    """ARG mode (arg.py:128-133): per-branch 0/1 masks are handled by the engine -- masked up-messages dropped
    (treeanc.py:862-872), masked children keep their subtree profile (:909-917), masked multiplicities in the
    branch-length objective (:1294,1326-1333) and in the substitution statistics (:1564-1572).  Fractional masks
    and joint reconstruction with masks fall back to the reference."""
    rt, dt = _pair(seed=34)
    # joint (device path, N2) then marginal: N_diff is counted against the joint sequences
    assert rt.infer_ancestral_sequences(marginal=False) == dt.infer_ancestral_sequences(marginal=False)
    assert rt.infer_ancestral_sequences(marginal=True) == dt.infer_ancestral_sequences(marginal=True)
    L = rt.data.compressed_length
    seg = np.zeros(L); seg[:L // 2] = 1
    total = np.ones(L)
    rn, dn = list(rt.tree.find_clades()), list(dt.tree.find_clades())
    for k, (a, b) in enumerate(zip(rn, dn)):          # every node carries a mask, like setup_arg leaves them
        a.mask = b.mask = (seg if k % 3 == 0 else total)
    for tips in (False, True):
        assert (rt.infer_ancestral_sequences(marginal=True, reconstruct_tip_states=tips)
                == dt.infer_ancestral_sequences(marginal=True, reconstruct_tip_states=tips))
        assert dt._b200_live
        assert np.isclose(rt.sequence_LH(), dt.sequence_LH(), rtol=1e-13)
        assert np.allclose(rt.tree.sequence_LH, dt.tree.sequence_LH, rtol=1e-12, atol=1e-12)
        for a, b in zip(rn, dn):
            if a.up is not None:
                assert np.allclose(a.marginal_outgroup_LH, b.marginal_outgroup_LH, rtol=1e-12, atol=1e-15)
            if tips or not a.is_terminal():
                assert np.allclose(a.marginal_profile, b.marginal_profile, rtol=1e-12, atol=1e-15)
                assert (a.cseq == b.cseq).all()
    for a, b in zip(rn[1:], dn[1:]):
        x, y = rt.optimal_marginal_branch_length(a), dt.optimal_marginal_branch_length(b)
        assert np.isclose(x, y, rtol=1e-6, atol=1e-12), (a.name, x, y)
    g1 = rt.infer_gtr(marginal=True, pc=1.0); g2 = dt.infer_gtr(marginal=True, pc=1.0)
    assert np.allclose(g1.W, g2.W, rtol=1e-9) and np.allclose(g1.Pi, g2.Pi, rtol=1e-9)
    rt.optimize_tree(branch_length_mode='marginal', max_iter=2, infer_gtr=False, prune_short=False)
    dt.optimize_tree(branch_length_mode='marginal', max_iter=2, infer_gtr=False, prune_short=False)
    assert dt._b200_live
    a = np.array([c.branch_length for c in rt.tree.find_clades()]); b = np.array([c.branch_length for c in dt.tree.find_clades()])
    assert np.allclose(a[1:], b[1:], rtol=1e-7, atol=1e-12)
    assert np.isclose(rt.sequence_LH(), dt.sequence_LH(), rtol=1e-10)
    # removing the masks again
    for a, b in zip(rn, dn):
        a.mask = b.mask = None
    assert rt.infer_ancestral_sequences(marginal=True) == dt.infer_ancestral_sequences(marginal=True)
    assert dt._b200_live and np.isclose(rt.sequence_LH(), dt.sequence_LH(), rtol=1e-13)
    # a fractional mask has no device form => reference implementation, identical numbers
    frac = np.ones(L); frac[::3] = 0.5
    rn[4].mask = dn[4].mask = frac
    assert rt.infer_ancestral_sequences(marginal=True) == dt.infer_ancestral_sequences(marginal=True)
    assert not dt._b200_live
    assert rt.sequence_LH() == dt.sequence_LH()
    for a, b in zip(rt.tree.find_clades(), dt.tree.find_clades()):
        if not a.is_terminal():
            assert np.array_equal(a.marginal_profile, b.marginal_profile)


def test_dropin_treetime_run_marginal():
    """TreeTime.run(branch_length_mode='marginal') -- the production caller (test_treetime.py:310-363)
    -- on top of the drop-in gives the reference's result."""
    refenv.activate()
    import oracle_engine
    from io import StringIO
    from Bio import Phylo
    from Bio.Align import MultipleSeqAlignment
    from Bio.SeqRecord import SeqRecord
    from Bio.Seq import Seq
    from treetime import GTR as RG, TreeTime
    from treetime_b200 import synth
    from treetime_b200.dropin import accelerate
    from treetime_b200.gtr import GTR
    pi = np.array([.3, .2, .2, .29, .01])
    T = synth.random_tree(25, seed=41, mean_bl=0.004)
    g = GTR.custom(pi=pi.copy(), W=np.ones((5, 5)), alphabet='nuc')
    idx = synth.evolve_alignment(T, 500, g.Pi, g.W, seed=41)
    aln = {k: g.alphabet[v] for k, v in idx.items()}
    # sampling dates proportional to root-to-tip distance (a clock-like tree)
    d2r = {}
    stack = [(T.root, 0.0)]
    while stack:
        n, d = stack.pop()
        if not n.clades:
            d2r[n.name] = d
        for c in n.clades:
            stack.append((c, d + c.branch_length))
    dates = {k: 2000.0 + v / 0.002 for k, v in d2r.items()}
    mk = lambda: RG.custom(pi=pi.copy(), W=np.ones((5, 5)), alphabet='nuc')  # noqa: E731
    mkaln = lambda: MultipleSeqAlignment([SeqRecord(Seq(''.join(aln[k])), id=k, name=k, description='') for k in aln])  # noqa: E731
    mktree = lambda: Phylo.read(StringIO(T.to_newick()), 'newick')  # noqa: E731
    kw = dict(root=None, infer_gtr=False, max_iter=1, branch_length_mode='marginal', time_marginal=False, resolve_polytomies=False)
    ref = TreeTime(tree=mktree(), aln=mkaln(), gtr=mk(), dates=dates, verbose=0, rng_seed=1)
    ref.run(**kw)
    ours = accelerate(TreeTime)(tree=mktree(), aln=mkaln(), gtr=mk(), dates=dates, verbose=0, rng_seed=1,
                                engine_factory=oracle_engine.factory)
    ours.run(**kw)
    assert ours._engine is not None and ours._engine.launch_count() > 0
    for a, b in zip(ref.tree.find_clades(), ours.tree.find_clades()):
        assert a.name == b.name
        assert np.isclose(a.numdate, b.numdate, rtol=0, atol=1e-6)
        assert np.isclose(a.branch_length, b.branch_length, rtol=1e-7, atol=1e-12)
    assert np.isclose(ref.tree.sequence_marginal_LH, ours.tree.sequence_marginal_LH, rtol=1e-10)


def test_dropin_batched_branch_grids_match_reference_interpolators():
    """N1: BranchLenInterpolator tables built from device-evaluated grids equal the reference's."""
    refenv.activate()
    import oracle_engine
    from io import StringIO
    from Bio import Phylo
    from Bio.Align import MultipleSeqAlignment
    from Bio.SeqRecord import SeqRecord
    from Bio.Seq import Seq
    from treetime import GTR as RG, ClockTree
    from treetime.branch_len_interpolator import BranchLenInterpolator
    from treetime_b200 import synth
    from treetime_b200.dropin import accelerate, branch_length_grid, B200ClockMixin
    from treetime_b200.gtr import GTR
    pi = np.array([.3, .2, .2, .29, .01])
    T = synth.random_tree(20, seed=51, mean_bl=0.004, zero_frac=0.2)
    g = GTR.custom(pi=pi.copy(), W=np.ones((5, 5)), alphabet='nuc')
    aln = {k: g.alphabet[v] for k, v in synth.evolve_alignment(T, 400, g.Pi, g.W, seed=51).items()}
    dates = {k: 2000.0 + i * 0.1 for i, k in enumerate(sorted(aln))}
    mk = lambda: RG.custom(pi=pi.copy(), W=np.ones((5, 5)), alphabet='nuc')  # noqa: E731
    mkaln = lambda: MultipleSeqAlignment([SeqRecord(Seq(''.join(aln[k])), id=k, name=k, description='') for k in aln])  # noqa: E731
    mktree = lambda: Phylo.read(StringIO(T.to_newick()), 'newick')  # noqa: E731
    kw = dict(dates=dates, verbose=0, rng_seed=1, branch_length_mode='marginal')
    ref = ClockTree(tree=mktree(), aln=mkaln(), gtr=mk(), **kw)
    ours = accelerate(ClockTree)(tree=mktree(), aln=mkaln(), gtr=mk(), engine_factory=oracle_engine.factory, **kw)
    assert isinstance(ours, B200ClockMixin)
    ref.init_date_constraints()
    ours.init_date_constraints()
    assert ours._b200_live and 'prob_t_profiles' not in ours.gtr.__dict__
    n_checked = 0
    for a, b in zip(ref.tree.find_clades(), ours.tree.find_clades()):
        if a.up is None:
            continue
        ia, ib = a.branch_length_interpolator, b.branch_length_interpolator
        # same grid (our restated grid construction) and same tabulated values
        assert np.array_equal(ia.x, ib.x)
        assert np.allclose(ia.y, ib.y, rtol=1e-9, atol=1e-8), np.abs(ia.y - ib.y).max()
        assert np.isclose(ia.peak_pos, ib.peak_pos)
        gexp = branch_length_grid(a.mutation_length, ref.one_mutation, ref.branch_grid_points)
        assert set(np.unique(gexp)) >= set(ia.x) or len(ia.x) <= len(gexp)
        n_checked += 1
    assert n_checked == 38
    # the proxy still behaves like the (pp, pc) tuple when somebody indexes it
    n = list(ours.tree.find_clades())[3]
    m = list(ref.tree.find_clades())[3]
    assert np.array_equal(n.profile_pair[0], m.profile_pair[0]) and np.array_equal(n.profile_pair[1], m.profile_pair[1])


def test_dropin_joint_reconstruction_and_treetime_run_joint():
    """N2: joint ML reconstruction through the drop-in equals the reference (incl. root sampling with the
    shared RNG and a full TreeTime.run in joint mode)."""
    rt, dt = _pair(seed=36)
    assert rt.infer_ancestral_sequences(marginal=False) == dt.infer_ancestral_sequences(marginal=False)
    assert dt._b200_live and dt._engine.launch_count() > 0
    assert rt.tree.sequence_joint_LH == dt.tree.sequence_joint_LH and np.array_equal(rt.tree.sequence_LH, dt.tree.sequence_LH)
    for a, b in zip(rt.tree.find_clades(), dt.tree.find_clades()):
        if not a.is_terminal():
            assert (a.cseq == b.cseq).all() and a.mutations == b.mutations
    assert rt.infer_ancestral_sequences(marginal=False) == dt.infer_ancestral_sequences(marginal=False) == 0
    # root sampling consumes the RNG identically
    n1 = rt.infer_ancestral_sequences(marginal=False, sample_from_profile='root', reconstruct_tip_states=True)
    n2 = dt.infer_ancestral_sequences(marginal=False, sample_from_profile='root', reconstruct_tip_states=True)
    assert n1 == n2
    for a, b in zip(rt.tree.find_clades(), dt.tree.find_clades()):
        assert (a.cseq == b.cseq).all()
    # joint after marginal and back
    assert rt.infer_ancestral_sequences(marginal=True) == dt.infer_ancestral_sequences(marginal=True)
    assert rt.infer_ancestral_sequences(marginal=False) == dt.infer_ancestral_sequences(marginal=False)
    # joint branch-length optimisation: the reference's per-branch code on device pair counts (node.branch_state)
    for a, b in zip(rt.tree.find_clades(), dt.tree.find_clades()):
        if a.up is not None:
            rt.add_branch_state(a)
            assert np.array_equal(a.branch_state['pair'], b.branch_state['pair'])
            assert np.array_equal(a.branch_state['multiplicity'], b.branch_state['multiplicity'])
            assert rt.optimal_branch_length(a) == dt.optimal_branch_length(b)
    launches = dt._engine.launch_count()
    rt.optimize_tree(branch_length_mode='joint', max_iter=2, prune_short=False)
    dt.optimize_tree(branch_length_mode='joint', max_iter=2, prune_short=False)
    assert dt._engine.launch_count() > launches
    a = np.array([c.branch_length for c in rt.tree.find_clades()]); b = np.array([c.branch_length for c in dt.tree.find_clades()])
    assert np.array_equal(a[1:], b[1:])
    # ... and with every branch in one lock-step Brent
    dt.batched_joint_branch_lengths = True
    rt.optimize_tree(branch_length_mode='joint', max_iter=2, prune_short=True)
    dt.optimize_tree(branch_length_mode='joint', max_iter=2, prune_short=True)
    a = np.array([c.branch_length for c in rt.tree.find_clades()]); b = np.array([c.branch_length for c in dt.tree.find_clades()])
    # a minimum located from function values is only defined to ~sqrt(eps) relative: the batched objective
    # sums the same terms in a different order
    assert a.shape == b.shape and np.allclose(a[1:], b[1:], rtol=5e-6, atol=1e-10)


def test_state_pair_restatement_matches_reference():
    """oracle state_pair / fold_state_pairs == GTR.state_pair for the small- and large-alphabet code paths,
    with and without ignore_gaps, ambiguous characters included."""
    refenv.activate()
    import flat_numpy as O
    from treetime import GTR as RG
    from treetime_b200.pairs import fold_state_pairs, NONE
    rng = np.random.default_rng(5)
    for name in ('nuc', 'aa'):
        g = RG.standard('JC69', alphabet=name)
        ab = [str(c) for c in g.alphabet]
        chars = sorted(g.profile_map.keys())
        L = 400
        sp = rng.choice(ab, size=L)
        sc = np.where(rng.random(L) < 0.2, rng.choice(chars, size=L), sp)
        mult = rng.integers(1, 5, size=L).astype(float)
        for ig in (False, True):
            ref = g.state_pair(sp, sc, pattern_multiplicity=mult, ignore_gaps=ig)
            mine = O.state_pair(ab, g.gap_index, sp, sc, mult, ignore_gaps=ig)
            assert np.array_equal(ref[0], mine[0]) and np.array_equal(ref[1], mine[1])
            # the device table form: parent state x child code
            q, W = len(ab), len(chars)
            C = np.zeros((q, W)); F = np.full((q, W), NONE, dtype=np.int32)
            pi = np.array([ab.index(c) for c in sp]); ci = np.array([chars.index(c) for c in sc])
            np.add.at(C, (pi, ci), mult); np.minimum.at(F, (pi, ci), np.arange(L))
            fp, fm = fold_state_pairs(C, F, ab, chars, g.gap_index, ig)
            assert np.array_equal(ref[0], fp) and np.array_equal(ref[1], fm)
            if len(ref[1]):
                t = 0.1
                assert O.prob_t_compressed(g, mine[0], mine[1], t) == g.prob_t_compressed(ref[0], ref[1], t, return_log=True)
                assert O.optimal_t_compressed(g, mine[0], mine[1]) == g.optimal_t_compressed(ref[0], ref[1])


def test_seqgen_reproduces_reference_sequences():
    """N4: treetime_b200.SeqGen(reference_rng=True) == treetime.seqgen.SeqGen with the same seed -- single-model
    and site-specific GTR (oracle-backed engine here; the GPU test checks the kernel against the same oracle)."""
    refenv.activate()
    import oracle_engine
    from io import StringIO
    from Bio import Phylo
    from treetime import GTR as RG
    from treetime.gtr_site_specific import GTR_site_specific
    from treetime.seqgen import SeqGen as RefSeqGen
    from treetime_b200 import synth
    from treetime_b200.gtr import GTR, GTRSiteSpecific
    from treetime_b200.seqgen import SeqGen
    T = synth.random_tree(25, seed=61, mean_bl=0.15, polytomy_frac=0.2)
    L = 300
    pi = np.array([.3, .2, .2, .29, .01])
    cases = [(RG.custom(pi=pi.copy(), W=np.ones((5, 5)), alphabet='nuc'), GTR.custom(pi=pi.copy(), W=np.ones((5, 5)), alphabet='nuc'))]
    rs = GTR_site_specific.random(L=L, alphabet='nuc', rng=np.random.default_rng(62))
    ms = GTRSiteSpecific(alphabet='nuc', seq_len=L)
    ms.assign_rates(mu=np.array(rs.mu), pi=np.array(rs.Pi), W=np.array(rs.W))
    cases.append((rs, ms))
    for rg, mg in cases:
        ref = RefSeqGen(L, tree=Phylo.read(StringIO(T.to_newick()), 'newick'), gtr=rg, rng_seed=7, verbose=0)
        ref.evolve()
        mine = SeqGen(L, tree=T.to_newick(), gtr=mg, rng_seed=7, engine_factory=oracle_engine.factory)
        aln = mine.evolve(reference_rng=True)
        want = {r.id: np.array(list(str(r.seq))) for r in ref.get_aln(internal=True)}
        got = mine.get_aln(internal=True)
        named = [k for k in want if k in got]
        assert len(named) >= 25
        for k in named:
            assert (want[k] == got[k]).all(), k
        assert set(aln) == set(n.name for n in T.get_terminals())


def test_sample_from_profile_all_nodes_matches_reference():
    """sample_from_profile=True (treeanc.py:786-798,919-923): every node's sequence is drawn from its marginal
    profile with the caller's generator; the drop-in and the mirror consume the RNG like the reference, so the
    sampled sequences, N_diff and the following draws are identical -- also with reconstructed tips and across
    repeated calls (N_diff against the previous sampled states)."""
    rt, dt = _pair(seed=41)
    for kw in (dict(), dict(reconstruct_tip_states=True), dict(), dict(reconstruct_tip_states=True)):
        n1 = rt.infer_ancestral_sequences(marginal=True, sample_from_profile=True, **kw)
        n2 = dt.infer_ancestral_sequences(marginal=True, sample_from_profile=True, **kw)
        assert dt._b200_live
        assert n1 == n2 and n1 > 0
        for a, b in zip(rt.tree.find_clades(), dt.tree.find_clades()):
            if kw or not a.is_terminal():
                assert (a.cseq == b.cseq).all()
                assert np.array_equal(a.marginal_profile, b.marginal_profile)
    assert rt.rng.random() == dt.rng.random()
    # a sampled pass followed by an argmax pass: N_diff counts against the sampled states
    assert rt.infer_ancestral_sequences(marginal=True) == dt.infer_ancestral_sequences(marginal=True)
    # the mirror (own containers) does the same
    refenv.activate()
    import oracle_engine
    from treetime import GTR as RG
    from treetime_b200 import synth
    from treetime_b200.gtr import GTR
    from treetime_b200.treeanc import TreeAnc
    pi = np.array([.3, .2, .2, .29, .01])
    T = synth.random_tree(25, seed=43, mean_bl=0.02)
    g = GTR.custom(pi=pi.copy(), W=np.ones((5, 5)), alphabet='nuc')
    idx = synth.evolve_alignment(T, 200, g.Pi, g.W, seed=43)
    aln = {k: g.alphabet[v] for k, v in idx.items()}
    r2 = refenv.reference_treeanc(T.to_newick(), aln, RG.custom(pi=pi.copy(), W=np.ones((5, 5)), alphabet='nuc'), rng_seed=7)
    m2 = TreeAnc(tree=T.to_newick(), aln=aln, gtr=g, rng_seed=7, engine_factory=oracle_engine.factory)
    for _ in range(2):
        assert (r2.infer_ancestral_sequences(marginal=True, sample_from_profile=True)
                == m2.infer_ancestral_sequences(marginal=True, sample_from_profile=True))
        rn = {n.name: n for n in r2.tree.find_clades()}
        for n in m2.tree.find_clades():
            if not n.is_terminal():
                assert (n.cseq == rn[n.name].cseq).all()


def test_site_specific_gtr_inference_matches_reference():
    """infer_gtr(marginal=True, site_specific=True) (treeanc.py:1551-1627 + GTR_site_specific.infer): per-pattern
    statistics from the engine, first under a single model, then under the inferred site-specific model (per-pattern
    transition matrices), through the drop-in and through the mirror."""
    rt, dt = _pair(seed=52, n=20, L=120, compress=False)
    assert rt.infer_ancestral_sequences(marginal=True) == dt.infer_ancestral_sequences(marginal=True)
    for it in range(2):
        g1 = rt.infer_gtr(marginal=True, site_specific=True, pc=1.0)
        launches = dt._engine.launch_count()
        g2 = dt.infer_gtr(marginal=True, site_specific=True, pc=1.0)
        assert dt._b200_live
        assert np.allclose(g1.Pi, g2.Pi, rtol=1e-10, atol=1e-14) and np.allclose(g1.mu, g2.mu, rtol=1e-10)
        assert np.allclose(g1.W, g2.W, rtol=1e-10)
        assert rt.infer_ancestral_sequences(marginal=True) == dt.infer_ancestral_sequences(marginal=True)
        assert np.isclose(rt.sequence_LH(), dt.sequence_LH(), rtol=1e-11)
    # a single model inferred from a site-specific one: totals of the per-pattern statistics
    g1 = rt.infer_gtr(marginal=True, site_specific=False, pc=1.0)
    g2 = dt.infer_gtr(marginal=True, site_specific=False, pc=1.0)
    assert np.allclose(g1.Pi, g2.Pi, rtol=1e-10) and np.allclose(g1.W, g2.W, rtol=1e-10)
    # compressed data + site-specific inference is refused like in the reference
    rt2, dt2 = _pair(seed=53, n=10, L=60)
    dt2.infer_ancestral_sequences(marginal=True)
    with pytest.raises(TypeError):
        dt2.infer_gtr(marginal=True, site_specific=True)
    # the mirror: own GTRSiteSpecific from the same statistics
    refenv.activate()
    import oracle_engine
    from treetime import GTR as RG
    from treetime_b200 import synth
    from treetime_b200.gtr import GTR
    from treetime_b200.treeanc import TreeAnc
    pi = np.array([.3, .2, .2, .29, .01])
    T = synth.random_tree(15, seed=54, mean_bl=0.03)
    g = GTR.custom(pi=pi.copy(), W=np.ones((5, 5)), alphabet='nuc')
    idx = synth.evolve_alignment(T, 90, g.Pi, g.W, seed=54)
    aln = {k: g.alphabet[v] for k, v in idx.items()}
    r3 = refenv.reference_treeanc(T.to_newick(), aln, RG.custom(pi=pi.copy(), W=np.ones((5, 5)), alphabet='nuc'), rng_seed=7, compress=False)
    m3 = TreeAnc(tree=T.to_newick(), aln=aln, gtr=g, rng_seed=7, compress=False, engine_factory=oracle_engine.factory)
    r3.infer_ancestral_sequences(marginal=True); m3.infer_ancestral_sequences(marginal=True)
    a = r3.infer_gtr(marginal=True, site_specific=True, pc=2.0)
    b = m3.infer_gtr(marginal=True, site_specific=True, pc=2.0)
    assert np.allclose(a.Pi, b.Pi, rtol=1e-10, atol=1e-14) and np.allclose(a.mu, b.mu, rtol=1e-10) and np.allclose(a.W, b.W, rtol=1e-10)
    assert r3.infer_ancestral_sequences(marginal=True) == m3.infer_ancestral_sequences(marginal=True)
    assert np.isclose(r3.sequence_LH(), m3.sequence_LH(), rtol=1e-11)


def test_gtr_inference_from_reconstructed_sequences_matches_reference():
    """infer_gtr(marginal=False) (treeanc.py:1573-1589): the mutation / state-time counts the reference collects from
    node.mutations and node.cseq come from the device's pair counts -- after a joint and after a marginal
    reconstruction, with and without reconstructed tips, through the drop-in and the mirror."""
    for recon in (dict(marginal=False), dict(marginal=True), dict(marginal=False, reconstruct_tip_states=True)):
        rt, dt = _pair(seed=61)
        assert rt.infer_ancestral_sequences(**recon) == dt.infer_ancestral_sequences(**recon)
        calls, orig = [], dt._engine.branch_state_pairs
        dt._engine.branch_state_pairs = lambda *a, **k: (calls.append(1), orig(*a, **k))[1]
        g1 = rt.infer_gtr(marginal=False, pc=2.0)
        g2 = dt.infer_gtr(marginal=False, pc=2.0)
        assert calls                                          # counted from the engine's pair tables
        assert np.allclose(g1.W, g2.W, rtol=1e-12) and np.allclose(g1.Pi, g2.Pi, rtol=1e-12) and np.isclose(g1.mu, g2.mu, rtol=1e-12)
    # no reconstruction yet: infer_gtr runs the joint reconstruction itself (treeanc.py:1546-1547)
    rt, dt = _pair(seed=62)
    g1 = rt.infer_gtr(marginal=False, normalized_rate=False); g2 = dt.infer_gtr(marginal=False, normalized_rate=False)
    assert np.allclose(g1.W, g2.W, rtol=1e-12) and np.allclose(g1.Pi, g2.Pi, rtol=1e-12) and np.isclose(g1.mu, g2.mu, rtol=1e-12)
    assert dt.sequence_reconstruction == rt.sequence_reconstruction == 'joint'
    # the mirror: infer_ancestral_sequences(infer_gtr=True) in joint mode
    refenv.activate()
    import oracle_engine
    from treetime import GTR as RG
    from treetime_b200 import synth
    from treetime_b200.gtr import GTR
    from treetime_b200.treeanc import TreeAnc
    pi = np.array([.3, .2, .2, .29, .01])
    T = synth.random_tree(30, seed=63, mean_bl=0.02)
    g = GTR.custom(pi=pi.copy(), W=np.ones((5, 5)), alphabet='nuc')
    idx = synth.evolve_alignment(T, 250, g.Pi, g.W, seed=63)
    aln = synth.sprinkle_ambiguous({k: g.alphabet[v] for k, v in idx.items()}, 0.03, 'N-RY', seed=5)
    r = refenv.reference_treeanc(T.to_newick(), aln, RG.custom(pi=pi.copy(), W=np.ones((5, 5)), alphabet='nuc'), rng_seed=1)
    m = TreeAnc(tree=T.to_newick(), aln=aln, gtr=g, rng_seed=1, engine_factory=oracle_engine.factory)
    assert r.infer_ancestral_sequences(marginal=False, infer_gtr=True) == m.infer_ancestral_sequences(marginal=False, infer_gtr=True)
    assert np.allclose(r.gtr.W, m.gtr.W, rtol=1e-10) and np.allclose(r.gtr.Pi, m.gtr.Pi, rtol=1e-10)
    assert np.isclose(r.tree.sequence_joint_LH, m.tree.sequence_joint_LH, rtol=1e-12)
