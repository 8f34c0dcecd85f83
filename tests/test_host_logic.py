"""CPU: host-side pieces -- Brent, pattern compression, flattening, sharding, the C-ABI surface."""
import os
import re
import numpy as np
import pytest

from treetime_b200 import synth, _lib
from treetime_b200.brent import brent_lockstep, BracketError
from treetime_b200.dist import shard_bounds
from treetime_b200.flatten import FlatTopology
from treetime_b200.sequence_data import SequenceData
from treetime_b200.tree import read_newick
import util

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_brent_lockstep_equals_scipy_per_element():
    from scipy.optimize import minimize_scalar
    rng = np.random.default_rng(0)
    n = 60
    c = rng.uniform(-1.5, 1.5, n); k = rng.uniform(0.5, 3, n); q4 = rng.uniform(0, 2, n)

    def f(i, x):
        return float(k[i] * (x - c[i]) ** 2 + q4[i] * abs(x - c[i]) ** 3)

    for tol in (1e-2 + 1e-8, 1e-10):
        r = brent_lockstep(lambda idx, u: np.array([f(i, float(x)) for i, x in zip(idx, u)]),
                           np.full(n, -2.0), c + 0.05, np.full(n, 2.0), tol=tol)
        for i in range(n):
            o = minimize_scalar(lambda x: f(i, x), bracket=[-2.0, c[i] + 0.05, 2.0], tol=tol, method='brent')
            assert o.x == r['x'][i] and o.nfev == r['nfev'][i] and o.nit == r['nit'][i]
        assert r['success'].all()


def test_brent_invalid_bracket_raises_like_scipy():
    with pytest.raises(ValueError):
        brent_lockstep(lambda idx, u: (u - 5.0) ** 2, [-2.0], [0.0], [2.0])     # minimum outside the bracket
    with pytest.raises(BracketError):
        brent_lockstep(lambda idx, u: u ** 2, [-2.0], [np.nan], [2.0])


def test_pattern_compression_semantics():
    """sequence_data.py:325-464: constant columns merge (ambiguous replaced), variable columns are private."""
    aln = {'a': 'AAAC-ANA', 'b': 'AAACTAAA', 'c': 'ANAGTAAA'}
    sd = SequenceData(aln, ambiguous='N', fill_overhangs=False)
    #            cols: 0 A  1 A(N->A) 2 A  3 C/C/G var 4 -/T/T var 5 A 6 N->A 7 A
    assert sd.compressed_length == 3
    assert list(sd.multiplicity()) == [6.0, 1.0, 1.0]
    assert list(sd.full_to_compressed_sequence_map) == [0, 0, 0, 1, 2, 0, 0, 0]
    assert ''.join(sd.compressed_alignment['c']) == 'AGT'
    assert ''.join(sd.compressed_alignment['a']) == 'AC-'
    full = sd.compressed_to_full_sequence(np.array(list('AGT')), as_string=True)
    assert full == 'AAAGTAAA'
    sd2 = SequenceData(aln, ambiguous='N', compress=False, fill_overhangs=False)
    assert sd2.compressed_length == 8 and sd2.multiplicity().sum() == 8
    sd3 = SequenceData({'a': '--AC--', 'b': 'ACACGT'}, ambiguous='N', fill_overhangs=True)
    assert ''.join(sd3.aln['a']) == 'NNACNN'


def test_flatten_preorder_and_csr():
    t = read_newick('((A:0.1,B:0.2)X:0.3,(C:0.1,(D:0.1,E:0.2):0.05):0.2,F:0.4);')
    topo = FlatTopology(t.root)
    names = [n.name for n in topo.nodes]
    assert names[:4] == [None, 'X', 'A', 'B'] and topo.n_tips == 6 and topo.n_nodes == 10
    assert topo.parent[0] == -1 and all(topo.parent[1:] < np.arange(1, 10))
    assert list(np.diff(topo.child_ptr))[0] == 3
    for n in range(10):
        for c in topo.child_idx[topo.child_ptr[n]:topo.child_ptr[n + 1]]:
            assert topo.parent[c] == n
    assert sorted(topo.tip_row[topo.tip_row >= 0]) == list(range(6))
    t2 = read_newick(t.to_newick())
    assert t2.to_newick() == t.to_newick()


def test_synth_generator_is_pinned():
    """bench/test inputs are regenerated from seeds: pin the generator's output."""
    import hashlib
    tree = synth.random_tree(50, seed=3, mean_bl=0.01)
    g = util.nuc_gtr()
    idx = synth.evolve_alignment(tree, 200, g.Pi, g.W, seed=3)
    sha = hashlib.sha256(np.vstack([idx[k] for k in sorted(idx)]).tobytes()).hexdigest()
    assert sha == SYNTH_SHA, sha


SYNTH_SHA = '4f06b5574c6402d9409187c5a9637ffa32327e24c1dd1b50f0590f0fb52d325d'


def test_shard_bounds_cover_and_balance():
    for n in (1, 7, 100, 22171):
        for w in (1, 2, 3, 8):
            b = [shard_bounds(n, r, w) for r in range(w)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 1


def test_c_abi_exports_every_declared_symbol():
    """include/ttb.h <-> libttb.so <-> the ctypes table: no compute calls, just the surface."""
    hdr = open(os.path.join(ROOT, 'include', 'ttb.h')).read()
    declared = set(re.findall(r'\b(ttb_[a-z_0-9]+)\s*\(', hdr))
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    lib = _lib.load()
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.ttb_version() >= 100
    assert [lib.ttb_supports_n_states(q) for q in (1, 4, 5, 9, 20, 22)] == [0, 1, 1, 0, 1, 1]
    assert set(_lib.Q_VALUES) == {q for q in range(1, 40) if lib.ttb_supports_n_states(q)}


def test_product_path_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    from treetime_b200.engine import Engine
    with pytest.raises(_lib.TTBError):
        Engine(5)
    from treetime_b200.treeanc import TreeAnc
    tree = synth.random_tree(8, seed=1, mean_bl=0.01)
    g = util.nuc_gtr()
    aln = {k: g.alphabet[v] for k, v in synth.evolve_alignment(tree, 50, g.Pi, g.W, seed=1).items()}
    tt = TreeAnc(tree=tree, aln=aln, gtr=g)
    with pytest.raises(_lib.TTBError):
        tt.infer_ancestral_sequences(marginal=True)       # no CPU fallback


def test_threaded_column_scans_equal_plain_numpy():
    """The host side of pattern compression works on column / row blocks on a thread pool: same statistics, same
    gather, same overhang filling as the one-shot numpy expressions they replaced."""
    from treetime_b200 import sequence_data as sdm
    rng = np.random.default_rng(3)
    A = rng.choice(np.frombuffer(b'ACGT-NRY', dtype=np.uint8), size=(700, 5000), p=[.22, .22, .22, .22, .05, .05, .01, .01])
    A[:, 10] = ord('N'); A[:, 11] = ord('A'); A[5, 11] = ord('N')        # all-ambiguous and constant-after-replacement columns
    amb = ord('N')
    lo, hi, aa = sdm._column_stats(A, amb)
    is_amb = A == amb
    assert np.array_equal(lo, np.where(is_amb, 255, A).min(axis=0)) and np.array_equal(hi, np.where(is_amb, 0, A).max(axis=0))
    assert np.array_equal(aa, is_amb.all(axis=0))
    cols = rng.permutation(5000)[:1234]
    G_ = sdm._gather_columns(A, cols)
    assert G_.flags['C_CONTIGUOUS'] and np.array_equal(G_, A[:, cols])
    # overhangs: only rows that start / end with a gap are touched; an all-gap row becomes all ambiguous
    B = A[:40, :200].copy()
    B[3] = ord('-'); B[5, :20] = ord('-'); B[7, -30:] = ord('-'); B[9, 0] = ord('-'); B[11, 50:60] = ord('-')
    sd = SequenceData({'s%02d' % i: B[i].copy() for i in range(40)}, ambiguous='N', fill_overhangs=True)
    nongap = B != ord('-'); any_ng = nongap.any(axis=1)
    first = np.where(any_ng, nongap.argmax(axis=1), B.shape[1]); last = np.where(any_ng, B.shape[1] - 1 - nongap[:, ::-1].argmax(axis=1), -1)
    pos = np.arange(B.shape[1])[None, :]
    ref = B.copy(); ref[(pos < first[:, None]) | (pos > last[:, None])] = ord('N')
    assert np.array_equal(sd.matrix, ref)


def test_vectorised_branch_length_floor_and_override():
    """_branch_lengths_to_gtr applies treeanc.py:752-760 to all nodes at once and still honours a subclass that
    overrides the per-node method."""
    import oracle_engine
    from treetime_b200.treeanc import TreeAnc
    tree = synth.random_tree(30, seed=2, mean_bl=1e-4, zero_frac=0.3)
    g = util.nuc_gtr()
    aln = {k: g.alphabet[v] for k, v in synth.evolve_alignment(tree, 120, g.Pi, g.W, seed=2).items()}
    tt = TreeAnc(tree=tree.to_newick(), aln=aln, gtr=g, engine_factory=oracle_engine.factory)
    nodes = tt._flat().nodes
    per_node = np.array([tt._branch_length_to_gtr(n) for n in nodes])
    assert np.array_equal(tt._branch_lengths_to_gtr(nodes), per_node) and (per_node[1:] > 0).all()
    tt.use_mutation_length = True
    for n in nodes:
        n.mutation_length = 2.0 * (n.branch_length or 0.0)
    assert np.array_equal(tt._branch_lengths_to_gtr(nodes), np.array([tt._branch_length_to_gtr(n) for n in nodes]))

    class Doubling(TreeAnc):
        def _branch_length_to_gtr(self, node):
            return 2.0 * TreeAnc._branch_length_to_gtr(self, node)
    t2 = Doubling(tree=tree.to_newick(), aln=aln, gtr=util.nuc_gtr(), engine_factory=oracle_engine.factory)
    n2 = t2._flat().nodes
    assert np.array_equal(t2._branch_lengths_to_gtr(n2), np.array([t2._branch_length_to_gtr(n) for n in n2]))
    assert np.array_equal(t2._branch_lengths_to_gtr(n2)[1:], 2.0 * per_node[1:])


def test_fastscan_equals_python_scans(monkeypatch):
    """csrc/ttb_fastscan.c (optional host helper): the C loops over the nodes' dicts give what the numpy / map() forms
    give, report 'cannot' on anything they do not handle, and the mirror's per-pass scans agree with and without them."""
    import oracle_engine
    from treetime_b200 import _fastscan
    from treetime_b200.treeanc import TreeAnc
    _fastscan.build()
    assert _fastscan.load() is not None

    class Node(object):
        def __init__(self, i):
            self.branch_length = 0.25 * i
            self.mask = None
    nodes = [Node(i) for i in range(1000)]
    nodes[3].branch_length = 7                     # int
    nodes[4].branch_length = np.float64(1.5)       # float subclass
    dicts = [n.__dict__ for n in nodes]
    out = np.full(1000, -1.0)
    assert _fastscan.scan_float_attr(dicts, 'branch_length', out, 1)
    assert out[0] == -1.0 and np.array_equal(out[1:], np.array([float(n.branch_length) for n in nodes[1:]]))
    assert _fastscan.any_not_none(dicts, 'mask') is False
    nodes[999].mask = np.zeros(3)
    assert _fastscan.any_not_none(dicts, 'mask') is True
    nodes[10].branch_length = None
    assert not _fastscan.scan_float_attr(dicts, 'branch_length', out, 1)          # None: the caller's generic path decides
    del nodes[10].__dict__['branch_length']
    assert not _fastscan.scan_float_attr(dicts, 'branch_length', out, 1)          # missing attribute
    assert not _fastscan.scan_float_attr(dicts, 'branch_length', np.zeros(1000, dtype=np.float32), 1)
    assert _fastscan.any_not_none(tuple(dicts), 'mask') is None                     # not a list: cannot tell
    # both scans in one walk
    for n in nodes:
        n.branch_length = 0.5
    nodes[999].mask = None
    both = np.full(1000, -1.0)
    assert _fastscan.scan_nodes(dicts, 'branch_length', 'mask', both, 1) is False and both[0] == -1.0 and (both[1:] == 0.5).all()
    nodes[0].mask = np.ones(2)                                                     # the root's mask counts as well
    assert _fastscan.scan_nodes(dicts, 'branch_length', 'mask', both, 1) is True
    nodes[500].branch_length = 'x'
    assert _fastscan.scan_nodes(dicts, 'branch_length', 'mask', both, 1) is None

    tree = synth.random_tree(40, seed=5, mean_bl=1e-3, zero_frac=0.2)
    g = util.nuc_gtr()
    aln = {k: g.alphabet[v] for k, v in synth.evolve_alignment(tree, 90, g.Pi, g.W, seed=5).items()}
    tt = TreeAnc(tree=tree.to_newick(), aln=aln, gtr=g, engine_factory=oracle_engine.factory)
    flat_nodes = tt._flat().nodes
    fast = tt._branch_lengths_to_gtr(flat_nodes)
    monkeypatch.setattr(_fastscan, 'scan_float_attr', lambda *a, **k: False)
    monkeypatch.setattr(_fastscan, 'any_not_none', lambda *a, **k: None)
    monkeypatch.setattr(_fastscan, 'scan_nodes', lambda *a, **k: None)
    assert np.array_equal(tt._branch_lengths_to_gtr(flat_nodes), fast)
    tt.infer_ancestral_sequences(marginal=True)
    lh = tt.tree.total_sequence_LH
    monkeypatch.undo()
    tt.infer_ancestral_sequences(marginal=True)
    assert tt.tree.total_sequence_LH == lh
