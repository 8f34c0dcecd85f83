"""CPU, world_size 2 over gloo: pattern-sharded TreeAnc (oracle-backed engines) equals the
unsharded run -- total LH, N_diff, gathered per-node arrays, optimised branch lengths."""
import os
import sys
import numpy as np
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    for p in (ROOT, os.path.join(ROOT, 'oracle'), os.path.join(ROOT, 'tests')):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch.distributed as dist
    dist.init_process_group('gloo', init_method='tcp://127.0.0.1:%d' % port, rank=rank, world_size=world)
    import oracle_engine
    import util
    from treetime_b200 import synth
    from treetime_b200.dist import TorchComm
    from treetime_b200.treeanc import TreeAnc
    tree = synth.random_tree(24, seed=5, mean_bl=0.02)
    g = util.nuc_gtr()
    aln = {k: g.alphabet[v] for k, v in synth.evolve_alignment(tree, 301, g.Pi, g.W, seed=5).items()}
    tt = TreeAnc(tree=tree.to_newick(), aln=aln, gtr=g, comm=TorchComm(), engine_factory=oracle_engine.factory)
    lo, hi = tt._shard()
    n1 = tt.infer_ancestral_sequences(marginal=True)
    out = dict(rank=rank, shard=(lo, hi), n1=n1, lh=tt.sequence_LH(), site=tt.tree.sequence_LH.copy())
    nodes = list(tt.tree.find_clades())
    out['prof'] = nodes[3].marginal_profile.copy() if not nodes[3].is_terminal() else nodes[0].marginal_profile.copy()
    out['outg'] = nodes[5].marginal_outgroup_LH.copy()
    out['cseq'] = ''.join(nodes[0].cseq)
    tt.optimize_tree(branch_length_mode='marginal', max_iter=2, prune_short=False)
    out['bl'] = np.array([n.branch_length for n in tt.tree.find_clades()])
    out['lh2'] = tt.sequence_LH()
    tt.infer_gtr(marginal=True)
    out['W'] = np.array(tt.gtr.W)
    out.update(_extras(TreeAnc, tree, aln, util, dict(comm=TorchComm(), engine_factory=oracle_engine.factory)))
    q.put(out)
    dist.barrier()
    dist.destroy_process_group()


def _extras(TreeAnc, tree, aln, util, kw):
    """The paths with pattern-axis state besides the plain pass: sampled sequences (every rank draws the full
    uniforms and uses its slice), per-branch masks (sliced with the patterns), per-pattern statistics + site-specific
    GTR inference (all-gathered)."""
    out = {}
    t2 = TreeAnc(tree=tree.to_newick(), aln=aln, gtr=util.nuc_gtr(), rng_seed=11, compress=False, **kw)
    out['s_n'] = [t2.infer_ancestral_sequences(marginal=True, sample_from_profile=True, reconstruct_tip_states=tips)
                  for tips in (False, True, True)]
    out['s_seq'] = [''.join(n.cseq) for n in t2.tree.find_clades()]
    out['s_rng'] = t2.rng.random()
    L = t2.data.compressed_length
    seg = np.zeros(L); seg[: L // 3] = 1
    for k, n in enumerate(t2.tree.find_clades()):
        n.mask = seg if k % 2 else np.ones(L)
    out['m_n'] = t2.infer_ancestral_sequences(marginal=True)
    out['m_lh'] = t2.sequence_LH()
    nodes = list(t2.tree.find_clades())
    out['m_bl'] = np.array([t2.optimal_marginal_branch_length(n) for n in nodes[1:6]])
    for n in nodes:
        n.mask = None
    t2.infer_ancestral_sequences(marginal=True)
    g = t2.infer_gtr(marginal=True, site_specific=True, pc=1.0)
    out['ss_pi'] = np.array(g.Pi); out['ss_mu'] = np.array(g.mu)
    t2.infer_ancestral_sequences(marginal=True)
    out['ss_lh'] = t2.sequence_LH()
    return out


def test_two_rank_sharding_equals_single_rank():
    import socket
    s = socket.socket(); s.bind(('127.0.0.1', 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    outs = [q.get(timeout=300) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    outs.sort(key=lambda o: o['rank'])
    # single-rank reference run
    import oracle_engine
    import util
    from treetime_b200 import synth
    from treetime_b200.treeanc import TreeAnc
    tree = synth.random_tree(24, seed=5, mean_bl=0.02)
    g = util.nuc_gtr()
    aln = {k: g.alphabet[v] for k, v in synth.evolve_alignment(tree, 301, g.Pi, g.W, seed=5).items()}
    tt = TreeAnc(tree=tree.to_newick(), aln=aln, gtr=g, engine_factory=oracle_engine.factory)
    n1 = tt.infer_ancestral_sequences(marginal=True)
    nodes = list(tt.tree.find_clades())
    L = tt.data.compressed_length
    assert outs[0]['shard'][0] == 0 and outs[0]['shard'][1] == outs[1]['shard'][0] and outs[1]['shard'][1] == L
    for o in outs:
        assert o['n1'] == n1
        assert abs(o['lh'] - tt.sequence_LH()) < 1e-12 * abs(tt.sequence_LH())
        assert np.array_equal(o['site'], tt.tree.sequence_LH)
        ref_prof = nodes[3].marginal_profile if not nodes[3].is_terminal() else nodes[0].marginal_profile
        assert np.array_equal(o['prof'], ref_prof) and np.array_equal(o['outg'], nodes[5].marginal_outgroup_LH)
        assert o['cseq'] == ''.join(nodes[0].cseq)
    tt.optimize_tree(branch_length_mode='marginal', max_iter=2, prune_short=False)
    bl = np.array([n.branch_length for n in tt.tree.find_clades()])
    tt.infer_gtr(marginal=True)
    for o in outs:
        assert np.allclose(o['bl'][1:], bl[1:], rtol=1e-8, atol=1e-13)     # partial sums are added in a different order
        assert np.isclose(o['lh2'], tt.sequence_LH(), rtol=1e-12)
        assert np.allclose(o['W'], tt.gtr.W, rtol=1e-8)
    assert np.array_equal(outs[0]['bl'], outs[1]['bl'])                      # ranks stay in lock-step
    ref = _extras(TreeAnc, tree, aln, util, dict(engine_factory=oracle_engine.factory))
    for o in outs:
        assert o['s_n'] == ref['s_n'] and o['s_seq'] == ref['s_seq'] and o['s_rng'] == ref['s_rng']
        assert o['m_n'] == ref['m_n'] and np.isclose(o['m_lh'], ref['m_lh'], rtol=1e-12)
        assert np.allclose(o['m_bl'], ref['m_bl'], rtol=1e-7, atol=1e-12)
        assert np.allclose(o['ss_pi'], ref['ss_pi'], rtol=1e-9, atol=1e-13) and np.allclose(o['ss_mu'], ref['ss_mu'], rtol=1e-9)
        assert np.isclose(o['ss_lh'], ref['ss_lh'], rtol=1e-11)
