"""Run under torchrun on N GPUs: the pattern-sharded TreeAnc (NCCL collectives) must reproduce the
single-GPU result -- total LH, N_diff, gathered per-node arrays, optimised branch lengths, inferred GTR."""
import os
import sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'tests')):
    sys.path.insert(0, p)
import torch
import torch.distributed as dist
import util
from treetime_b200 import synth
from treetime_b200.dist import TorchComm, SingleComm
from treetime_b200.treeanc import TreeAnc

rank = int(os.environ['RANK']); local = int(os.environ['LOCAL_RANK']); world = int(os.environ['WORLD_SIZE'])
torch.cuda.set_device(local)
dist.init_process_group('nccl', device_id=torch.device('cuda', local))
tree = synth.random_tree(300, seed=5, mean_bl=0.004)
g = util.nuc_gtr()
aln = {k: g.alphabet[v] for k, v in synth.evolve_alignment(tree, 5000, g.Pi, g.W, seed=5).items()}
nwk = tree.to_newick()


def run(comm):
    tt = TreeAnc(tree=nwk, aln=aln, gtr=util.nuc_gtr(), device=local, comm=comm)
    out = dict(n1=tt.infer_ancestral_sequences(marginal=True), lh=tt.sequence_LH(), site=tt.tree.sequence_LH.copy())
    nodes = list(tt.tree.find_clades())
    out['prof'] = nodes[0].marginal_profile.copy()
    out['outg'] = nodes[7].marginal_outgroup_LH.copy()
    out['cseq'] = ''.join(nodes[0].cseq)
    tt.optimize_tree(branch_length_mode='marginal', max_iter=2, prune_short=False)
    out['bl'] = np.array([n.branch_length for n in tt.tree.find_clades()])
    out['lh2'] = tt.sequence_LH()
    tt.infer_gtr(marginal=True)
    out['W'] = np.array(tt.gtr.W)
    out['shard'] = tt._shard()
    return out


sharded = run(TorchComm())
dist.barrier()
if rank == 0:
    single = run(SingleComm())
    assert sharded['n1'] == single['n1']
    assert abs(sharded['lh'] - single['lh']) <= 1e-12 * abs(single['lh'])
    assert np.array_equal(sharded['site'], single['site'])
    assert np.array_equal(sharded['prof'], single['prof']) and np.array_equal(sharded['outg'], single['outg'])
    assert sharded['cseq'] == single['cseq']
    assert np.allclose(sharded['bl'][1:], single['bl'][1:], rtol=1e-7, atol=1e-12), np.abs(sharded['bl'] - single['bl']).max()
    assert abs(sharded['lh2'] - single['lh2']) <= 1e-10 * abs(single['lh2'])
    assert np.allclose(sharded['W'], single['W'], rtol=1e-7)
    print('DIST CHECK OK: world=%d shard0=%s LH=%.6f -> %.6f' % (world, sharded['shard'], single['lh'], single['lh2']))
dist.barrier()
dist.destroy_process_group()
