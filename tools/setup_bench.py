"""N3 measurement: TreeAnc set-up (alignment -> patterns on the device) with host vs device pattern compression."""
import sys, time; sys.path.insert(0, '.')
import numpy as np
from treetime_b200 import synth
from treetime_b200.gtr import GTR
from treetime_b200.treeanc import TreeAnc
n_tips, L, mbl = (int(sys.argv[1]), int(sys.argv[2]), float(sys.argv[3])) if len(sys.argv) > 3 else (2000, 10000, 5e-4)
g = GTR.custom(pi=np.array([.3, .2, .2, .29, .01]), W=np.ones((5, 5)), alphabet='nuc')
tree = synth.random_tree(n_tips, seed=1, mean_bl=mbl)
idx = synth.evolve_alignment(tree, L, g.Pi, g.W, seed=1)
ab = np.asarray(g.alphabet).astype('S1').view(np.uint8)
aln = {k: ab[v] for k, v in idx.items()}          # ASCII byte rows
nwk = tree.to_newick()
res = {}
for dc in (False, True, False, True):
    t0 = time.perf_counter()
    tt = TreeAnc(tree=nwk, aln=aln, gtr=g, device_compress=dc)
    t1 = time.perf_counter()
    tt._sync_device(); tt._engine.sync()
    t2 = time.perf_counter()
    tt.infer_ancestral_sequences(marginal=True)
    t3 = time.perf_counter()
    res[dc] = (t1 - t0, t2 - t1, t3 - t2, tt.data.compressed_length, tt.sequence_LH())
    print('device_compress=%-5s  constructor %.2f s  first upload %.2f s  first pass %.3f s  L\'=%d  LH=%.6f' % ((dc,) + res[dc]))
    tt._engine.close()
assert res[False][3:] == res[True][3:]
