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
