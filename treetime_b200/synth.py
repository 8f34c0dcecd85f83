"""Deterministic synthetic inputs (numpy only): random trees and alignments
evolved along them.  Used by the parity tests and by bench.py on the GPU box,
where neither the reference nor its SeqGen exist.  The generator is a fixed
function of (numpy version, seed): tests pin a checksum of its output.

`evolve_alignment` simulates the same continuous-time Markov chain as the
reference's SeqGen (seqgen.py:38-67: child ~ expQt(t)[:, parent] per site) but
by uniformisation, touching only the sites that mutate, so a 20k-tip x 30 kb
alignment takes seconds instead of an hour.
"""
import numpy as np
from .tree import Node, Tree


def random_tree(n_tips, seed=1, mean_bl=1e-3, polytomy_frac=0.0, zero_frac=0.0):
    """Random rooted tree: repeatedly join two random lineages; exponential
    branch lengths with mean `mean_bl`; tips named t000000...  With
    `polytomy_frac`>0 a fraction of internal branches is collapsed (children
    spliced into the grandparent) to create multifurcations; `zero_frac` sets a
    fraction of branch lengths to exactly 0."""
    rng = np.random.default_rng(seed)
    lineages = [Node(name='t%06d' % i) for i in range(n_tips)]
    bls = rng.exponential(mean_bl, size=2 * n_tips)
    if zero_frac > 0:
        bls[rng.random(2 * n_tips) < zero_frac] = 0.0
    k = 0
    while len(lineages) > 1:
        i = int(rng.integers(len(lineages)))
        a = lineages[i]
        lineages[i] = lineages[-1]
        lineages.pop()
        j = int(rng.integers(len(lineages)))
        b = lineages[j]
        a.branch_length = float(bls[k]); b.branch_length = float(bls[k + 1]); k += 2
        lineages[j] = Node(clades=[a, b])
    root = lineages[0]
    root.branch_length = None
    if polytomy_frac > 0:
        # collapse internal, non-root branches top-down; every internal node is drawn exactly once
        stack = [root]
        while stack:
            n = stack.pop()
            queue, new = list(n.clades), []
            while queue:
                c = queue.pop(0)
                if c.clades and rng.random() < polytomy_frac:
                    for g in c.clades:
                        g.branch_length = (g.branch_length or 0.0) + (c.branch_length or 0.0)
                    queue.extend(c.clades)
                else:
                    new.append(c)
            n.clades = new
            stack.extend(c for c in n.clades if c.clades)
    return Tree(root=root)


def caterpillar_tree(n_tips, mean_bl=1e-3, seed=1):
    """Maximally unbalanced (ladder) tree: depth = n_tips - 1 levels."""
    rng = np.random.default_rng(seed)
    bls = rng.exponential(mean_bl, size=2 * n_tips)
    cur = Node(name='t%06d' % 0, branch_length=float(bls[0]))
    for i in range(1, n_tips):
        tip = Node(name='t%06d' % i, branch_length=float(bls[2 * i]))
        cur = Node(clades=[tip, cur], branch_length=float(bls[2 * i + 1]))
    cur.branch_length = None
    return Tree(root=cur)


def evolve_alignment(tree, L, Pi, W, mu=1.0, seed=1, dtype=np.uint8):
    """State-index alignment (dict tip name -> uint8[L] of state indices) evolved
    down `tree` under rate matrix Q_ij = mu W_ij Pi_i (i != j; column j = from
    state j, as in gtr.py:289-299), root drawn from Pi.  Exact CTMC sampling by
    uniformisation: per branch Poisson(Lambda t L) candidate events at uniformly
    random sites, each moving state j -> i with probability Q_ij/Lambda (else a
    self-transition)."""
    rng = np.random.default_rng(seed)
    Pi = np.asarray(Pi, dtype=float)
    q = Pi.shape[0]
    Q = mu * np.asarray(W, dtype=float) * Pi[:, None]
    np.fill_diagonal(Q, 0.0)
    out_rate = Q.sum(axis=0)
    Lam = out_rate.max() * 1.000001 + 1e-300
    # per from-state j: cumulative distribution over target i (self-transition last)
    cum = np.zeros((q, q + 1))
    for j in range(q):
        p = np.concatenate([Q[:, j] / Lam, [1.0 - out_rate[j] / Lam]])
        cum[j] = np.cumsum(p)
    root_seq = np.searchsorted(np.cumsum(Pi), rng.random(L)).clip(0, q - 1).astype(dtype)
    aln = {}
    stack = [(tree.root, root_seq)]
    while stack:
        node, seq = stack.pop()
        if not node.clades:
            aln[node.name] = seq
            continue
        for c in node.clades:
            cs = seq.copy()
            n_ev = rng.poisson(Lam * (c.branch_length or 0.0) * L)
            if n_ev:
                sites = rng.integers(0, L, size=n_ev)
                u = rng.random(n_ev)
                for s, uu in zip(sites, u):       # sequential: events may hit the same site twice
                    j = cs[s]
                    i = int(np.searchsorted(cum[j], uu))
                    if i < q:
                        cs[s] = i
            stack.append((c, cs))
    return aln


def sprinkle_ambiguous(aln_chars, frac, ambiguous_chars, seed=1):
    """Replace a fraction of characters by ambiguity codes (in place copy)."""
    rng = np.random.default_rng(seed)
    out = {}
    amb = np.array(list(ambiguous_chars))
    for k in sorted(aln_chars):
        s = np.array(aln_chars[k]).copy()
        m = rng.random(s.shape[0]) < frac
        s[m] = amb[rng.integers(0, len(amb), size=int(m.sum()))]
        out[k] = s
    return out


def flat_problem(tree, sd, gtr):
    """tree + SequenceData + GTR -> (FlatTopology, flat dict, gtr dict): the arrays that cross
    the C-ABI (include/ttb.h) and that the CPU oracle consumes.  Mirrors what
    TreeAnc._sync_device uploads: ladderized child order (treeanc.py:456-457), branch lengths
    floored at MIN_BRANCH_LENGTH * one_mutation (treeanc.py:752-760)."""
    from . import config as ttconf
    from .flatten import FlatTopology, code_table, gtr_arrays
    tree.ladderize()
    topo = FlatTopology(tree.root)
    chars, lut, table = code_table(gtr.profile_map, gtr.n_states)
    lut8 = np.full(256, 255, dtype=np.uint8)
    for c, i in lut.items():
        lut8[ord(c)] = i
    rows = np.array([sd._row.get(topo.nodes[n].name, -1) for n in topo.tip_nodes])
    codes = np.full((topo.n_tips, sd.compressed_length), len(chars), dtype=np.uint8)
    have = rows >= 0
    codes[have] = lut8[sd.compressed_matrix[rows[have]]]
    if (codes == 255).any():
        raise KeyError('alignment contains characters that are not in the profile map')
    one_mutation = 1.0 / sd.full_length
    floor = ttconf.MIN_BRANCH_LENGTH * one_mutation
    t = np.array([max(floor, n.branch_length if n.branch_length else 0.0) for n in topo.nodes])
    t[0] = max(floor, 0.001)
    flat = topo.as_dict()
    flat.update(tip_codes=codes, code_profiles=table, multiplicity=np.array(sd.multiplicity(), dtype=np.float64), t=t)
    return topo, flat, gtr_arrays(gtr)


def make_flat_problem(tree, gtr, L, seed, amb_frac=0.0, amb_chars='N-RY', compress=True, mu_sim=1.0):
    """Simulate an alignment of length L down `tree` and flatten everything."""
    from .sequence_data import SequenceData
    Pi = gtr.Pi if np.ndim(gtr.Pi) == 1 else gtr.Pi.mean(axis=1)
    idx = evolve_alignment(tree, L, Pi, gtr.W, mu=mu_sim, seed=seed)
    ab = np.asarray(gtr.alphabet).astype('S1').view(np.uint8)
    aln = {k: ab[v] for k, v in idx.items()}                      # ASCII bytes
    if amb_frac:
        amb = np.frombuffer(amb_chars.encode('ascii'), dtype=np.uint8)
        rng = np.random.default_rng(seed + 1)
        for k in sorted(aln):
            s = aln[k]
            m = rng.random(s.shape[0]) < amb_frac
            s[m] = amb[rng.integers(0, len(amb), size=int(m.sum()))]
    sd = SequenceData(aln, compress=compress, ambiguous=gtr.ambiguous)
    return flat_problem(tree, sd, gtr)
