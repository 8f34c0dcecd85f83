"""Flattening boundary: tree / alignment / GTR objects -> plain arrays.

Duck-typed on purpose: it accepts treetime_b200's own Tree/TreeAnc objects and
the reference's (Bio.Phylo clades held by treetime.TreeAnc) alike.  The arrays
are what crosses the C-ABI (include/ttb.h) and what the CPU oracle consumes.

Node numbering = the reference's preorder `tree.find_clades()` (root = 0), so
children always have larger ids than their parent.  Child order = `node.clades`
order, which fixes the floating point summation order of treeanc.py:861-875.
"""
import numpy as np


def preorder_nodes(root):
    """Iterative preorder in `clades` order (Bio.Phylo find_clades semantics)."""
    out, stack = [], [root]
    while stack:
        n = stack.pop()
        out.append(n)
        stack.extend(reversed(n.clades))
    return out


class FlatTopology(object):
    """parent / CSR children / tip rows for a rooted tree."""

    def __init__(self, root):
        nodes = preorder_nodes(root)
        index = {id(n): i for i, n in enumerate(nodes)}
        n_nodes = len(nodes)
        parent = np.full(n_nodes, -1, dtype=np.int32)
        child_ptr = np.zeros(n_nodes + 1, dtype=np.int32)
        child_idx = np.zeros(max(n_nodes - 1, 0), dtype=np.int32)
        tip_row = np.full(n_nodes, -1, dtype=np.int32)
        k = 0
        n_tips = 0
        for i, n in enumerate(nodes):
            child_ptr[i] = k
            if not n.clades:
                tip_row[i] = n_tips
                n_tips += 1
            for c in n.clades:
                ci = index[id(c)]
                parent[ci] = i
                child_idx[k] = ci
                k += 1
        child_ptr[n_nodes] = k
        self.nodes = nodes
        self.index = index
        self.n_nodes = n_nodes
        self.n_tips = n_tips
        self.parent = parent
        self.child_ptr = child_ptr
        self.child_idx = child_idx
        self.tip_row = tip_row
        self.tip_nodes = np.nonzero(tip_row >= 0)[0].astype(np.int32)
        self.internal_nodes = np.nonzero(tip_row < 0)[0].astype(np.int32)

    def signature(self):
        """Cheap topology fingerprint used to decide whether the device copy is stale."""
        return (self.n_nodes, hash(self.parent.tobytes()), hash(self.child_idx.tobytes()))

    def as_dict(self):
        return dict(parent=self.parent, child_ptr=self.child_ptr, child_idx=self.child_idx, tip_row=self.tip_row)


def code_table(profile_map, n_states):
    """Stable character -> uint8 code assignment and the (n_codes, q) 0/1 table
    (the values of gtr.profile_map, seq_utils.py:28-122).  One extra trailing
    code = all ones, used for tips that have no sequence (treeanc.py:850-851)."""
    chars = sorted(profile_map.keys())
    if len(chars) > 254:
        raise ValueError('too many distinct characters for uint8 tip codes')
    table = np.ones((len(chars) + 1, n_states), dtype=np.float64)
    for i, c in enumerate(chars):
        table[i] = np.asarray(profile_map[c], dtype=np.float64)
    lut = {c: i for i, c in enumerate(chars)}
    return chars, lut, table


def encode_chars(seq, chars):
    """Vectorised char array ('U1'/'S1') -> uint8 codes via a codepoint table."""
    seq = np.asarray(seq)
    if seq.dtype.kind == 'U':
        cp = seq.view(np.uint32).reshape(seq.shape + (-1,))[..., 0] if seq.dtype.itemsize > 4 else seq.view(np.uint32)
    elif seq.dtype.kind == 'S':
        cp = seq.view(np.uint8).astype(np.uint32)
    else:
        raise TypeError('sequence must be a numpy character array')
    cps = np.array([ord(c) for c in chars], dtype=np.uint32)
    lut = np.full(int(max(cps.max(), cp.max())) + 1, 255, dtype=np.uint8)
    lut[cps] = np.arange(len(chars), dtype=np.uint8)
    codes = lut[cp]
    if (codes == 255).any():
        bad = np.unique(seq[codes == 255])
        raise KeyError('characters %s are not in the profile map' % list(bad))
    return codes


def encode_tips(topo, compressed_alignment, profile_map, n_states, Lp):
    """uint8 tip codes [n_tips, L'] in tip_row order + the code profile table."""
    chars, lut, table = code_table(profile_map, n_states)
    missing = len(chars)
    codes = np.empty((topo.n_tips, Lp), dtype=np.uint8)
    for i in topo.tip_nodes:
        node = topo.nodes[i]
        row = topo.tip_row[i]
        if node.name in compressed_alignment:
            s = compressed_alignment[node.name]
            if getattr(s, 'dtype', None) is not None and s.dtype == np.uint8:
                codes[row] = s          # already encoded by treetime_b200.SequenceData
            else:
                codes[row] = encode_chars(s, chars)
        else:
            codes[row] = missing
    return codes, table, chars


def expqt_t_grid(rate_scale):
    """The 61-point grid on which the reference interpolates site-specific exp(Qt)
    (gtr_site_specific.py:336-344)."""
    return (1.0 / rate_scale) * np.concatenate((np.linspace(0, 0.1, 11)[:-1], np.linspace(0.1, 1, 21)[:-1],
                                                np.linspace(1, 5, 21)[:-1], np.linspace(5, 10, 11)))


def gtr_arrays(gtr):
    """GTR object (reference's or ours) -> dict of arrays for the engine/oracle."""
    Pi = np.ascontiguousarray(gtr.Pi, dtype=np.float64)
    d = dict(eigenvals=np.ascontiguousarray(gtr.eigenvals, dtype=np.float64),
             v=np.ascontiguousarray(gtr.v, dtype=np.float64),
             v_inv=np.ascontiguousarray(gtr.v_inv, dtype=np.float64),
             Pi=Pi, gap_index=getattr(gtr, 'gap_index', None))
    if Pi.ndim == 2:
        d['site_specific'] = True
        d['mu'] = np.ascontiguousarray(gtr.mu, dtype=np.float64)
        d['rate_scale'] = float(gtr.rate_scale)
        d['approximate'] = bool(getattr(gtr, 'approximate', True))
        d['t_grid'] = expqt_t_grid(d['rate_scale'])
    else:
        d['site_specific'] = False
        d['mu'] = float(gtr.mu)
    return d


def flatten_treeanc(tt):
    """Everything the oracle / engine needs from a TreeAnc-like object `tt`
    (needs .tree.root, .data.compressed_alignment/.multiplicity()/.compressed_length,
    .gtr and ._branch_length_to_gtr)."""
    topo = FlatTopology(tt.tree.root)
    Lp = int(tt.data.compressed_length)
    q = int(tt.gtr.n_states) if hasattr(tt.gtr, 'n_states') else len(tt.gtr.alphabet)
    codes, table, chars = encode_tips(topo, tt.data.compressed_alignment, tt.gtr.profile_map, q, Lp)
    t = np.array([tt._branch_length_to_gtr(n) for n in topo.nodes], dtype=np.float64)
    flat = topo.as_dict()
    flat.update(tip_codes=codes, code_profiles=table, multiplicity=np.asarray(tt.data.multiplicity(), dtype=np.float64), t=t)
    return topo, flat, gtr_arrays(tt.gtr)
