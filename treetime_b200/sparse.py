"""Sparse forms at the host boundary.

* input: a tip-code matrix as (reference row, list of differences) -- the analogue of TreeTime's
  dict-of-differences (VCF) alignments (sequence_data.py:363-383); uploaded with
  Engine.set_patterns_sparse, expanded on the device;
* output: the reconstructed sequences as (root row, list of states that differ from the parent) --
  what `node.mutations` (treeanc.py:27-42) enumerates; produced by Engine.mutations.
"""
import numpy as np


def sparse_from_dense(tip_codes):
    """(ref_codes[L'], row[], pos[], code[]) with ref = per-column majority code."""
    tip_codes = np.asarray(tip_codes, dtype=np.uint8)
    n_codes = int(tip_codes.max()) + 1 if tip_codes.size else 1
    counts = np.stack([(tip_codes == c).sum(axis=0) for c in range(n_codes)])
    ref = counts.argmax(axis=0).astype(np.uint8)
    row, pos = np.nonzero(tip_codes != ref[None, :])
    return ref, row.astype(np.int32), pos.astype(np.int32), tip_codes[row, pos]


def expand_mutations(parent, tip_row, root_idx, node, pos, state):
    """Dense state indices [n_internal, L'] (internal nodes in node order) from the sparse result."""
    parent = np.asarray(parent)
    internal = np.nonzero(np.asarray(tip_row) < 0)[0]
    slot = -np.ones(parent.shape[0], dtype=np.int64)
    slot[internal] = np.arange(internal.shape[0])
    out = np.empty((internal.shape[0], root_idx.shape[0]), dtype=np.uint8)
    out[0] = root_idx
    start = np.searchsorted(node, internal, side='left')
    stop = np.searchsorted(node, internal, side='right')
    for k, n in enumerate(internal):
        if n == 0:
            continue
        out[k] = out[slot[parent[n]]]               # preorder ids: the parent's row is already final
        lo, hi = start[k], stop[k]
        if hi > lo:
            out[k, pos[lo:hi]] = state[lo:hi]
    return out
