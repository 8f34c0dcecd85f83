"""Host side of the joint branch-length optimisation: turn the device's per-branch
pair-count tables (ttb_branch_state_pairs) into the reference's `branch_state` dictionaries and
optimise all branch lengths in one lock-step Brent.

Reference: TreeAnc.add_branch_state (treeanc.py:1148-1163), GTR.state_pair (gtr.py:631-705),
GTR.prob_t_compressed (gtr.py:710-745), GTR.optimal_t_compressed (gtr.py:816-925, profiles=False)."""
import numpy as np

from . import config as ttconf
from .brent import brent_lockstep

NONE = 0x7fffffff


def fold_state_pairs(C, F, alphabet, col_chars, gap_index, ignore_gaps):
    """One branch: C[q, W] multiplicity sums, F[q, W] first pattern of (parent state i, child column c);
    col_chars[c] = character shown by column c (None: no character).  Returns (pairs[n, 2] int,
    multiplicity[n] int) with the content AND order of GTR.state_pair:
      * alphabets of < 10 letters: pairs (n1, n2) in alphabet order, positions whose child character is
        not a letter of the alphabet are not counted, gaps skipped with ignore_gaps (gtr.py:676-689);
      * larger alphabets: characters outside the alphabet count as state 1 (`np.ones_like`, :694-698),
        pairs listed in order of first occurrence (dict insertion order, :699-705)."""
    q = len(alphabet)
    letter = {c: i for i, c in enumerate(alphabet)}
    out = []
    if q < 10:
        col_of = {}
        for c, ch in enumerate(col_chars):
            if ch in letter and letter[ch] not in col_of:
                col_of[letter[ch]] = c
        for n1 in range(q):
            if gap_index is None or not ignore_gaps or n1 != gap_index:
                for n2 in range(q):
                    if (gap_index is None or not ignore_gaps or n2 != gap_index) and n2 in col_of:
                        count = C[n1, col_of[n2]]
                        if count:
                            out.append(((n1, n2), count))
    else:
        acc = {}
        ii, cc = np.nonzero(F != NONE)
        for i, c in zip(ii, cc):
            if c >= len(col_chars) or col_chars[c] is None:
                continue
            s = letter.get(col_chars[c], 1)
            if ignore_gaps and (gap_index == i or gap_index == s):
                continue
            cnt, fst = acc.get((i, s), (0.0, NONE))
            acc[(i, s)] = (cnt + C[i, c], min(fst, int(F[i, c])))
        out = [(k, v[0]) for k, v in sorted(acc.items(), key=lambda kv: kv[1][1])]
    return (np.array([x[0] for x in out], dtype=int).reshape(-1, 2) if out else np.zeros((0, 2), dtype=int),
            np.array([x[1] for x in out], dtype=int))


def count_matrices(C, alphabet, col_chars, gap_index, ignore_gaps):
    """All branches at once: C[n, q, W] -> M[n, parent, child] with the same counting rules as
    fold_state_pairs (order is irrelevant for the dense form)."""
    q = len(alphabet)
    letter = {c: i for i, c in enumerate(alphabet)}
    M = np.zeros((C.shape[0], q, q))
    for c, ch in enumerate(col_chars):
        if c >= C.shape[2] or ch is None:
            continue
        if ch in letter:
            s = letter[ch]
        elif q < 10:
            continue
        else:
            s = 1
        M[:, :, s] += C[:, :, c]
    M = np.trunc(M)                      # state_pair returns integer multiplicities (:704)
    if ignore_gaps and gap_index is not None:
        M[:, gap_index, :] = 0
        M[:, :, gap_index] = 0
    return M


def optimal_t_from_counts(gtr, M, tol=1e-10):
    """GTR.optimal_t_compressed(profiles=False) for many branches: minimise over s = sqrt(t)
       -sum_{p,c} M[b, p, c] * log(max(expQt(s^2)[c, p], SUPERTINY))
    with the bracket (-sqrt(MAX_BL), sqrt(hamming), sqrt(MAX_BL)) (gtr.py:866-885); failures fall back
    to the hamming distance (:916-918).  M[b, parent, child]."""
    n = M.shape[0]
    tot = M.sum(axis=(1, 2))
    diag = np.einsum('bii->b', M)
    with np.errstate(invalid='ignore', divide='ignore'):
        hamming = (tot - diag) / tot
    v, vinv, lam = np.asarray(gtr.v), np.asarray(gtr.v_inv), np.asarray(gtr.eigenvals) * gtr.mu
    Mt = np.ascontiguousarray(np.transpose(M, (0, 2, 1)))        # [b, child, parent] like logQt[child, parent]

    def neg_prob(idx, s):
        e = np.exp(np.multiply.outer(s ** 2, lam))               # [k, q]
        P = np.maximum(0.0, np.einsum('ij,bj,jk->bik', v, e, vinv))
        logP = np.log(np.maximum(P, ttconf.SUPERTINY_NUMBER))
        return -1.0 * (logP * Mt[idx]).sum(axis=(1, 2))

    smax = np.sqrt(ttconf.MAX_BRANCH_LENGTH)
    with np.errstate(invalid='ignore'):
        xb = np.sqrt(hamming)
    opt = brent_lockstep(neg_prob, np.full(n, -smax), xb, np.full(n, smax), tol=tol)
    new_len = opt['x'] ** 2
    return np.where(opt['success'], new_len, hamming), opt
