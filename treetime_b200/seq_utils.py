"""Alphabets and ambiguity profiles for the host side.

Same content as the reference's tables (treetime/seq_utils.py:19-122) -- they
are the IUPAC ambiguity codes -- but built from the IUPAC definitions rather
than written out.  `profile_maps[name][char]` is the 0/1 vector over
`alphabets[name]` of states compatible with `char`.
"""
import numpy as np

alphabet_synonyms = {
    'nuc': 'nuc', 'nucleotide': 'nuc', 'DNA': 'nuc',
    'nuc_nogap': 'nuc_nogap', 'nucleotide_nogap': 'nuc_nogap', 'DNA_nogap': 'nuc_nogap',
    'aa': 'aa', 'aminoacid': 'aa',
    'aa_nogap': 'aa_nogap', 'aminoacid_nogap': 'aa_nogap',
}

_AA = 'ACDEFGHIKLMNPQRSTVWY'
alphabets = {
    'nuc': np.array(list('ACGT-')),
    'nuc_nogap': np.array(list('ACGT')),
    'aa': np.array(list(_AA + '*-')),
    'aa_nogap': np.array(list(_AA)),
}

# IUPAC nucleotide ambiguity codes
_IUPAC_NUC = {
    'A': 'A', 'C': 'C', 'G': 'G', 'T': 'T',
    'R': 'AG', 'Y': 'CT', 'S': 'CG', 'W': 'AT', 'K': 'GT', 'M': 'AC',
    'D': 'AGT', 'H': 'ACT', 'B': 'CGT', 'V': 'ACG',
}
# IUPAC amino-acid ambiguity codes
_IUPAC_AA = {'B': 'ND', 'Z': 'QE'}


def _vec(alphabet, members):
    return np.array([1.0 if a in members else 0.0 for a in alphabet], dtype=float)


def _build_profile_maps():
    maps = {}
    for name in ('nuc', 'nuc_nogap'):
        ab = alphabets[name]
        m = {c: _vec(ab, s) for c, s in _IUPAC_NUC.items()}
        # gap is a state in 'nuc'; in 'nuc_nogap' it is missing data
        m['-'] = _vec(ab, '-') if name == 'nuc' else np.ones(len(ab))
        m['N'] = np.ones(len(ab))
        m['X'] = np.ones(len(ab))
        maps[name] = m
    for name in ('aa', 'aa_nogap'):
        ab = alphabets[name]
        m = {c: _vec(ab, c) for c in _AA}
        if name == 'aa':
            m['*'] = _vec(ab, '*')
            m['-'] = _vec(ab, '-')
        # 'aa_nogap' has no '-' entry: unknown characters are added as all-ones
        # by extend_profile at set-up (treeanc.py:441)
        m['X'] = np.ones(len(ab))
        for c, s in _IUPAC_AA.items():
            m[c] = _vec(ab, s)
        maps[name] = m
    return maps


profile_maps = _build_profile_maps()


def normalize_profile(in_profile, log=False, return_offset=True):
    """Row-normalise an (L, q) profile; with log=True the input holds log
    probabilities.  Host-side helper with the semantics of seq_utils.py:279-307
    (the device kernels do this in registers)."""
    if log:
        pre = in_profile.max(axis=1)
        prof = np.exp(in_profile - pre[:, None])
    else:
        pre = 0.0
        prof = in_profile
    norm = prof.sum(axis=1)
    return prof / norm[:, None], ((np.log(norm) + pre) if return_offset else None)


def seq2prof(seq, profile_map):
    """Character array -> (L, q) profile (seq_utils.py:207-229)."""
    return np.array([profile_map[k] for k in seq])


def prof2seq(profile, gtr, sample_from_prof=False, normalize=True, rng=None):
    """(L, q) profile -> (sequence, values, indices); argmax = first maximum, or
    inverse-CDF sampling with one uniform per row (seq_utils.py:232-276)."""
    if rng is None:
        rng = np.random.default_rng()
    tmp = normalize_profile(profile, return_offset=False)[0] if normalize else profile
    if sample_from_prof:
        cumdis = tmp.cumsum(axis=1).T
        randnum = rng.random(size=cumdis.shape[1])
        idx = np.argmax(cumdis >= randnum, axis=0)
    else:
        idx = tmp.argmax(axis=1)
    return gtr.alphabet[idx], tmp[np.arange(tmp.shape[0]), idx], idx
