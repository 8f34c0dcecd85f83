"""Alignment container and pattern compression for the host side.

Input boundary of the hot path: turns an alignment
into the compressed pattern matrix the engine consumes.  Same semantics as the
reference's SequenceData.make_compressed_alignment (sequence_data.py:325-464):
  * columns are visited in order; a column that is constant -- possibly after
    replacing the ambiguous character when exactly one other letter occurs
    (:386-392) -- shares one pattern with all identical constant columns;
  * every variable column gets a private pattern (:394-402);
  * patterns are numbered by first occurrence; multiplicity = number of columns.
Unlike the reference (a Python loop over columns building strings) the columns
are classified with vectorised byte arithmetic on an (n_seq, L) uint8 matrix.
"""
import numpy as np


def _to_bytes(seq, convert_upper=True):
    """str / char array / bytes -> uint8 ASCII codes."""
    if isinstance(seq, str):
        b = np.frombuffer(seq.encode('ascii'), dtype=np.uint8)
    elif isinstance(seq, (bytes, bytearray)):
        b = np.frombuffer(bytes(seq), dtype=np.uint8)
    else:
        seq = np.asarray(seq)
        if seq.dtype.kind == 'U':
            b = seq.astype('U1').view(np.uint32).astype(np.uint8)
        elif seq.dtype.kind == 'S':
            b = seq.astype('S1').view(np.uint8)
        elif seq.dtype == np.uint8:
            b = seq
        else:
            raise TypeError('unsupported sequence type %r' % (seq.dtype,))
    if convert_upper:
        b = np.where((b >= 97) & (b <= 122), b - 32, b).astype(np.uint8)
    return b


def _blocks(L, width=2048):
    return [(a, min(L, a + width)) for a in range(0, L, width)]


def _pool_map(fn, jobs):
    """numpy releases the GIL inside its loops: a few threads over column blocks keep the temporaries cache-sized and use
    the host's cores (the column scans are the host side of N3)."""
    import os
    from concurrent.futures import ThreadPoolExecutor
    n = max(1, min(len(jobs), os.cpu_count() or 1, 16))
    if n == 1:
        return [fn(j) for j in jobs]
    with ThreadPoolExecutor(max_workers=n) as ex:
        return list(ex.map(fn, jobs))


def _column_stats(A, amb):
    """Per column: min and max over the entries != amb (255 / 0 if there is none) and whether all entries == amb."""
    L = A.shape[1]
    lo = np.empty(L, dtype=np.uint8); hi = np.empty(L, dtype=np.uint8); all_amb = np.empty(L, dtype=bool)

    def one(b):
        a0, a1 = b
        X = A[:, a0:a1]
        is_amb = X == amb
        lo[a0:a1] = np.where(is_amb, 255, X).min(axis=0)
        hi[a0:a1] = np.where(is_amb, 0, X).max(axis=0)
        all_amb[a0:a1] = is_amb.all(axis=0)
    _pool_map(one, _blocks(L))
    return lo, hi, all_amb


def _gather_columns(A, cols):
    """A[:, cols] as a new C-contiguous matrix, row blocks in parallel."""
    cols = np.asarray(cols)
    out = np.empty((A.shape[0], cols.shape[0]), dtype=A.dtype)

    def one(b):
        r0, r1 = b
        np.take(A[r0:r1], cols, axis=1, out=out[r0:r1])
    _pool_map(one, _blocks(A.shape[0], 512))
    return out


def read_fasta(path):
    names, seqs, cur = [], [], []
    with open(path) as fh:
        for line in fh:
            line = line.strip()
            if not line:
                continue
            if line[0] == '>':
                if names:
                    seqs.append(''.join(cur))
                names.append(line[1:].split()[0])
                cur = []
            else:
                cur.append(line)
    if names:
        seqs.append(''.join(cur))
    return list(zip(names, seqs))


class SequenceData(object):
    """Holds the alignment as ASCII bytes and its compressed patterns.

    aln : fasta file name, dict name -> sequence, or list of (name, sequence);
          sequences may be str, numpy char arrays or uint8 ASCII arrays.
    """

    def __init__(self, aln, compress=True, convert_upper=True, fill_overhangs=True, ambiguous='N',
                 sequence_length=None, logger=None, device_stats=None):
        """device_stats: optional callable (matrix, fill_overhangs, gap, fill, ambiguous) -> (lo, hi, all_amb)
        that computes the per-column statistics on the GPU (Engine.alignment_stats) and
        keeps the alignment resident there; the host then only numbers the patterns."""
        self.logger = logger or (lambda *a, **k: None)
        self.compress = compress
        self.ambiguous = ambiguous
        self.is_sparse = False
        self.additional_constant_sites = 0
        if isinstance(aln, str):
            aln = read_fasta(aln)
        if isinstance(aln, dict):
            items = list(aln.items())
        else:
            items = [(getattr(r, 'id', None) or r[0], str(getattr(r, 'seq', None) or r[1])) if not isinstance(r, tuple) else r
                     for r in aln]
        if not items:
            raise ValueError('SequenceData: empty alignment')
        self.sequence_names = [k for k, _ in items]
        if all(isinstance(s, np.ndarray) and s.dtype == np.uint8 for _, s in items):
            rows = [s for _, s in items]              # already ASCII bytes: upper-case once, vectorised, below
            bulk_upper = convert_upper
        else:
            rows = [_to_bytes(s, convert_upper) for _, s in items]
            bulk_upper = False
        L = rows[0].shape[0]
        if any(r.shape[0] != L for r in rows):
            raise ValueError('SequenceData: sequences differ in length')
        self._matrix = np.vstack(rows) if len(rows) > 1 else rows[0][None, :].copy()
        if bulk_upper and self._matrix.max() >= 97:        # one pass without temporaries decides whether any letter is lower case
            lower = (self._matrix >= 97) & (self._matrix <= 122)
            if lower.any():
                self._matrix[lower] -= 32
        # with no ambiguous character the reference assigns None into a 'U1' array, which numpy stores
        # as 'N' (seq_utils.py:196-202): reproduce that
        self._fill_char = (ord(ambiguous) if ambiguous is not None else ord('N')) if fill_overhangs else None
        self._filled = self._fill_char is None
        self._col_stats = None
        self.device_resident = False
        if self.ambiguous is None:                      # sequence_data.py:322-323
            nuc = np.isin(self._matrix, np.frombuffer(b'acgtACGT-N', dtype=np.uint8)).sum()
            self.ambiguous = 'N' if nuc > 0.9 * self._matrix.size else 'X'
        if device_stats is not None and compress and not (sequence_length and int(sequence_length) > L):
            self._col_stats = device_stats(self._matrix, self._fill_char is not None, '-',
                                           chr(self._fill_char) if self._fill_char is not None else 'N', self.ambiguous)
            self.device_resident = True
        self.full_length = int(sequence_length) if sequence_length else L
        if self.full_length < L:
            raise AttributeError('SequenceData: specified sequence length is smaller than alignment length!')
        self.additional_constant_sites = self.full_length - L
        self._row = {k: i for i, k in enumerate(self.sequence_names)}
        self.make_compressed_alignment()

    @property
    def matrix(self):
        """Host copy of the alignment with overhangs filled (lazily when the device did the statistics)."""
        if not self._filled:
            self._fill_overhangs(self._fill_char)
            self._filled = True
        return self._matrix

    def _fill_overhangs(self, amb):
        """Leading/trailing gaps -> ambiguous (seq2array fill_overhangs, seq_utils.py:196-202)."""
        A = self._matrix
        gap = ord('-')
        # only rows that start or end with a gap have overhangs: no full-matrix temporaries for the others
        for r in np.nonzero((A[:, 0] == gap) | (A[:, -1] == gap))[0]:
            row = A[r]
            ng = np.nonzero(row != gap)[0]
            if ng.size == 0:
                row[:] = amb
            else:
                row[:ng[0]] = amb
                row[ng[-1] + 1:] = amb

    @property
    def aln(self):
        return _AlnView(self)

    @property
    def compressed_length(self):
        return self._compressed_length

    def multiplicity(self, mask=None):
        return self._multiplicity if mask is None else self._multiplicity * mask

    def make_compressed_alignment(self):
        n_seq, L = self._matrix.shape
        if self._col_stats is not None:
            self._compress_from_stats(*self._col_stats)
            return
        A = self.matrix
        if not self.compress:
            self._multiplicity = np.ones(self.full_length, dtype=float)
            self.full_to_compressed_sequence_map = np.arange(self.full_length)
            self._compressed_length = self.full_length
            self.compressed_matrix = A
            self.pattern_first_position = np.arange(L)
            self._finish()
            return
        amb = ord(self.ambiguous) if self.ambiguous is not None else 256
        lo, hi, all_amb = _column_stats(A, amb)        # extrema over non-ambiguous entries
        # constant (possibly after replacing the ambiguous character by the single other letter)
        const = (lo == hi) | all_amb
        letter = np.where(all_amb, amb, lo).astype(np.uint8)
        # pattern id by first occurrence: variable columns are private, constant ones keyed by letter
        pid = np.empty(L, dtype=np.int64)
        first_seen = {}
        order = []
        key = np.where(const, letter.astype(np.int64), -1 - np.arange(L))
        uniq, first_idx, inverse = np.unique(key, return_index=True, return_inverse=True)
        rank = np.argsort(np.argsort(first_idx))       # patterns ordered by first occurrence
        pid = rank[inverse]
        n_pat = uniq.shape[0]
        first_pos = np.empty(n_pat, dtype=np.int64)
        first_pos[rank] = first_idx
        C = _gather_columns(A, first_pos)
        cc = const[first_pos]
        if cc.any():                                    # constant patterns: ambiguous replaced (:391)
            C[:, cc] = letter[first_pos][cc][None, :]
        mult = np.bincount(pid, minlength=n_pat).astype(float)
        f2c = pid.copy()
        if self.additional_constant_sites:
            # extra constant columns distributed over the unambiguous states by composition (:417-441)
            from .seq_utils import alphabets
            likely = 'nuc' if np.isin(A, np.frombuffer(b'ACGT-N', dtype=np.uint8)).mean() > 0.9 else 'aa'
            chars = [c for c in alphabets[likely + '_nogap'] if c not in (self.ambiguous, '-')]
            counts = [(c, int((A == ord(c)).sum())) for c in chars]
            total = sum(n for _, n in counts)
            left = self.additional_constant_sites
            extra_f2c = []
            for k, (c, n) in enumerate(counts):
                add = left if k == len(counts) - 1 else int(np.round(self.additional_constant_sites * n / total))
                if add:
                    hit = np.nonzero(cc & (C[0] == ord(c)) & (C == ord(c)).all(axis=0))[0]
                    if hit.size:
                        p = int(hit[0])
                    else:
                        p = C.shape[1]
                        C = np.hstack([C, np.full((n_seq, 1), ord(c), dtype=np.uint8)])
                        cc = np.append(cc, True)
                        mult = np.append(mult, 0.0)
                        first_pos = np.append(first_pos, L + len(extra_f2c))
                    mult[p] += add
                    extra_f2c.extend([p] * add)
                    left -= add
            f2c = np.concatenate([f2c, np.array(extra_f2c, dtype=np.int64)])
        self._multiplicity = mult
        self.full_to_compressed_sequence_map = f2c
        self._compressed_length = int(mult.shape[0])
        self.compressed_matrix = np.ascontiguousarray(C)
        self.pattern_first_position = first_pos
        self._finish()

    def _compress_from_stats(self, lo, hi, all_amb):
        """Number the patterns from per-column statistics computed on the device (same rules as the
        host path below; sequence_data.py:386-402)."""
        L = lo.shape[0]
        amb = ord(self.ambiguous) if self.ambiguous is not None else 256
        const = (lo == hi) | all_amb
        letter = np.where(all_amb, amb, lo).astype(np.uint8)
        key = np.where(const, letter.astype(np.int64), -1 - np.arange(L))
        uniq, first_idx, inverse = np.unique(key, return_index=True, return_inverse=True)
        rank = np.argsort(np.argsort(first_idx))
        pid = rank[inverse]
        n_pat = uniq.shape[0]
        first_pos = np.empty(n_pat, dtype=np.int64)
        first_pos[rank] = first_idx
        self._multiplicity = np.bincount(pid, minlength=n_pat).astype(float)
        self.full_to_compressed_sequence_map = pid
        self._compressed_length = int(n_pat)
        self.pattern_first_position = first_pos
        self.pattern_const_letter = np.where(const[first_pos], letter[first_pos], 0).astype(np.uint8)
        self._compressed_matrix = None                 # gathered lazily on the host if anybody asks
        self._finish()

    @property
    def compressed_matrix(self):
        if self._compressed_matrix is None:
            C = _gather_columns(self.matrix, self.pattern_first_position)
            cc = self.pattern_const_letter != 0
            if cc.any():
                C[:, cc] = self.pattern_const_letter[cc][None, :]
            self._compressed_matrix = np.ascontiguousarray(C)
        return self._compressed_matrix

    @compressed_matrix.setter
    def compressed_matrix(self, value):
        self._compressed_matrix = value

    def _finish(self):
        self.compressed_alignment = _CompressedView(self)

    def compressed_to_full_sequence(self, sequence, include_additional_constant_sites=False, as_string=False):
        """Expand a compressed sequence (sequence_data.py:513-543)."""
        L = self.full_length if include_additional_constant_sites else self.full_length - self.additional_constant_sites
        tmp = np.asarray(sequence)[self.full_to_compressed_sequence_map[:L]]
        return ''.join(tmp.astype('U')) if as_string else tmp


class _CompressedView(object):
    """dict-like view name -> compressed char array (like data.compressed_alignment)."""

    def __init__(self, sd):
        self._sd = sd

    def __contains__(self, name):
        return name in self._sd._row

    def __getitem__(self, name):
        return self._sd.compressed_matrix[self._sd._row[name]].view('S1').astype('U1')

    def codes(self, name):
        return self._sd.compressed_matrix[self._sd._row[name]]

    def keys(self):
        return self._sd.sequence_names

    def __len__(self):
        return len(self._sd.sequence_names)


class _AlnView(_CompressedView):
    def __getitem__(self, name):
        return self._sd.matrix[self._sd._row[name]].view('S1').astype('U1')

    def values(self):
        return [self[k] for k in self.keys()]
