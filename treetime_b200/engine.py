"""Python handle on one device engine (one GPU, one shard of the pattern axis)."""
import ctypes
import numpy as np
from . import _lib

SUBTREE, OUTGROUP, PROFILE, JOINT_ROOT_LX = 0, 1, 2, 3
RECONSTRUCT_TIPS, LH_ONLY, KEEP_PREV_STATES = 1, 2, 32
BRANCH, BRANCH_ROOT = 0, 1


def _dp(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


def _ip(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_int32))


def _up(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8))


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


class Engine(object):
    """Owns a ttb_handle.  All array arguments are numpy (host) arrays."""

    def __init__(self, n_states, device=0):
        self.lib = _lib.load()
        self.n_states = int(n_states)
        self.device = int(device)
        h = ctypes.c_void_p()
        _lib.check(self.lib.ttb_create(ctypes.byref(h), self.device, self.n_states))
        self.h = h
        self.n_nodes = 0
        self.n_patterns = 0
        self.n_codes = 0
        self._keep = {}

    def close(self):
        if getattr(self, 'h', None):
            self.lib.ttb_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- inputs ---------------------------------------------------------------
    def set_stream(self, cuda_stream):
        _lib.check(self.lib.ttb_set_stream(self.h, ctypes.c_void_p(int(cuda_stream) if cuda_stream else None)))

    def set_message_storage(self, dtype):
        """'f64' (default, the reference's precision) or 'f32': S / M stored as float, arithmetic in double."""
        code = {'f64': 0, 'float64': 0, 'f32': 1, 'float32': 1}.get(str(dtype))
        if code is None:
            raise ValueError("message storage must be 'f64' or 'f32'")
        _lib.check(self.lib.ttb_set_message_storage(self.h, code))

    def set_tree(self, parent, child_ptr, child_idx, tip_row):
        parent, child_ptr, child_idx, tip_row = _i32(parent), _i32(child_ptr), _i32(child_idx), _i32(tip_row)
        _lib.check(self.lib.ttb_set_tree(self.h, parent.shape[0], _ip(parent), _ip(child_ptr), _ip(child_idx), _ip(tip_row)))
        self.n_nodes = int(parent.shape[0])
        self.tip_row = tip_row

    def set_patterns(self, tip_codes, code_profiles, multiplicity, validate=True):
        """validate=False skips the host-side range scan of the code matrix (it costs a pass over it)."""
        tip_codes = np.ascontiguousarray(tip_codes, dtype=np.uint8)
        code_profiles = _f64(code_profiles)
        multiplicity = _f64(multiplicity)
        if code_profiles.shape[1] != self.n_states:
            raise ValueError('code_profiles must have n_states columns')
        if tip_codes.shape[1] != multiplicity.shape[0]:
            raise ValueError('tip_codes and multiplicity disagree on the number of patterns')
        if validate and tip_codes.size and int(tip_codes.max()) >= code_profiles.shape[0]:
            raise ValueError('tip code out of range of code_profiles')
        _lib.check(self.lib.ttb_set_patterns(self.h, tip_codes.shape[1], _up(tip_codes), code_profiles.shape[0],
                                             _dp(code_profiles), _dp(multiplicity)))
        self.n_patterns = int(tip_codes.shape[1])
        self.n_codes = int(code_profiles.shape[0])

    def alignment_stats(self, aln, fill_overhangs=False, gap='-', fill='N', ambiguous='N'):
        """Upload the raw ASCII alignment [n_seq, L] (kept resident) and return per column
        (lo, hi, all_ambiguous): extrema over the non-ambiguous characters."""
        aln = np.ascontiguousarray(aln, dtype=np.uint8)
        L = aln.shape[1]
        lo = np.empty(L, dtype=np.uint8); hi = np.empty(L, dtype=np.uint8); aa = np.empty(L, dtype=np.uint8)
        amb = ord(ambiguous) if ambiguous is not None else 256
        _lib.check(self.lib.ttb_alignment_stats(self.h, aln.shape[0], L, _up(aln), 1 if fill_overhangs else 0, ord(gap),
                                                ord(fill), amb, _up(lo), _up(hi), _up(aa)))
        return lo, hi, aa.astype(bool)

    def set_patterns_from_alignment(self, first_pos, const_letter, tip_seq_row, lut, missing_code, code_profiles, multiplicity):
        """Gather the compressed code matrix on the device from the resident alignment."""
        first_pos = np.ascontiguousarray(first_pos, dtype=np.int64)
        const_letter = np.ascontiguousarray(const_letter, dtype=np.uint8)
        tip_seq_row = _i32(tip_seq_row)
        lut = np.ascontiguousarray(lut, dtype=np.uint8)
        code_profiles, multiplicity = _f64(code_profiles), _f64(multiplicity)
        _lib.check(self.lib.ttb_set_patterns_from_alignment(
            self.h, first_pos.shape[0], first_pos.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)), _up(const_letter), _ip(tip_seq_row),
            _up(lut), int(missing_code), code_profiles.shape[0], _dp(code_profiles), _dp(multiplicity)))
        self.n_patterns = int(first_pos.shape[0])
        self.n_codes = int(np.asarray(code_profiles).shape[0])

    def set_patterns_sparse(self, ref_codes, entry_row, entry_pos, entry_code, code_profiles, multiplicity):
        """Every tip row = ref_codes except at the listed (tip row, pattern, code) entries."""
        ref_codes = np.ascontiguousarray(ref_codes, dtype=np.uint8)
        entry_row, entry_pos = _i32(entry_row), _i32(entry_pos)
        entry_code = np.ascontiguousarray(entry_code, dtype=np.uint8)
        code_profiles, multiplicity = _f64(code_profiles), _f64(multiplicity)
        if code_profiles.shape[1] != self.n_states or ref_codes.shape[0] != multiplicity.shape[0]:
            raise ValueError('inconsistent sparse pattern arguments')
        _lib.check(self.lib.ttb_set_patterns_sparse(self.h, ref_codes.shape[0], _up(ref_codes), entry_row.shape[0],
                                                    _ip(entry_row), _ip(entry_pos), _up(entry_code), code_profiles.shape[0],
                                                    _dp(code_profiles), _dp(multiplicity)))
        self.n_patterns = int(ref_codes.shape[0])
        self.n_codes = int(np.asarray(code_profiles).shape[0])

    def set_gtr(self, g):
        """g: dict from flatten.gtr_arrays()."""
        gap = -1 if g.get('gap_index') is None else int(g['gap_index'])
        if g.get('site_specific', False):
            ev, v, vi, Pi, mu = _f64(g['eigenvals']), _f64(g['v']), _f64(g['v_inv']), _f64(g['Pi']), _f64(g['mu'])
            tg = _f64(g['t_grid'])
            _lib.check(self.lib.ttb_set_gtr_site_specific(self.h, _dp(ev), _dp(v), _dp(vi), _dp(Pi), _dp(mu), _dp(tg),
                                                          tg.shape[0], float(g['rate_scale']),
                                                          1 if g.get('approximate', True) else 0, gap))
        else:
            ev, v, vi, Pi = _f64(g['eigenvals']), _f64(g['v']), _f64(g['v_inv']), _f64(g['Pi'])
            if ev.shape[0] != self.n_states:
                raise ValueError('GTR has %d states, engine was created for %d' % (ev.shape[0], self.n_states))
            _lib.check(self.lib.ttb_set_gtr(self.h, _dp(ev), _dp(v), _dp(vi), _dp(Pi), float(g['mu']), gap))

    def set_branch_lengths(self, t):
        t = _f64(t)
        if t.shape[0] != self.n_nodes:
            raise ValueError('t must have one entry per node')
        _lib.check(self.lib.ttb_set_branch_lengths(self.h, _dp(t)))

    def set_branch_masks(self, masks, node_mask):
        """masks [n_masks, n_patterns] of 0/1 (or None to remove them), node_mask[n_nodes] = mask row or -1."""
        if masks is None or len(masks) == 0:
            _lib.check(self.lib.ttb_set_branch_masks(self.h, 0, None, None))
            return
        masks = np.ascontiguousarray(masks, dtype=np.uint8)
        node_mask = _i32(node_mask)
        if masks.shape[1] != self.n_patterns or node_mask.shape[0] != self.n_nodes:
            raise ValueError('masks must be [n_masks, n_patterns] and node_mask [n_nodes]')
        _lib.check(self.lib.ttb_set_branch_masks(self.h, masks.shape[0], _up(masks), _ip(node_mask)))

    # -- the pass -------------------------------------------------------------
    def marginal(self, reconstruct_tips=False, lh_only=False, keep_prev=False):
        """keep_prev: keep the states this pass overwrites (sample_states counts N_diff against them)."""
        flags = (RECONSTRUCT_TIPS if reconstruct_tips else 0) | (LH_ONLY if lh_only else 0) | (KEEP_PREV_STATES if keep_prev else 0)
        _lib.check(self.lib.ttb_marginal(self.h, flags))

    def sample_states(self, nodes, uniforms):
        """ttb_sample_states: draw the states of `nodes` from their marginal profiles with the caller's uniforms
        [len(nodes), n_patterns]; returns (changed states of internal nodes, of tips) w.r.t. the previous pass."""
        nodes = _i32(np.atleast_1d(nodes))
        u = _f64(uniforms)
        if u.shape != (nodes.shape[0], self.n_patterns):
            raise ValueError('uniforms must be [len(nodes), n_patterns]')
        nd, ndt = ctypes.c_int64(), ctypes.c_int64()
        _lib.check(self.lib.ttb_sample_states(self.h, nodes.shape[0], _ip(nodes), _dp(u), ctypes.byref(nd), ctypes.byref(ndt)))
        return nd.value, ndt.value

    def joint(self, reconstruct_tips=False, trace=True):
        """Joint (max-product) reconstruction; results() returns (sequence_joint_LH, N_diff).
        trace=False stops after the root so that the caller can sample the root (joint_retrace)."""
        _lib.check(self.lib.ttb_joint(self.h, (RECONSTRUCT_TIPS if reconstruct_tips else 0) | (0 if trace else 4)))

    def joint_retrace(self, root_idx, reconstruct_tips=False):
        root_idx = np.ascontiguousarray(root_idx, dtype=np.uint8)
        if root_idx.shape[0] != self.n_patterns:
            raise ValueError('root_idx must have one state per pattern')
        _lib.check(self.lib.ttb_joint_retrace(self.h, _up(root_idx), RECONSTRUCT_TIPS if reconstruct_tips else 0))

    def results(self):
        tot = ctypes.c_double()
        nd = ctypes.c_int64()
        _lib.check(self.lib.ttb_results(self.h, ctypes.byref(tot), ctypes.byref(nd)))
        return tot.value, nd.value

    def results_tips(self):
        """The share of N_diff that comes from terminal nodes."""
        nd = ctypes.c_int64()
        _lib.check(self.lib.ttb_results_tips(self.h, ctypes.byref(nd)))
        return nd.value

    def results_device_ptr(self):
        p = ctypes.c_void_p()
        _lib.check(self.lib.ttb_results_device_ptr(self.h, ctypes.byref(p)))
        return p.value

    def sync(self):
        _lib.check(self.lib.ttb_sync(self.h))

    # -- outputs --------------------------------------------------------------
    def site_lh(self):
        out = np.empty(self.n_patterns, dtype=np.float64)
        _lib.check(self.lib.ttb_fetch_site_lh(self.h, _dp(out)))
        return out

    def node_array(self, node, which):
        out = np.empty((self.n_patterns, self.n_states), dtype=np.float64)
        _lib.check(self.lib.ttb_fetch_node(self.h, int(node), int(which), _dp(out)))
        return out

    def seq_idx(self, nodes):
        nodes = _i32(np.atleast_1d(nodes))
        out = np.empty((nodes.shape[0], self.n_patterns), dtype=np.uint8)
        _lib.check(self.lib.ttb_fetch_seq_idx(self.h, nodes.shape[0], _ip(nodes), _up(out)))
        return out

    def all_seq_idx(self, out=None):
        """State indices of all internal nodes, [n_internal, n_patterns] in node order."""
        n_int = int((self.tip_row < 0).sum())
        if out is None:
            out = np.empty((n_int, self.n_patterns), dtype=np.uint8)
        _lib.check(self.lib.ttb_fetch_all_seq_idx(self.h, _up(out)))
        return out

    def mutations(self, max_n=None):
        """Sparse form of all reconstructed sequences: (root_idx[L'], node[], pos[], state[]) with one entry per
        (internal node, pattern) whose state differs from the parent's, sorted by (node, pos)."""
        if max_n is None:
            max_n = max(1 << 16, 8 * self.n_nodes)
        root = np.empty(self.n_patterns, dtype=np.uint8)
        while True:
            node = np.empty(max_n, dtype=np.int32); pos = np.empty(max_n, dtype=np.int32); st = np.empty(max_n, dtype=np.uint8)
            n = ctypes.c_int64()
            _lib.check(self.lib.ttb_fetch_mutations(self.h, _up(root), max_n, _ip(node), _ip(pos), _up(st), ctypes.byref(n)))
            if n.value <= max_n:
                break
            max_n = int(n.value)
        m = int(n.value)
        return root, node[:m], pos[:m], st[:m]          # the library returns them ordered by (node, pos)

    def enqueue_site_lh(self, out):
        """Stream-ordered D2H of tree.sequence_LH into `out` (pinned); valid after sync()."""
        _lib.check(self.lib.ttb_enqueue_fetch_site_lh(self.h, _dp(out)))

    def enqueue_all_seq_idx(self, out):
        """Stream-ordered D2H of all internal state indices into `out` (pinned); valid after sync()."""
        _lib.check(self.lib.ttb_enqueue_fetch_all_seq_idx(self.h, _up(out)))

    def profile_marginal(self, reconstruct_tips=False, lh_only=False):
        """Un-graphed pass with per-phase CUDA-event times: dict phase -> (ms, launches)."""
        flags = (RECONSTRUCT_TIPS if reconstruct_tips else 0) | (LH_ONLY if lh_only else 0)
        ms = np.zeros(4, dtype=np.float64)
        nl = np.zeros(4, dtype=np.int32)
        _lib.check(self.lib.ttb_profile_marginal(self.h, flags, _dp(ms), _ip(nl)))
        return {k: (float(ms[i]), int(nl[i])) for i, k in enumerate(('expqt', 'postorder', 'root', 'preorder'))}

    def branch_objective(self, nodes, t, kinds=None):
        nodes, t = _i32(nodes), _f64(t)
        out = np.empty(nodes.shape[0], dtype=np.float64)
        kp = None
        if kinds is not None:
            kinds = _i32(kinds)
            kp = _ip(kinds)
        _lib.check(self.lib.ttb_branch_objective(self.h, nodes.shape[0], _ip(nodes), kp, _dp(t), _dp(out)))
        return out

    def brent_minimize(self, nodes, kinds, xa, xb, xc, tol, maxiter=500, allreduce=None, check_every=4):
        """Lock-step Brent over s = sqrt(t) for all listed branches with the state machine on the device
        (ttb_brent_*): per iteration one objective launch + one state-update launch, no host round trip; the host only
        reads the number of unfinished branches every `check_every` iterations.  allreduce(dev_ptr, n): sums the n
        objective values in place over the pattern shards (NCCL), or None.  Returns the dict of brent_lockstep."""
        from .brent import BracketError
        nodes = _i32(nodes)
        n = nodes.shape[0]
        kp = None
        if kinds is not None:
            kinds = _i32(kinds)
            kp = _ip(kinds)
        xa, xb, xc = _f64(xa), _f64(xb), _f64(xc)
        swap = xa > xc
        if swap.any():
            xa, xc = np.where(swap, xc, xa), np.where(swap, xa, xc)
        try:
            _lib.check(self.lib.ttb_brent_begin(self.h, n, _ip(nodes), kp, _dp(xa), _dp(xb), _dp(xc), float(tol), int(maxiter)))
            ptr, cnt = ctypes.c_void_p(), ctypes.c_int32()
            if allreduce is not None:
                _lib.check(self.lib.ttb_brent_f_device_ptr(self.h, ctypes.byref(ptr), ctypes.byref(cnt)))
            it = 0
            while True:
                _lib.check(self.lib.ttb_brent_eval(self.h))
                if allreduce is not None:
                    allreduce(ptr.value, n)
                it += 1
                sync = it >= 3 and (it - 3) % check_every == 0
                na = ctypes.c_int32(-1)
                _lib.check(self.lib.ttb_brent_update(self.h, 1 if sync else 0, ctypes.byref(na)))
                if (sync and na.value == 0) or it > maxiter + check_every + 3:
                    break
        except _lib.TTBError as e:
            if 'Bracketing values' in str(e):
                raise BracketError(str(e).split(': ', 1)[-1])
            raise
        x = np.empty(n); fun = np.empty(n)
        nit = np.empty(n, dtype=np.int32); nfev = np.empty(n, dtype=np.int32)
        _lib.check(self.lib.ttb_brent_result(self.h, _dp(x), _dp(fun), _ip(nit), _ip(nfev)))
        success = (nit < maxiter) & ~(np.isnan(x) | np.isnan(fun))
        return dict(x=x, fun=fun, nit=nit.astype(np.int64), nfev=nfev.astype(np.int64), success=success)

    def branch_hamming(self, nodes, kinds=None):
        nodes = _i32(nodes)
        num = np.empty(nodes.shape[0], dtype=np.float64)
        den = ctypes.c_double()
        kp = None
        if kinds is not None:
            kinds = _i32(kinds)
            kp = _ip(kinds)
        _lib.check(self.lib.ttb_branch_hamming(self.h, nodes.shape[0], _ip(nodes), kp, _dp(num), ctypes.byref(den)))
        return num, den.value

    def mutation_counts(self):
        q = self.n_states
        n_ij = np.empty((q, q), dtype=np.float64)
        T_i = np.empty(q, dtype=np.float64)
        _lib.check(self.lib.ttb_mutation_counts(self.h, _dp(n_ij), _dp(T_i)))
        return n_ij, T_i

    def mutation_counts_per_site(self):
        """(n_ija[q, q, L'], T_ia[q, L']): the statistics of mutation_counts() before the sum over patterns."""
        q = self.n_states
        n_ija = np.empty((q, q, self.n_patterns), dtype=np.float64)
        T_ia = np.empty((q, self.n_patterns), dtype=np.float64)
        _lib.check(self.lib.ttb_mutation_counts_per_site(self.h, _dp(n_ija), _dp(T_ia)))
        return n_ija, T_ia

    def branch_state_pairs(self, nodes, tip_states=False):
        """(counts[n, q, W], first[n, q, W]) of ttb_branch_state_pairs: parent-state x child-state (or tip code)
        multiplicity sums and the first pattern showing each pair; W = q with tip_states else max(q, n_codes)."""
        nodes = _i32(np.atleast_1d(nodes))
        q = self.n_states
        W = q if tip_states else max(q, self.n_codes)
        counts = np.empty((nodes.shape[0], q, W), dtype=np.float64)
        first = np.empty((nodes.shape[0], q, W), dtype=np.int32)
        _lib.check(self.lib.ttb_branch_state_pairs(self.h, nodes.shape[0], _ip(nodes), 1 if tip_states else 0, W, _dp(counts), _ip(first)))
        return counts, first

    def seqgen(self, seed, state2code, root_idx=None, uniforms=None, return_states=True):
        """ttb_seqgen: evolve sequences down the tree; the tips become the engine's alignment.  Returns the
        state indices of every node [n_nodes, n_patterns] (or None)."""
        state2code = np.ascontiguousarray(state2code, dtype=np.uint8)
        if state2code.shape[0] != self.n_states:
            raise ValueError('state2code needs one entry per state')
        r = None if root_idx is None else np.ascontiguousarray(root_idx, dtype=np.uint8)
        u = None if uniforms is None else _f64(uniforms)
        if u is not None and u.shape != (self.n_nodes, self.n_patterns):
            raise ValueError('uniforms must be [n_nodes, n_patterns]')
        if r is not None and r.shape[0] != self.n_patterns:
            raise ValueError('root_idx must have n_patterns entries')
        out = np.empty((self.n_nodes, self.n_patterns), dtype=np.uint8) if return_states else None
        _lib.check(self.lib.ttb_seqgen(self.h, int(seed) & 0xFFFFFFFFFFFFFFFF, _up(r) if r is not None else None,
                                       _dp(u) if u is not None else None, _up(state2code), _up(out) if out is not None else None))
        return out

    def device_bytes(self):
        b = ctypes.c_int64()
        _lib.check(self.lib.ttb_device_bytes(self.h, ctypes.byref(b)))
        return b.value

    def launch_count(self):
        b = ctypes.c_int64()
        _lib.check(self.lib.ttb_launch_count(self.h, ctypes.byref(b)))
        return b.value
