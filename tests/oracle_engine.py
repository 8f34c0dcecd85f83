"""TEST-ONLY stand-in for treetime_b200.engine.Engine backed by the CPU oracle.

Lets the host logic (TreeAnc mirror, lock-step Brent, pattern sharding, the
TreeTime drop-in) be exercised without a GPU.  It is injected explicitly through
`TreeAnc(..., engine_factory=oracle_engine.factory)`; nothing in treetime_b200/
knows about it and the product path never falls back to it."""
import numpy as np
import flat_numpy as O


class OracleEngine(object):
    def __init__(self, n_states, device=0):
        self.n_states = n_states
        self.flat = {}
        self.g = None
        self.res = None
        self.prev_idx = None
        self._prev_before = None
        self.launches = 0
        self.masks = None

    def set_branch_masks(self, masks, node_mask):
        if masks is None or len(masks) == 0:
            self.masks = None
        else:
            M = np.asarray(masks, dtype=float)
            self.masks = {n: M[k] for n, k in enumerate(node_mask) if k >= 0}

    def _mult(self, n, kind=0):
        """data.multiplicity(mask=node.mask); merged root branch: mask(n1) * mask(n2) if both exist (treeanc.py:1326-1333)."""
        m = self.flat['multiplicity']
        if self.masks is None:
            return m
        if kind == 1:
            c0 = self.flat['child_ptr'][0]
            n1, n2 = int(self.flat['child_idx'][c0]), int(self.flat['child_idx'][c0 + 1])
            if n1 in self.masks and n2 in self.masks:
                return m * (self.masks[n1] * self.masks[n2])
            return m
        return m * self.masks[n] if n in self.masks else m

    def set_stream(self, s):
        pass

    def set_tree(self, parent, child_ptr, child_idx, tip_row):
        self.flat.update(parent=np.array(parent), child_ptr=np.array(child_ptr), child_idx=np.array(child_idx),
                         tip_row=np.array(tip_row))
        self.n_nodes = len(parent)
        self.tip_row = np.array(tip_row)
        self.res = None
        self.prev_idx = None
        self.masks = None

    def set_patterns(self, tip_codes, code_profiles, multiplicity, validate=True):
        self.flat.update(tip_codes=np.array(tip_codes), code_profiles=np.array(code_profiles, dtype=float),
                         multiplicity=np.array(multiplicity, dtype=float))
        self.n_patterns = tip_codes.shape[1]
        self.res = None
        self.prev_idx = None
        self.masks = None

    def set_patterns_sparse(self, ref_codes, entry_row, entry_pos, entry_code, code_profiles, multiplicity):
        n_tips = int((self.tip_row >= 0).sum())
        codes = np.repeat(np.asarray(ref_codes, dtype=np.uint8)[None, :], n_tips, axis=0)
        codes[np.asarray(entry_row), np.asarray(entry_pos)] = entry_code
        self.set_patterns(codes, code_profiles, multiplicity)

    def mutations(self, max_n=None):
        idx = self.all_seq_idx()
        internal = np.nonzero(self.tip_row < 0)[0]
        slot = {int(n): k for k, n in enumerate(internal)}
        node, pos, st = [], [], []
        for k, n in enumerate(internal):
            if n == 0:
                continue
            d = np.nonzero(idx[k] != idx[slot[int(self.flat['parent'][n])]])[0]
            node.append(np.full(d.shape[0], n, dtype=np.int32)); pos.append(d.astype(np.int32)); st.append(idx[k][d])
        cat = lambda x, dt: np.concatenate(x).astype(dt) if x else np.zeros(0, dtype=dt)  # noqa: E731
        return idx[0].copy(), cat(node, np.int32), cat(pos, np.int32), cat(st, np.uint8)

    def alignment_stats(self, aln, fill_overhangs=False, gap='-', fill='N', ambiguous='N'):
        A = np.array(aln, dtype=np.uint8)
        if fill_overhangs:
            ng = A != ord(gap); anyng = ng.any(axis=1)
            first = np.where(anyng, ng.argmax(axis=1), A.shape[1]); last = np.where(anyng, A.shape[1] - 1 - ng[:, ::-1].argmax(axis=1), -1)
            pos = np.arange(A.shape[1])[None, :]
            A[(pos < first[:, None]) | (pos > last[:, None])] = ord(fill)
        self._aln = A
        a = ord(ambiguous) if ambiguous is not None else 256
        isa = A == a
        return np.where(isa, 255, A).min(axis=0).astype(np.uint8), np.where(isa, 0, A).max(axis=0).astype(np.uint8), isa.all(axis=0)

    def set_patterns_from_alignment(self, first_pos, const_letter, tip_seq_row, lut, missing_code, code_profiles, multiplicity):
        C = self._aln[:, first_pos].copy()
        cc = const_letter != 0
        C[:, cc] = const_letter[cc][None, :]
        codes = np.full((len(tip_seq_row), len(first_pos)), missing_code, dtype=np.uint8)
        have = np.asarray(tip_seq_row) >= 0
        codes[have] = lut[C[np.asarray(tip_seq_row)[have]]]
        assert (codes != 255).all()
        self.set_patterns(codes, code_profiles, multiplicity)

    def set_gtr(self, g):
        self.g = dict(g)

    def set_branch_lengths(self, t):
        self.flat['t'] = np.array(t, dtype=float)

    def sample_states(self, nodes, uniforms):
        """Engine.sample_states: redraw the states of `nodes` from their profiles, count changes vs the previous pass."""
        nd = ndt = 0
        for k, n in enumerate(np.atleast_1d(nodes)):
            n = int(n)
            idx = O.sample_idx(self.res.profile[n], np.asarray(uniforms)[k])
            prev = self._prev_before[n] if self._prev_before is not None else None
            d = int((idx != prev).sum()) if prev is not None else self.n_patterns
            if self.tip_row[n] >= 0:
                ndt += d
            else:
                nd += d
            self.res.seq_idx[n] = idx
            self.prev_idx[n] = idx
        return nd, ndt

    def marginal(self, reconstruct_tips=False, lh_only=False, keep_prev=False):
        self.launches += 1
        if lh_only:
            r = O.sequence_LH_only(self.flat, self.g, masks=self.masks)
            self._tot, self._nd = r.total_LH, 0
            self._site = r.sequence_LH
            return
        self.t_pass = self.flat['t'].copy()
        self.g_pass = dict(self.g)
        self._prev_before = list(self.prev_idx) if self.prev_idx is not None else None
        r = O.marginal(self.flat, self.g, reconstruct_tip_states=reconstruct_tips, prev_seq_idx=self.prev_idx, masks=self.masks)
        if self.prev_idx is None:
            pass
        self.res = r
        self.prev_idx = list(r.seq_idx)
        self._tot, self._nd, self._site = r.total_LH, r.N_diff, r.sequence_LH

    def joint(self, reconstruct_tips=False, trace=True):
        self.launches += 1
        self._prev_before = list(self.prev_idx) if self.prev_idx is not None else None
        r = O.joint(self.flat, self.g, reconstruct_tip_states=reconstruct_tips, prev_seq_idx=self.prev_idx)
        self.jres = r
        self.res = None
        self._tot, self._site = r.total_LH, r.sequence_LH
        if trace:
            self._nd = r.N_diff
            self.prev_idx = list(r.seq_idx)
            self.seqs = r.seq_idx

    def joint_retrace(self, root_idx, reconstruct_tips=False):
        r = self.jres
        n_nodes = self.n_nodes
        seq = [None] * n_nodes
        seq[0] = np.asarray(root_idx).astype(int)
        order = [n for n in range(1, n_nodes) if self.tip_row[n] < 0]
        if reconstruct_tips:
            order += [n for n in range(1, n_nodes) if self.tip_row[n] >= 0]
        nd = 0
        L = self.n_patterns
        for n in order:
            seq[n] = np.choose(seq[self.flat['parent'][n]], r.joint_Cx[n].T)
            if self.prev_idx is not None and self.prev_idx[n] is not None:
                nd += int((seq[n] != self.prev_idx[n]).sum())
            else:
                nd += L
        self._site = np.choose(seq[0], r.joint_Lx[0].T)
        self._tot = (self._site * self.flat['multiplicity']).sum()
        self._nd = nd
        self.prev_idx = list(seq)
        self.seqs = seq

    def branch_state_pairs(self, nodes, tip_states=False):
        src = self.res.seq_idx if self.res is not None else self.seqs
        C, F = O.branch_pair_tables(self.flat, src, tip_states=tip_states)
        k = np.atleast_1d(nodes).astype(int) - 1
        return C[k], F[k]

    def seqgen(self, seed, state2code, root_idx=None, uniforms=None, return_states=True):
        L = self.n_patterns
        if uniforms is None:        # the device uses Philox; any stream of uniforms has the same distribution
            uniforms = np.random.default_rng(seed).random((self.n_nodes, L))
        st = O.seqgen(self.flat, O.make_gtr(self.g), np.asarray(uniforms), root_idx=root_idx)
        tips = self.tip_row >= 0
        self.flat['tip_codes'] = np.asarray(state2code, dtype=np.uint8)[st[tips]][np.argsort(self.tip_row[tips])]
        self.res = None
        self.prev_idx = None
        return st if return_states else None

    def results_tips(self):
        # share of the last N_diff that came from tips (recomputed: the oracle returns only the total)
        src = self.res.seq_idx if self.res is not None else self.seqs
        nd = 0
        for n in range(self.n_nodes):
            if self.tip_row[n] >= 0 and src[n] is not None:
                prev = self._prev_before[n] if self._prev_before is not None else None
                nd += int((src[n] != prev).sum()) if prev is not None else self.n_patterns
        return nd

    def results(self):
        return self._tot, self._nd

    def sync(self):
        pass

    def site_lh(self):
        return self._site.copy()

    def node_array(self, node, which):
        if which == 3:
            return np.array(self.jres.joint_Lx[0])
        r = self.res
        return np.array((r.subtree_LH, r.outgroup_LH, r.profile)[which][node])

    def seq_idx(self, nodes):
        src = self.res.seq_idx if self.res is not None else self.seqs
        return np.array([src[int(n)] for n in np.atleast_1d(nodes)], dtype=np.uint8)

    def all_seq_idx(self, out=None):
        src = self.res.seq_idx if self.res is not None else self.seqs
        rows = [src[n] for n in range(self.n_nodes) if self.tip_row[n] < 0]
        return np.array(rows, dtype=np.uint8)

    def _pair(self, n, kind):
        if kind == 1:
            return O.root_branch_profiles(self._pass_flat(), self.g_pass, self.res)
        return self.res.outgroup_LH[n], self.res.subtree_LH[n]

    def _pass_flat(self):
        f = dict(self.flat)
        f['t'] = self.t_pass
        return f

    def branch_objective(self, nodes, t, kinds=None):
        G = O.make_gtr(self.g_pass)
        kinds = np.zeros(len(nodes), dtype=int) if kinds is None else kinds
        return np.array([G.prob_t_profiles(self._pair(int(n), int(k)), self._mult(int(n), int(k)), float(tt), return_log=True)
                         for n, k, tt in zip(nodes, kinds, t)])

    def branch_hamming(self, nodes, kinds=None):
        kinds = np.zeros(len(nodes), dtype=int) if kinds is None else kinds
        m = self.flat['multiplicity']
        num = []
        for n, k in zip(nodes, kinds):
            pp, pc = self._pair(int(n), int(k))
            num.append(np.sum(self._mult(int(n), int(k)) * np.sum(pp * pc, axis=1)))
        return np.array(num), m.sum()

    def mutation_counts(self):
        n_ija, T_ia = O.mutation_counts(self._pass_flat(), self.g_pass, self.res, masks=self.masks)
        return n_ija.sum(axis=-1), T_ia.sum(axis=-1)

    def mutation_counts_per_site(self):
        return O.mutation_counts(self._pass_flat(), self.g_pass, self.res, masks=self.masks)

    def launch_count(self):
        return self.launches

    def device_bytes(self):
        return 0


def factory(n_states, device):
    return OracleEngine(n_states, device)
