"""CPU ORACLE (test infrastructure, NOT product code).

Flat-array numpy restatement of TreeTime's marginal ancestral reconstruction
(reference: /root/reference/treetime, v0.12.1).  Every function cites the
reference lines it restates.  It deliberately keeps the reference's per-node
numpy call structure (one `.dot`, one `np.log`, one `normalize_profile` ... per
node on (L', q) arrays) so that (a) the arithmetic -- including the order of
floating point operations inside numpy -- is the reference's, and (b) timing it
is a faithful stand-in ("port") for the reference's CPU path on machines where
the reference itself cannot be imported (the GPU box has no /root/reference).

Parity status: PINNED.  tests/test_oracle_golden.py checks this file against
 * the reference's own known-answer test (test/test_treetime.py:137-155:
   sum_patterns exp(LH) == 1) and
 * golden vectors produced by the UNMODIFIED reference in the build container
   (oracle/make_golden.py, committed under tests/golden/).
oracle/validate_against_reference.py re-runs the comparison live when
/root/reference is present.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this module.  treetime_b200/ never does.

Flat problem layout (all numpy, built by treetime_b200.flatten or by tests):
  parent[n]            int32, -1 for the root; nodes are numbered in the
                       reference's preorder (tree.find_clades()), root = 0
  child_ptr[n+1], child_idx[...]   CSR children in `node.clades` order
  tip_row[n]           row of `tip_codes` for terminal nodes, -1 for internal
  tip_codes[n_tips,L'] uint8 indices into code_profiles
  code_profiles[c,q]   0/1 ambiguity profiles (gtr.profile_map values)
  multiplicity[L']     float64 pattern weights
  t[n]                 branch length used by the GTR (already floored,
                       treeanc.py:752-760); t[root] unused
  gtr                  dict(eigenvals, v, v_inv, Pi, mu) -- or the per-site
                       variant with 'site_specific': True (see SiteSpecific below)
"""
import numpy as np

TINY_NUMBER = 1e-12        # treetime/config.py:4
SUPERTINY_NUMBER = 1e-24   # treetime/config.py:5
MIN_BRANCH_LENGTH = 1e-3   # treetime/config.py:7
MAX_BRANCH_LENGTH = 4.0    # treetime/config.py:34
BIG_NUMBER = 1e10          # treetime/config.py:3


# ----------------------------------------------------------------------------
# L1 primitives
# ----------------------------------------------------------------------------
def normalize_profile(in_profile, log=False, return_offset=True):
    """seq_utils.py:279-307."""
    if log:
        tmp_prefactor = in_profile.max(axis=1)
        tmp_prof = np.exp(in_profile.T - tmp_prefactor).T
    else:
        tmp_prefactor = 0.0
        tmp_prof = in_profile
    norm_vector = tmp_prof.sum(axis=1)
    return (
        np.einsum('ai,a->ai', tmp_prof, 1.0 / norm_vector),
        (np.log(norm_vector) + tmp_prefactor) if return_offset else None,
    )


def eig_single_site(W, p):
    """gtr.py:612-629 (symmetrised eigendecomposition of the rate matrix)."""
    assert np.abs(np.diag(W).sum()) < 1e-10
    tmpp = np.sqrt(p)
    symQ = W * np.outer(tmpp, tmpp)
    np.fill_diagonal(symQ, -np.sum(W * p, axis=1))
    eigvals, eigvecs = np.linalg.eigh(symQ)
    tmp_v = eigvecs.T * tmpp
    one_norm = np.sum(np.abs(tmp_v), axis=1)
    return eigvals, tmp_v.T / one_norm, (eigvecs * one_norm).T / tmpp


# ----------------------------------------------------------------------------
# L2: single-site GTR arithmetic
# ----------------------------------------------------------------------------
class FlatGTR(object):
    """gtr.py arithmetic on (eigenvals, v, v_inv, Pi, mu)."""
    site_specific = False

    def __init__(self, eigenvals, v, v_inv, Pi, mu, gap_index=None):
        self.eigenvals = np.asarray(eigenvals, dtype=float)
        self.v = np.asarray(v, dtype=float)
        self.v_inv = np.asarray(v_inv, dtype=float)
        self.Pi = np.asarray(Pi, dtype=float)
        self.mu = float(mu)
        self.gap_index = gap_index
        self.n_states = self.Pi.shape[0]

    def _exp_lt(self, t):
        """gtr.py:1027-1049."""
        log_val = self.mu * t * self.eigenvals
        if any(i > 10 for i in log_val):
            raise ValueError('Error in computing exp(Q * t): Q has positive eigenvalues or the branch length t is too large.')
        return np.exp(log_val)

    def expQt(self, t):
        """gtr.py:1051-1067. expQt[i,j] = P(child=i | parent=j)."""
        eLambdaT = np.diag(self._exp_lt(t))
        Qs = self.v.dot(eLambdaT.dot(self.v_inv))
        return np.maximum(0, Qs)

    def propagate_profile(self, profile, t, return_log=False):
        """gtr.py:965-995 (child -> parent message)."""
        Qt = self.expQt(t)
        res = profile.dot(Qt)
        return np.log(res) if return_log else res

    def evolve(self, profile, t, return_log=False):
        """gtr.py:997-1025 (parent -> child message)."""
        Qt = self.expQt(t).T
        res = profile.dot(Qt)
        return np.log(res) if return_log else res

    def prob_t_profiles(self, profile_pair, multiplicity, t, return_log=False, ignore_gaps=True):
        """gtr.py:922-963."""
        if t < 0:
            logP = -BIG_NUMBER
        else:
            Qt = self.expQt(t)
            res = np.einsum('ai,ij,aj->a', profile_pair[1], Qt, profile_pair[0])
            if ignore_gaps and (self.gap_index is not None):
                non_gap_frac = (1 - profile_pair[0][:, self.gap_index]) * (1 - profile_pair[1][:, self.gap_index])
                logP = np.sum(multiplicity * np.log(res + SUPERTINY_NUMBER) * non_gap_frac)
            else:
                logP = np.sum(multiplicity * np.log(res + SUPERTINY_NUMBER))
        return logP if return_log else np.exp(logP)

    def optimal_t_compressed(self, seq_pair, multiplicity, tol=1e-10, return_nfev=False):
        """gtr.py:816-920, profiles=True branch (scipy Brent on s = sqrt(t))."""
        from scipy.optimize import minimize_scalar

        def _neg_prob(t, seq_pair, multiplicity):
            res = -1.0 * self.prob_t_profiles(seq_pair, multiplicity, t**2, return_log=True)
            return res + np.exp(t**4 / 10000)

        hamming_distance = 1 - np.sum(multiplicity * np.sum(seq_pair[0] * seq_pair[1], axis=1)) / np.sum(multiplicity)
        opt = minimize_scalar(
            _neg_prob,
            bracket=[-np.sqrt(MAX_BRANCH_LENGTH), np.sqrt(hamming_distance), np.sqrt(MAX_BRANCH_LENGTH)],
            args=(seq_pair, multiplicity), tol=tol, method='brent',
        )
        new_len = opt['x'] ** 2
        if opt.get('success', True) != True:  # noqa: E712  (gtr.py:916-918)
            new_len = hamming_distance
        if return_nfev:
            return new_len, opt['nfev']
        return new_len


# ----------------------------------------------------------------------------
# L2: site-specific GTR arithmetic
# ----------------------------------------------------------------------------
class FlatGTRSiteSpecific(object):
    """gtr_site_specific.py arithmetic: per-site Pi (q,L), mu (L), eigen-systems
    v, v_inv (q,q,L), eigenvals (q,L); linear-in-t interpolated expQt
    (gtr_site_specific.py:331-371)."""
    site_specific = True

    def __init__(self, eigenvals, v, v_inv, Pi, mu, rate_scale, approximate=True, gap_index=None):
        self.eigenvals = np.asarray(eigenvals, dtype=float)
        self.v = np.asarray(v, dtype=float)
        self.v_inv = np.asarray(v_inv, dtype=float)
        self.Pi = np.asarray(Pi, dtype=float)
        self.mu = np.asarray(mu, dtype=float)
        self.rate_scale = float(rate_scale)
        self.approximate = approximate
        self.gap_index = gap_index
        self.n_states = self.Pi.shape[0]
        self.t_grid = self.make_t_grid(self.rate_scale)
        self._stack = None

    @staticmethod
    def make_t_grid(rate_scale):
        """gtr_site_specific.py:336-344 (61 points)."""
        return (1.0 / rate_scale) * np.concatenate((
            np.linspace(0, 0.1, 11)[:-1], np.linspace(0.1, 1, 21)[:-1],
            np.linspace(1, 5, 21)[:-1], np.linspace(5, 10, 11)))

    def _expQt(self, t):
        """gtr_site_specific.py:350-365."""
        eLambdaT = np.exp(t * self.mu * self.eigenvals)
        return np.einsum('jia,ja,kja->ika', self.v, eLambdaT, self.v_inv)

    def expQt(self, t):
        """gtr_site_specific.py:367-371; the interpolation is scipy interp1d
        kind='linear' over the stacked matrices (:345-348): slope*(x-x_lo)+y_lo."""
        if t * self.rate_scale < 10 and self.approximate:
            if self._stack is None:
                self._stack = np.stack([self._expQt(tg) for tg in self.t_grid], axis=0)
            from scipy.interpolate import interp1d
            return interp1d(self.t_grid, self._stack, axis=0, assume_sorted=True, copy=False, kind='linear')(t)
        return self._expQt(t)

    def propagate_profile(self, profile, t, return_log=False):
        """gtr_site_specific.py:376-406."""
        Qt = self.expQt(t)
        res = np.einsum('ai,ija->aj', profile, Qt)
        return np.log(np.maximum(TINY_NUMBER, res)) if return_log else np.maximum(0, res)

    def evolve(self, profile, t, return_log=False):
        """gtr_site_specific.py:408-437."""
        Qt = self.expQt(t)
        res = np.einsum('ai,jia->aj', profile, Qt)
        return np.log(res) if return_log else res

    def prob_t_profiles(self, profile_pair, multiplicity, t, return_log=False, ignore_gaps=True):
        """gtr.py:922-963, 3-d branch (:951-952)."""
        if t < 0:
            logP = -BIG_NUMBER
        else:
            Qt = self.expQt(t)
            res = np.einsum('ai,ija,aj->a', profile_pair[1], Qt, profile_pair[0])
            if ignore_gaps and (self.gap_index is not None):
                non_gap_frac = (1 - profile_pair[0][:, self.gap_index]) * (1 - profile_pair[1][:, self.gap_index])
                logP = np.sum(multiplicity * np.log(res + SUPERTINY_NUMBER) * non_gap_frac)
            else:
                logP = np.sum(multiplicity * np.log(res + SUPERTINY_NUMBER))
        return logP if return_log else np.exp(logP)

    optimal_t_compressed = FlatGTR.optimal_t_compressed


def make_gtr(g):
    """dict -> FlatGTR / FlatGTRSiteSpecific."""
    if isinstance(g, (FlatGTR, FlatGTRSiteSpecific)):
        return g
    if g.get('site_specific', False):
        return FlatGTRSiteSpecific(g['eigenvals'], g['v'], g['v_inv'], g['Pi'], g['mu'], g['rate_scale'],
                                   approximate=g.get('approximate', True), gap_index=g.get('gap_index'))
    return FlatGTR(g['eigenvals'], g['v'], g['v_inv'], g['Pi'], g['mu'], gap_index=g.get('gap_index'))


# ----------------------------------------------------------------------------
# L4: the marginal engine on flat arrays
# ----------------------------------------------------------------------------
class MarginalResult(object):
    """Per-node arrays as the reference leaves them on the clades."""
    def __init__(self, n_nodes):
        self.subtree_LH = [None] * n_nodes        # node.marginal_subtree_LH
        self.prefactor = [None] * n_nodes         # node.marginal_subtree_LH_prefactor
        self.log_Lx = [None] * n_nodes            # node.marginal_log_Lx
        self.outgroup_LH = [None] * n_nodes       # node.marginal_outgroup_LH
        self.profile = [None] * n_nodes           # node.marginal_profile
        self.seq_idx = [None] * n_nodes           # argmax state index (alphabet[idx] = node._cseq)
        self.sequence_LH = None                   # tree.sequence_LH
        self.total_LH = None                      # tree.total_sequence_LH
        self.N_diff = None


def postorder(flat, gtr, res=None, masks=None):
    """treeanc.py:840-878 (postorder_traversal_marginal)."""
    gtr = make_gtr(gtr)
    parent, cptr, cidx = flat['parent'], flat['child_ptr'], flat['child_idx']
    n_nodes = parent.shape[0]
    L = flat['multiplicity'].shape[0]
    q = gtr.n_states
    t = flat['t']
    res = res or MarginalResult(n_nodes)
    # leaves (:846-853): seq2prof = table lookup (seq_utils.py:207-229)
    for n in range(n_nodes):
        if flat['tip_row'][n] >= 0:
            if res.subtree_LH[n] is None:
                res.subtree_LH[n] = flat['code_profiles'][flat['tip_codes'][flat['tip_row'][n]]]
            res.prefactor[n] = np.zeros(L)
    # internal nodes, children before parents (:857-878). Preorder ids => descending ids.
    for n in range(n_nodes - 1, -1, -1):
        if flat['tip_row'][n] >= 0:
            continue
        tmp_log_subtree_LH = np.zeros((L, q), dtype=float)
        pref = np.zeros(L, dtype=float)
        for ch in cidx[cptr[n]:cptr[n + 1]]:
            lx = gtr.propagate_profile(res.subtree_LH[ch], t[ch], return_log=True)
            if masks is not None and masks.get(int(ch)) is not None:
                lx = (lx.T * masks[int(ch)]).T              # :867-872
            res.log_Lx[ch] = lx
            tmp_log_subtree_LH += lx
            pref += res.prefactor[ch]
        res.subtree_LH[n], offset = normalize_profile(tmp_log_subtree_LH, log=True)
        res.prefactor[n] = pref + offset
    return res


def total_LH_and_root(flat, gtr, res):
    """treeanc.py:814-838 (argmax root sequence; sampling stays host-side)."""
    gtr = make_gtr(gtr)
    L = flat['multiplicity'].shape[0]
    if len(gtr.Pi.shape) == 1:
        res.outgroup_LH[0] = np.repeat([gtr.Pi], L, axis=0)
    else:
        res.outgroup_LH[0] = np.copy(gtr.Pi.T)
    res.profile[0], pre = normalize_profile(res.outgroup_LH[0] * res.subtree_LH[0])
    res.sequence_LH = res.prefactor[0] + pre
    res.total_LH = (res.sequence_LH * flat['multiplicity']).sum()
    res.seq_idx[0] = res.profile[0].argmax(axis=1)          # seq_utils.py:271
    return res


def sample_idx(profile, u):
    """prof2seq with sample_from_prof=True (seq_utils.py:266-269): first state whose cumulative probability reaches u."""
    cumdis = profile.cumsum(axis=1).T
    return np.argmax(cumdis >= u, axis=0)


def preorder(flat, gtr, res, reconstruct_tip_states=False, prev_seq_idx=None, masks=None, uniforms=None):
    """treeanc.py:880-932 (preorder_traversal_marginal); argmax assignment, or -- uniforms = {node: u[L']}, the
    reference's rng.random(L') draws -- sampled from the profile (sample_from_profile=True)."""
    gtr = make_gtr(gtr)
    parent = flat['parent']
    n_nodes = parent.shape[0]
    L = flat['multiplicity'].shape[0]
    t = flat['t']
    N_diff = 0
    for n in range(1, n_nodes):                              # preorder ids, root skipped
        up = parent[n]
        res.outgroup_LH[n], _ = normalize_profile(
            np.log(np.maximum(TINY_NUMBER, res.profile[up])) - res.log_Lx[n], log=True, return_offset=False)
        if flat['tip_row'][n] >= 0 and not reconstruct_tip_states:
            continue
        msg = gtr.evolve(res.outgroup_LH[n], t[n], return_log=False)
        if masks is not None and masks.get(n) is not None:
            m = masks[n]
            res.profile[n], _ = normalize_profile(res.subtree_LH[n] * (m * msg.T + (1.0 - m)).T, return_offset=False)
        else:
            res.profile[n], _ = normalize_profile(res.subtree_LH[n] * msg, return_offset=False)
        if uniforms is not None:
            idx = sample_idx(res.profile[n], uniforms[n])
        else:
            idx = res.profile[n].argmax(axis=1)
        if prev_seq_idx is not None and prev_seq_idx[n] is not None:
            N_diff += int((idx != prev_seq_idx[n]).sum())
        else:
            N_diff += L
        res.seq_idx[n] = idx
    res.N_diff = N_diff
    return res


def marginal(flat, gtr, reconstruct_tip_states=False, prev_seq_idx=None, masks=None, uniforms=None):
    """treeanc.py:762-812 (_ml_anc_marginal); the root is always the argmax here (root sampling stays with the caller)."""
    res = postorder(flat, gtr, masks=masks)
    total_LH_and_root(flat, gtr, res)
    preorder(flat, gtr, res, reconstruct_tip_states=reconstruct_tip_states, prev_seq_idx=prev_seq_idx, masks=masks,
             uniforms=uniforms)
    return res


def sequence_LH_only(flat, gtr, masks=None):
    """The LH-only path of optimize_gtr_rate's cost function (treeanc.py:1685-1689)."""
    res = postorder(flat, gtr, masks=masks)
    total_LH_and_root(flat, gtr, res)
    return res


# ----------------------------------------------------------------------------
# branch-length surface (A8) and substitution statistics (A10)
# ----------------------------------------------------------------------------
def marginal_branch_profile(res, n):
    """treeanc.py:1122-1146: (pp, pc) = (outgroup_LH, subtree_LH)."""
    return res.outgroup_LH[n], res.subtree_LH[n]


def optimal_marginal_branch_length(flat, gtr, res, n, tol=1e-10, return_nfev=False):
    """treeanc.py:1272-1295."""
    gtr = make_gtr(gtr)
    pp, pc = marginal_branch_profile(res, n)
    return gtr.optimal_t_compressed((pp, pc), flat['multiplicity'], tol=tol, return_nfev=return_nfev)


def branch_objective(flat, gtr, res, n, t):
    """prob_t_profiles(return_log=True) for branch n at length t (gtr.py:922-963)."""
    gtr = make_gtr(gtr)
    pp, pc = marginal_branch_profile(res, n)
    return gtr.prob_t_profiles((pp, pc), flat['multiplicity'], t, return_log=True)


def root_branch_profiles(flat, gtr, res):
    """treeanc.py:1317-1326: merged branch across a bifurcating root."""
    cptr, cidx = flat['child_ptr'], flat['child_idx']
    n1, n2 = cidx[cptr[0]:cptr[0] + 2]
    prof_c = res.subtree_LH[n1]
    prof_p = normalize_profile(res.subtree_LH[n2] * res.outgroup_LH[0])[0]
    return prof_p, prof_c


def optimize_branch_lengths_marginal_step(flat, gtr, res, branch_length, i_iter, damping=0.75):
    """One sweep of the per-branch loop of optimize_tree_marginal
    (treeanc.py:1312-1344) on flat arrays.  `branch_length` is node.branch_length
    (unfloored); returns the updated copy.  The bifurcating-root block runs once
    per root child, exactly as in the reference (Appendix C of SURVEY.md)."""
    gtr = make_gtr(gtr)
    parent, cptr = flat['parent'], flat['child_ptr']
    bl = np.array(branch_length, dtype=float)
    tol = 1e-8 + 0.01 ** (i_iter + 1)
    n_nodes = parent.shape[0]
    root_bif = (cptr[1] - cptr[0]) == 2
    for n in range(1, n_nodes):
        if parent[n] == 0 and root_bif:
            n1, n2 = flat['child_idx'][cptr[0]:cptr[0] + 2]
            total_bl = bl[n1] + bl[n2]
            bl_ratio = bl[n1] / total_bl
            prof_p, prof_c = root_branch_profiles(flat, gtr, res)
            new_bl = gtr.optimal_t_compressed((prof_p, prof_c), flat['multiplicity'], tol=tol)
            update_val = new_bl * (1 - damping ** (i_iter + 1)) + total_bl * damping ** (i_iter + 1)
            bl[n1] = update_val * bl_ratio
            bl[n2] = update_val * (1 - bl_ratio)
        else:
            new_val = optimal_marginal_branch_length(flat, gtr, res, n, tol=tol)
            bl[n] = new_val * (1 - damping ** (i_iter + 1)) + bl[n] * damping ** (i_iter + 1)
    return bl


def branch_mutation_matrix(flat, gtr, res, n):
    """treeanc.py:1085-1120 (compressed)."""
    gtr = make_gtr(gtr)
    pp, pc = marginal_branch_profile(res, n)
    expQt = gtr.expQt(flat['t'][n]) + SUPERTINY_NUMBER
    if len(expQt.shape) == 3:
        stack = np.einsum('ai,aj,ija->aij', pc, pp, expQt)
    else:
        stack = np.einsum('ai,aj,ij->aij', pc, pp, expQt)
    normalizer = stack.sum(axis=2).sum(axis=1)
    return np.einsum('aij,a->aij', stack, 1.0 / normalizer)


def mutation_counts(flat, gtr, res, masks=None):
    """treeanc.py:1556-1572: n_ija (q,q,L') and T_ia (q,L') accumulated over branches; masks = {node: mask[L']}
    as in data.multiplicity(mask=c.mask)."""
    gtr = make_gtr(gtr)
    q = gtr.n_states
    L = flat['multiplicity'].shape[0]
    n_ija = np.zeros((q, q, L))
    T_ia = np.zeros((q, L))
    m = flat['multiplicity']
    # reference loop: for node in get_nonterminals() (preorder): for c in node
    for node in range(flat['parent'].shape[0]):
        for c in flat['child_idx'][flat['child_ptr'][node]:flat['child_ptr'][node + 1]]:
            mut_stack = np.transpose(branch_mutation_matrix(flat, gtr, res, c), (1, 2, 0))
            mc = m if masks is None or masks.get(int(c)) is None else m * masks[int(c)]
            T_ia += 0.5 * flat['t'][c] * mut_stack.sum(axis=0) * mc
            T_ia += 0.5 * flat['t'][c] * mut_stack.sum(axis=1) * mc
            n_ija += mut_stack * mc
    return n_ija, T_ia


# ----------------------------------------------------------------------------
# N2: joint (max-product) ML reconstruction
# ----------------------------------------------------------------------------
class JointResult(object):
    def __init__(self, n_nodes):
        self.joint_Lx = [None] * n_nodes
        self.joint_Cx = [None] * n_nodes
        self.seq_idx = [None] * n_nodes
        self.sequence_LH = None
        self.total_LH = None      # tree.sequence_joint_LH
        self.N_diff = None


def joint(flat, gtr, reconstruct_tip_states=False, prev_seq_idx=None):
    """treeanc.py:934-1080 (_ml_anc_joint), argmax root (no sampling), no masks."""
    gtr = make_gtr(gtr)
    parent, cptr, cidx = flat['parent'], flat['child_ptr'], flat['child_idx']
    n_nodes = parent.shape[0]
    L = flat['multiplicity'].shape[0]
    q = gtr.n_states
    t = flat['t']
    res = JointResult(n_nodes)
    # postorder: children before parents = descending preorder ids (:957-1000)
    for n in range(n_nodes - 1, 0, -1):
        log_transitions = np.log(np.maximum(TINY_NUMBER, gtr.expQt(t[n])))
        if flat['tip_row'][n] >= 0:
            tmp_prof = flat['code_profiles'][flat['tip_codes'][flat['tip_row'][n]]]
            msg_from_children = np.log(np.maximum(tmp_prof, TINY_NUMBER))
            msg_from_children[np.isnan(msg_from_children) | np.isinf(msg_from_children)] = -BIG_NUMBER
        else:
            msg_from_children = np.sum(np.stack([res.joint_Lx[c] for c in cidx[cptr[n]:cptr[n + 1]]], axis=0), axis=0)
        Lx = np.zeros((L, q))
        Cx = np.zeros((L, q), dtype=np.uint16)
        for char_i in range(q):
            msg_to_parent = log_transitions[:, char_i].T + msg_from_children
            Cx[:, char_i] = msg_to_parent.argmax(axis=1)
            Lx[:, char_i] = msg_to_parent.max(axis=1)
        res.joint_Lx[n], res.joint_Cx[n] = Lx, Cx
    # root (:1003-1023)
    msg_from_children = np.sum(np.stack([res.joint_Lx[c] for c in cidx[cptr[0]:cptr[1]]], axis=0), axis=0)
    res.joint_Lx[0] = msg_from_children + np.log(gtr.Pi).T
    normalized_profile = (res.joint_Lx[0].T - res.joint_Lx[0].max(axis=1)).T
    prof, _ = normalize_profile(np.exp(normalized_profile), return_offset=False)     # prof2seq(normalize=True)
    idxs = prof.argmax(axis=1)
    res.sequence_LH = np.choose(idxs, res.joint_Lx[0].T)
    res.total_LH = (res.sequence_LH * flat['multiplicity']).sum()
    res.seq_idx[0] = idxs
    # preorder back-trace (:1029-1048): internal nodes in preorder, then the tips
    order = [n for n in range(1, n_nodes) if flat['tip_row'][n] < 0]
    if reconstruct_tip_states:
        order += [n for n in range(1, n_nodes) if flat['tip_row'][n] >= 0]
    N_diff = 0
    for n in order:
        res.seq_idx[n] = np.choose(res.seq_idx[parent[n]], res.joint_Cx[n].T)
        if prev_seq_idx is not None and prev_seq_idx[n] is not None:
            N_diff += int((res.seq_idx[n] != prev_seq_idx[n]).sum())
        else:
            N_diff += L
    res.N_diff = N_diff
    return res


# ---------------------------------------------------------------------------------------
# N2: sufficient statistics of the joint branch-length optimisation.
# Reference: TreeAnc.add_branch_state (treeanc.py:1148-1163) -> GTR.state_pair (gtr.py:631-705):
# per branch, the multiplicity-weighted number of patterns for every (parent character, child
# character) pair.  Restated on indices: parent = reconstructed state index, child = reconstructed
# state index (internal nodes, and tips with tip_states) or the tip's alignment code.
# ---------------------------------------------------------------------------------------
PAIR_NONE = 0x7fffffff


def branch_pair_tables(flat, seq_idx, tip_states=False):
    """C[n_nodes-1, q, W] multiplicity sums and F[...] index of the first pattern showing each pair
    (PAIR_NONE if none), W = q with tip_states else max(q, n_codes); row k belongs to node k+1."""
    parent, tip_row = flat['parent'], flat['tip_row']
    mult = np.asarray(flat['multiplicity'], dtype=float)
    q = flat['code_profiles'].shape[1]
    W = q if tip_states else max(q, flat['code_profiles'].shape[0])
    n_nodes = parent.shape[0]
    L = mult.shape[0]
    C = np.zeros((n_nodes - 1, q, W))
    F = np.full((n_nodes - 1, q, W), PAIR_NONE, dtype=np.int32)
    pos = np.arange(L)
    for n in range(1, n_nodes):
        p = np.asarray(seq_idx[parent[n]]).astype(int)
        if tip_row[n] >= 0 and not tip_states:
            c = flat['tip_codes'][tip_row[n]].astype(int)
        else:
            c = np.asarray(seq_idx[n]).astype(int)
        np.add.at(C[n - 1], (p, c), mult)
        np.minimum.at(F[n - 1], (p, c), pos)
    return C, F


def state_pair(alphabet, gap_index, seq_p, seq_ch, pattern_multiplicity, ignore_gaps=False):
    """GTR.state_pair (gtr.py:631-705) on character arrays -- the quirks included: alphabets of < 10
    letters count only positions where both characters are letters of the alphabet and list the pairs
    in alphabet order; larger alphabets map every other character to index 1 and list the pairs in
    order of first occurrence."""
    alphabet = [str(a) for a in alphabet]
    out = []
    if len(alphabet) < 10:
        bp = [seq_p == a for a in alphabet]
        bc = [seq_ch == a for a in alphabet]
        for n1 in range(len(alphabet)):
            if gap_index is None or not ignore_gaps or n1 != gap_index:
                for n2 in range(len(alphabet)):
                    if gap_index is None or not ignore_gaps or n2 != gap_index:
                        count = ((bp[n1] & bc[n2]) * pattern_multiplicity).sum()
                        if count:
                            out.append(((n1, n2), count))
    else:
        num = []
        for seq in (seq_p, seq_ch):
            tmp = np.ones(seq.shape[0], dtype=int)
            for ni, a in enumerate(alphabet):
                tmp[seq == a] = ni
            num.append(tmp)
        acc = {}
        for i in range(seq_p.shape[0]):
            if (not ignore_gaps) or (gap_index != num[0][i] and gap_index != num[1][i]):
                key = (num[0][i], num[1][i])
                acc[key] = acc.get(key, 0) + pattern_multiplicity[i]
        out = list(acc.items())
    return (np.array([x[0] for x in out], dtype=int).reshape(-1, 2) if out else np.zeros((0, 2), dtype=int),
            np.array([x[1] for x in out], dtype=int))


def prob_t_compressed(gtr, seq_pair, multiplicity, t):
    """GTR.prob_t_compressed(return_log=True), gtr.py:710-745."""
    if t < 0:
        return -BIG_NUMBER
    logQt = np.log(np.maximum(gtr.expQt(t), SUPERTINY_NUMBER))
    return np.sum(logQt[seq_pair[:, 1], seq_pair[:, 0]] * multiplicity)


def optimal_t_compressed(gtr, seq_pair, multiplicity, tol=1e-10):
    """GTR.optimal_t_compressed(profiles=False), gtr.py:816-925."""
    from scipy.optimize import minimize_scalar
    hamming = np.sum(multiplicity[seq_pair[:, 1] != seq_pair[:, 0]]) / np.sum(multiplicity)
    opt = minimize_scalar(lambda s: -1.0 * prob_t_compressed(gtr, seq_pair, multiplicity, s ** 2),
                          bracket=[-np.sqrt(MAX_BRANCH_LENGTH), np.sqrt(hamming), np.sqrt(MAX_BRANCH_LENGTH)], tol=tol, method='brent')
    new_len = opt['x'] ** 2
    if 'success' in opt and opt['success'] is not True and opt['success'] != True:      # noqa: E712
        new_len = hamming
    return new_len


# ---------------------------------------------------------------------------------------
# N4: SeqGen.  Reference: SeqGen.evolve + sample_from_profile (seqgen.py:19-67), GTR.evolve (gtr.py:997-1025).
# ---------------------------------------------------------------------------------------
def seqgen(flat, gtr, uniforms, root_idx=None):
    """State indices [n_nodes, L] evolved down the tree: root = argmax(cumsum(Pi) > u_0) unless given,
    child = argmax(cumsum(expQt(t_c)[:, parent state]) > u_c) -- exactly the reference's draw when
    `uniforms[n]` is the rng.random(L) vector it used for node n."""
    parent, t = flat['parent'], flat['t']
    n_nodes = parent.shape[0]
    L = uniforms.shape[1]
    states = np.zeros((n_nodes, L), dtype=np.uint8)
    site = np.arange(L)
    if root_idx is not None:
        states[0] = root_idx
    else:
        Pi = gtr.Pi
        prof = Pi.T if Pi.ndim == 2 else np.repeat([Pi], L, axis=0)         # seqgen.py:53-58
        states[0] = np.argmax(prof.cumsum(axis=1).T > uniforms[0], axis=0)
    for n in range(1, n_nodes):
        P = gtr.expQt(t[n])                          # (q, q) or (q, q, L): P[i, j(, a)] = Prob(child i | parent j)
        sp = states[parent[n]].astype(int)
        prof = P[:, sp, site].T if P.ndim == 3 else P[:, sp].T               # profile.dot(expQt(t).T) for one-hot profiles
        states[n] = np.argmax(prof.cumsum(axis=1).T > uniforms[n], axis=0)
    return states
