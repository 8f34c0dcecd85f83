"""Host-side substitution models (parameters + eigendecomposition).

Only what the marginal path needs lives here: the model parameters, the
symmetrised eigendecomposition (reference: gtr.py:612-629) and small host
helpers.  The exponentiation exp(Qt) for all branches runs on the device
(csrc/ttb_kernels.cu: expqt_batch_kernel); `expQt` below is the host formula
kept for the mirror API and unit tests.  The reference's standard-model
parameter tables (nuc_models.py / aa_models.py) are not rewritten: pass any
reference GTR object to the drop-in, or build models with `GTR.custom`.
"""
import numpy as np
from . import config as ttconf
from .seq_utils import alphabets, profile_maps, alphabet_synonyms


def avg_transition(W, pi, gap_index=None):
    """Average substitution rate of (W, pi), ignoring the gap state (gtr.py:7-11)."""
    if gap_index is None:
        return np.einsum('i,ij,j', pi, W, pi)
    return (np.einsum('i,ij,j', pi, W, pi) - np.sum(pi * W[:, gap_index]) * pi[gap_index]) / (1 - pi[gap_index])


def eig_single_site(W, p):
    """Eigen-system of Q_ij = W_ij p_i via the symmetric matrix W o sqrt(p p^T)
    (gtr.py:612-629): returns eigenvals, v (right, columns) and v_inv."""
    sp = np.sqrt(p)
    symQ = W * np.outer(sp, sp)
    np.fill_diagonal(symQ, -np.sum(W * p, axis=1))
    eigvals, eigvecs = np.linalg.eigh(symQ)
    tmp_v = eigvecs.T * sp
    one_norm = np.sum(np.abs(tmp_v), axis=1)
    return eigvals, tmp_v.T / one_norm, (eigvecs * one_norm).T / sp


class GTR(object):
    """General time-reversible model, single set of parameters for all sites."""
    is_site_specific = False

    def __init__(self, alphabet='nuc', prof_map=None, logger=None):
        if isinstance(alphabet, str):
            if alphabet not in alphabet_synonyms:
                raise AttributeError('Unknown alphabet type specified')
            name = alphabet_synonyms[alphabet]
            self.alphabet = alphabets[name]
            self.profile_map = dict(profile_maps[name])
        else:
            self.alphabet = np.array(alphabet)
            if prof_map is None:
                self.profile_map = {s: x for s, x in zip(self.alphabet, np.eye(len(self.alphabet)))}
            else:
                self.profile_map = dict(prof_map)
        self.logger = logger or (lambda *a, **k: None)
        self.n_states = len(self.alphabet)
        self.state_index = {s: i for i, s in enumerate(self.alphabet)}
        amb = [c for c, x in self.profile_map.items() if np.sum(x) == self.n_states]
        self.ambiguous = ('N' if 'N' in amb else amb[0]) if amb else None
        self.gap_index = self.state_index.get('-', None)
        self.assign_rates()

    # -- parameters ---------------------------------------------------------
    @property
    def mu(self):
        return self._mu

    @mu.setter
    def mu(self, value):
        self.assign_rates(mu=value, pi=self.Pi, W=self.W)

    @property
    def Pi(self):
        return self._Pi

    @property
    def W(self):
        return self._W

    @property
    def Q(self):
        Q = (self.W * self.Pi).T
        np.fill_diagonal(Q, -np.sum(Q, axis=0) + np.diag(Q))
        return Q

    def assign_rates(self, mu=1.0, pi=None, W=None):
        """Set (mu, Pi, W); W is rescaled so that the average rate is carried by
        mu alone (gtr.py:234-278)."""
        n = self.n_states
        self._mu = mu
        Pi = np.array(pi, dtype=float) if (pi is not None and len(pi) == n) else np.ones(n)
        self._Pi = Pi / np.sum(Pi)
        if W is None or np.shape(W) != (n, n):
            W = np.ones((n, n))
        else:
            W = np.array(W, dtype=float)
        np.fill_diagonal(W, 0)
        average_rate = avg_transition(W, self._Pi, gap_index=self.gap_index)
        self._W = W / average_rate
        self._mu = self._mu * average_rate
        self.eigenvals, self.v, self.v_inv = eig_single_site(self._W, self._Pi)

    @classmethod
    def custom(cls, mu=1.0, pi=None, W=None, **kwargs):
        gtr = cls(**kwargs)
        gtr.assign_rates(mu=mu, pi=pi, W=W)
        return gtr

    @classmethod
    def jc69(cls, mu=1.0, alphabet='nuc', **kwargs):
        """Equal rates, equal frequencies."""
        return cls.custom(mu=mu, alphabet=alphabet, **kwargs)

    @classmethod
    def random(cls, mu=1.0, alphabet='nuc', rng=None):
        """Random reversible model (gamma-distributed rates and frequencies)."""
        rng = rng or np.random.default_rng()
        gtr = cls(alphabet=alphabet)
        n = gtr.n_states
        pi = rng.gamma(2.0, size=n) + 0.1
        tmp = np.tril(rng.gamma(3.0, size=(n, n)), k=-1)
        gtr.assign_rates(mu=mu, pi=pi / pi.sum(), W=tmp + tmp.T)
        return gtr

    # -- host formulas (small inputs / tests only) --------------------------
    def _exp_lt(self, t):
        log_val = self.mu * t * self.eigenvals
        if np.any(log_val > 10):
            raise ValueError('Error in computing exp(Q * t): Q has positive eigenvalues or the branch length t is too large.')
        return np.exp(log_val)

    def expQt(self, t):
        """max(0, v diag(exp(mu t lambda)) v_inv); [i, j] = P(child=i | parent=j) (gtr.py:1051-1067)."""
        return np.maximum(0, self.v.dot(np.diag(self._exp_lt(t)).dot(self.v_inv)))

    def average_rate(self):
        return self.mu * avg_transition(self.W, self.Pi, gap_index=self.gap_index)


class GTRSiteSpecific(GTR):
    """Per-site frequencies Pi (q, L) and rates mu (L), shared W; the transition
    matrices are interpolated linearly in t on a 61-point grid unless
    approximate=False (gtr_site_specific.py:331-371)."""
    is_site_specific = True

    def __init__(self, seq_len=1, approximate=True, **kwargs):
        self.seq_len = seq_len
        self.approximate = approximate
        super(GTRSiteSpecific, self).__init__(**kwargs)

    @property
    def mu(self):
        return self._mu

    @mu.setter
    def mu(self, value):
        self.assign_rates(mu=value, pi=self.Pi, W=self.W)

    def assign_rates(self, mu=1.0, pi=None, W=None):
        """gtr_site_specific.py:47-114."""
        n = self.n_states
        if np.isscalar(mu):
            self._mu = mu * np.ones(self.seq_len)
        else:
            self._mu = np.array(mu, dtype=float)
            self.seq_len = self._mu.shape[0]
        if pi is not None and np.ndim(pi) == 2 and np.shape(pi)[0] == n:
            self.seq_len = np.shape(pi)[1]
            Pi = np.array(pi, dtype=float)
        elif pi is not None:
            if len(pi) != n:
                raise ValueError('GTRSiteSpecific: length of equilibrium frequency vector does not match alphabet length')
            Pi = np.repeat([np.asarray(pi, dtype=float)], self.seq_len, axis=0).T
        else:
            Pi = np.ones((n, self.seq_len))
        if self._mu.shape[0] != Pi.shape[1]:
            raise ValueError('GTRSiteSpecific: length of rate vector and equilibrium frequency vector must match!')
        self._Pi = Pi / np.sum(Pi, axis=0)
        if W is None:
            W = np.ones((n, n))
        elif np.shape(W) != (n, n):
            raise ValueError('GTRSiteSpecific: size of substitution matrix does not match alphabet length')
        else:
            W = 0.5 * (np.array(W, dtype=float) + np.array(W, dtype=float).T)
        np.fill_diagonal(W, 0)
        average_rate = np.einsum('ia,ij,ja', self._Pi, W, self._Pi) / self.seq_len
        self._W = W / average_rate
        self._mu = self._mu * average_rate
        self._eig()
        self.rate_scale = self.average_rate().mean()

    def _eig(self):
        """Per-site eigen-systems (gtr_site_specific.py:312-329), batched."""
        W, Pi = self._W, self._Pi                      # (q,q), (q,L)
        sp = np.sqrt(Pi)                               # (q,L)
        symQ = np.einsum('ij,ia,ja->aij', W, sp, sp)   # (L,q,q)
        diag = -np.einsum('ij,ja->ai', W, Pi)          # -sum_j W_ij p_j
        ar = np.arange(self.n_states)
        symQ[:, ar, ar] = diag
        eigvals, eigvecs = np.linalg.eigh(symQ)        # (L,q), (L,q,q)
        tmp_v = np.swapaxes(eigvecs, 1, 2) * sp.T[:, None, :]      # eigvecs.T * sp
        one_norm = np.sum(np.abs(tmp_v), axis=2)                   # (L,q)
        v = np.swapaxes(tmp_v, 1, 2) / one_norm[:, None, :]        # tmp_v.T / one_norm
        v_inv = np.swapaxes(eigvecs * one_norm[:, None, :], 1, 2) / sp.T[:, None, :]
        self.eigenvals = np.ascontiguousarray(eigvals.T)            # (q,L)
        # reference layout (np.swapaxes(list_of_matrices, 0, -1), gtr_site_specific.py:327-329):
        # v[k, i, a] = V_a[i, k] and v_inv[j, k, a] = Vinv_a[k, j]
        self.v = np.ascontiguousarray(np.transpose(v, (2, 1, 0)))
        self.v_inv = np.ascontiguousarray(np.transpose(v_inv, (2, 1, 0)))

    def average_rate(self):
        """Per-site average rate; no gap correction (gtr_site_specific.py:491-495)."""
        return np.einsum('a,ia,ij,ja->a', self.mu, self.Pi, self.W, self.Pi)

    @classmethod
    def custom(cls, mu=1.0, pi=None, W=None, **kwargs):
        gtr = cls(**kwargs)
        gtr.assign_rates(mu=mu, pi=pi, W=W)
        return gtr

    @classmethod
    def random(cls, L=1, avg_mu=1.0, alphabet='nuc', pi_dirichlet_alpha=1, W_dirichlet_alpha=3.0,
               mu_gamma_alpha=3.0, rng=None):
        """Random per-site model; same draw order as gtr_site_specific.py:116-172
        so that the same `rng` yields the same model."""
        rng = rng or np.random.default_rng()
        gtr = cls(alphabet=alphabet, seq_len=L)
        n = gtr.n_states
        pi = 1.0 * rng.gamma(pi_dirichlet_alpha, size=(n, L)) if pi_dirichlet_alpha else np.ones((n, L))
        pi /= pi.sum(axis=0)
        tmp = 1.0 * rng.gamma(W_dirichlet_alpha, size=(n, n)) if W_dirichlet_alpha else np.ones((n, n))
        tmp = np.tril(tmp, k=-1)
        W = tmp + tmp.T
        mu = rng.gamma(mu_gamma_alpha, size=(L,)) if mu_gamma_alpha else np.ones(L)
        gtr.assign_rates(mu=mu, pi=pi, W=W)
        gtr.assign_rates(mu=gtr.mu * (avg_mu / np.mean(gtr.average_rate())), pi=gtr.Pi, W=gtr.W)
        return gtr

    def expQt(self, t):
        """Per-site transition matrices stacked as (q, q, L): [i, j, a] = Prob(i <- j | site a, t)
        (gtr_site_specific.py:350-371).  Host-side accessor for single branches (get_branch_mutation_matrix); the
        passes evaluate it on the device.  With `approximate` and t * rate_scale < 10 the reference interpolates the
        matrices linearly on its 61-point grid; the matrices are linear in the eigen-factors exp(lambda mu t), so
        interpolating those gives the same result."""
        from .flatten import expqt_t_grid
        lam = self.eigenvals * self.mu
        if getattr(self, 'approximate', True) and t * self.rate_scale < 10:
            grid = expqt_t_grid(self.rate_scale)
            hi = int(np.clip(np.searchsorted(grid, t), 1, grid.shape[0] - 1))
            lo = hi - 1
            e = np.exp(lam * grid[lo])
            e = e + (np.exp(lam * grid[hi]) - e) * ((t - grid[lo]) / (grid[hi] - grid[lo]))
        else:
            e = np.exp(lam * t)
        return np.einsum('jia,ja,kja->ika', self.v, e, self.v_inv)


def infer_gtr_from_counts(nij, Ti, root_state, fixed_pi=None, pc=1.0, gap_limit=0.01, alphabet='nuc',
                          prof_map=None, logger=None):
    """Fit (mu, Pi, W) to the substitution statistics n_ij (changes j -> i), T_i
    (time spent in state i) and the root state counts: the fixed-point iteration of
    the reference's GTR.infer (gtr.py:491-599), consumer of the device-side counts
    (csrc counts_kernel).  Host-only: q x q arithmetic."""
    gtr = GTR(alphabet=alphabet, prof_map=prof_map, logger=logger)
    nij = np.array(nij, dtype=float)
    Ti = np.asarray(Ti, dtype=float)
    root_state = np.asarray(root_state, dtype=float)
    dp, Nit = 1e-5, 40
    pc_mat = pc * np.ones_like(nij)
    np.fill_diagonal(pc_mat, 0.0)
    np.fill_diagonal(nij, 0.0)
    pi_old = np.zeros_like(Ti)
    pi = np.ones_like(Ti) if fixed_pi is None else np.array(fixed_pi, dtype=float)
    pi /= pi.sum()
    W_ij = np.ones_like(nij)
    mu = (nij.sum() + pc) / (Ti.sum() + pc)
    count = 0
    while np.linalg.norm(pi_old - pi) > dp and count < Nit:
        count += 1
        pi_old = np.copy(pi)
        W_ij = (nij + nij.T + 2 * pc_mat) / mu / (np.outer(pi, Ti) + np.outer(Ti, pi) + ttconf.TINY_NUMBER + 2 * pc_mat)
        np.fill_diagonal(W_ij, 0)
        W_ij = W_ij / avg_transition(W_ij, pi, gap_index=gtr.gap_index)
        if fixed_pi is None:
            pi = (np.sum(nij + pc_mat, axis=1) + root_state) / (
                ttconf.TINY_NUMBER + mu * np.dot(W_ij, Ti) + root_state.sum() + np.sum(pc_mat, axis=1))
            pi /= pi.sum()
            mu = (nij.sum() + pc) / (np.sum(pi * (W_ij.dot(Ti))) + pc)
        else:
            mu = (nij.sum() + pc) / (np.sum(pi * (W_ij.dot(pi))) * Ti.sum() + pc)
    if gtr.gap_index is not None:
        # the reference resets the gap frequency to gap_limit unconditionally (gtr.py:584-596)
        pi[gtr.gap_index] = gap_limit
        pi /= pi.sum()
    gtr.assign_rates(mu=mu, W=W_ij, pi=pi)
    return gtr


def infer_site_specific_gtr_from_counts(sub_ija, T_ia, root_state, pc=1.0, gap_limit=0.01, Nit=30, dp=1e-5,
                                        alphabet='nuc', prof_map=None, logger=None):
    """Per-site model from per-site statistics: the fixed point of n_ija + pc = pi_ia W_ij mu_a (T_ja + pc + root)
    that the reference's GTR_site_specific.infer iterates (gtr_site_specific.py:207-310).  sub_ija (q, q, L):
    expected changes j -> i at site a, T_ia (q, L): time spent in state i, root_state (q, L): root profile.
    Consumer of the device-side per-pattern counts (ttb_mutation_counts_per_site).  Host-only arithmetic on
    (q, L) arrays, in the reference's order of operations."""
    gtr = GTRSiteSpecific(alphabet=alphabet, prof_map=prof_map, logger=logger)
    q = gtr.n_states
    L = sub_ija.shape[-1]
    n_ija = np.array(sub_ija, dtype=float)
    ar = np.arange(q)
    n_ija[ar, ar, :] = 0
    n_ij = n_ija.sum(axis=-1)
    m_ia = np.sum(n_ija, axis=1) + root_state + pc
    n_a = n_ija.sum(axis=1).sum(axis=0) + pc
    Lambda = np.sum(root_state, axis=0) + q * pc
    p_ia_old = np.zeros((q, L))
    p_ia = np.ones((q, L)) / q
    mu_a = np.ones(L)
    W_ij = np.ones((q, q)) - np.eye(q)
    n_iter = 0
    while np.linalg.norm(p_ia_old - p_ia) > dp and n_iter < Nit:
        n_iter += 1
        p_ia_old = np.copy(p_ia)
        S_ij = np.einsum('a,ia,ja', mu_a, p_ia, T_ia)
        W_ij = (n_ij + n_ij.T + pc) / (S_ij + S_ij.T + pc)
        avg_pi = p_ia.mean(axis=-1)
        average_rate = W_ij.dot(avg_pi).dot(avg_pi)
        W_ij = W_ij / average_rate
        mu_a *= average_rate
        p_ia = m_ia / (mu_a * np.dot(W_ij, T_ia) + Lambda)
        p_ia = p_ia / p_ia.sum(axis=0)
        mu_a = n_a / (pc + np.einsum('ia,ij,ja->a', p_ia, W_ij, T_ia))
    if n_iter >= Nit and logger is not None:
        logger('WARNING: maximum number of iterations has been reached in GTR inference', 3, warn=True)
    if gtr.gap_index is not None:
        # like the single-site model, the reference resets the gap frequency of EVERY site (:293-306)
        for a in range(L):
            p_ia[gtr.gap_index, a] = gap_limit
            p_ia[:, a] /= p_ia[:, a].sum()
    gtr.assign_rates(mu=mu_a, W=W_ij, pi=p_ia)
    return gtr
