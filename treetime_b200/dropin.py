"""Drop-in for the REAL TreeTime: a mixin that puts the device implementation of the marginal
path in front of `treetime.TreeAnc` in the MRO, so that `ClockTree` / `TreeTime.run` /
`treetime ancestral --marginal` call it unchanged.

    from treetime import TreeAnc, TreeTime
    from treetime_b200.dropin import accelerate
    B200TreeAnc = accelerate(TreeAnc)          # class B200TreeAnc(B200MarginalMixin, TreeAnc)
    B200TreeTime = accelerate(TreeTime)
    tt = B200TreeTime(tree=..., aln=..., gtr=..., dates=...)
    tt.run(branch_length_mode='marginal', ...)

What is overridden (reference lines in treeanc.py): `_ml_anc_marginal` (:762-812),
`optimize_tree_marginal` (:1297-1360), `optimal_marginal_branch_length` (:1272-1295),
`infer_gtr` marginal branch (:1500-1632), `optimize_gtr_rate` (:1679-1708).  Everything the
device path does not cover -- per-branch masks (ARG),
site-specific models with more than 8 states, alphabets without compiled kernels, joint / Fitch reconstruction --
falls through to the reference's own implementation (`super()`).

Per-node results stay on the device.  The reference reads them as plain attributes of the
Bio.Phylo clades (`node.marginal_profile`, `node.marginal_subtree_LH`,
`node.marginal_outgroup_LH`, `node._cseq`); a `__getattr__` hook installed on the clade class
fetches them on first access and caches them on the node until the next pass.
"""
import numpy as np

from .device_mixin import DeviceMarginalMixin, Unsupported, SUBTREE, OUTGROUP, PROFILE

_LAZY = {'marginal_subtree_LH': SUBTREE, 'marginal_outgroup_LH': OUTGROUP, 'marginal_profile': PROFILE}
_hooked = set()


def _lazy_getattr(node, name):
    """Called only when normal attribute lookup fails (so cached values and everything the
    reference sets itself take precedence)."""
    if name in _LAZY or name == '_cseq' or name == 'branch_state':
        tt = node.__dict__.get('tt')
        if tt is not None and getattr(tt, '_b200_live', False) and '_fid' in node.__dict__:
            try:
                if name == 'branch_state':
                    # N2: pair counts of the branch from the device instead of GTR.state_pair on host sequences
                    if node.up is None or getattr(node, 'mask', None) is not None or tt.gtr.is_site_specific:
                        raise AttributeError(name)
                    try:
                        val = tt._branch_state(node)
                    except tt._missing_data_error:
                        raise AttributeError(name)      # the reference raises its own error for a leaf without sequence
                elif name == '_cseq':
                    if node.is_terminal() and not tt.reconstructed_tip_sequences:
                        raise AttributeError(name)
                    val = tt._node_cseq(node)
                else:
                    val = tt._node_array(node, _LAZY[name])
            except (ValueError, RuntimeError) as e:
                raise AttributeError(str(e))
            node.__dict__[name] = val
            return val
    raise AttributeError(name)


def _install_hook(clade_cls):
    if clade_cls in _hooked:
        return
    if '__getattr__' in clade_cls.__dict__:
        prev = clade_cls.__dict__['__getattr__']

        def chained(node, name, _prev=prev):
            try:
                return _lazy_getattr(node, name)
            except AttributeError:
                return _prev(node, name)
        clade_cls.__getattr__ = chained
    else:
        clade_cls.__getattr__ = _lazy_getattr
    _hooked.add(clade_cls)


class B200MarginalMixin(DeviceMarginalMixin):
    """Place before treetime.TreeAnc (or ClockTree / TreeTime) in the MRO."""

    def __init__(self, *args, device=0, comm=None, engine_factory=None, **kwargs):
        self._init_device(device=device, comm=comm, engine_factory=engine_factory)
        self._b200_live = False
        super(B200MarginalMixin, self).__init__(*args, **kwargs)

    # -- the reference edits `clades` lists in place (prune_short_branches, polytomy resolution, reroot)
    #    without telling anybody: compare the cached flattening with the live tree at the start of
    #    every pass.  Between passes node._fid keeps addressing the device's (older) numbering, which
    #    mirrors the reference's stale per-node attributes.
    def _topology_changed(self):
        topo = self._topo
        if topo is None or topo.nodes[0] is not self.tree.root:
            return True
        cp, ci, nodes = topo.child_ptr, topo.child_idx, topo.nodes
        for i, n in enumerate(nodes):
            kids = n.clades
            b = cp[i]
            if len(kids) != cp[i + 1] - b or any(nodes[ci[b + k]] is not c for k, c in enumerate(kids)):
                return True
        return False

    def _refresh_topology(self):
        if self._topology_changed() and self._topo is not None:
            for n in self._topo.nodes:
                n.__dict__.pop('_fid', None)
        return DeviceMarginalMixin._refresh_topology(self)

    def _drop_node_caches(self):
        self._b200_live = False
        if self._topo is not None:
            for n in self._topo.nodes:
                d = n.__dict__
                for k in ('marginal_subtree_LH', 'marginal_outgroup_LH', 'marginal_profile', 'branch_state'):
                    d.pop(k, None)
        self._cache = {}
        self._seq_cache = {}

    def _device_ok(self):
        from . import _lib
        g = self.gtr
        if getattr(g, 'is_site_specific', False) and int(g.n_states) > 8:
            return 'site-specific models with more than 8 states run in the reference'
        if self._engine is None and self._engine_factory.__module__ == 'treetime_b200.device_mixin':
            if not _lib.load().ttb_supports_n_states(int(g.n_states)):
                return 'no kernels for %d states' % g.n_states
        if self.data.compressed_length < 1:
            return 'empty alignment'
        return None

    # -- the pass ---------------------------------------------------------------------------
    def _ml_anc_marginal(self, sample_from_profile=False, reconstruct_tip_states=False, debug=False, **kwargs):
        why = self._device_ok()
        if why is None:
            why = self._mask_problem()
        if why is None:
            try:
                # N_diff against a previous reconstruction that did not come from the device (joint /
                # Fitch / reference fallback) has to be counted on the host (treeanc.py:925-926)
                prev_live = self._b200_live and self.sequence_reconstruction == 'marginal'
                old = None
                if self.sequence_reconstruction and not prev_live:
                    old = {}
                    for n in self.tree.find_clades():
                        if n.up is not None and (reconstruct_tip_states or not n.is_terminal()):
                            try:
                                c = n.cseq
                            except ValueError:      # node created after the last reconstruction
                                c = None
                            if c is not None:
                                old[id(n)] = np.array(c)
                topo = self._refresh_topology()      # (snapshots the device's states if the tree was edited: N_diff)
                self._drop_node_caches()
                _install_hook(type(self.tree.root))
                # the reference keeps stale per-node arrays from an earlier (reference) pass: drop them
                for n in topo.nodes:
                    d = n.__dict__
                    for k in ('marginal_subtree_LH', 'marginal_outgroup_LH', 'marginal_profile', '_cseq',
                              'marginal_log_Lx', 'marginal_subtree_LH_prefactor', 'branch_state'):
                        d.pop(k, None)
                N_diff = DeviceMarginalMixin._ml_anc_marginal(self, sample_from_profile=sample_from_profile,
                                                              reconstruct_tip_states=reconstruct_tip_states, debug=debug)
                root = self.tree.root
                if getattr(root, '_cseq_override', None) is not None:
                    root.__dict__['_cseq'] = root._cseq_override
                self._b200_live = True
                if old is not None:
                    N_diff = 0
                    for n in topo.nodes[1:]:
                        if reconstruct_tip_states or not n.is_terminal():
                            N_diff += int((n.cseq != old[id(n)]).sum()) if id(n) in old else self.data.compressed_length
                return N_diff
            except Unsupported as e:
                why = str(e)
        self.logger('B200: falling back to the reference implementation (%s)' % why, 2)
        if self._b200_live and self._topo is not None:
            # the reference compares against node._cseq of the previous pass: bring them to the host
            for n in self._topo.nodes:
                if '_cseq' not in n.__dict__ and (self.reconstructed_tip_sequences or not n.is_terminal()):
                    n.__dict__['_cseq'] = self._node_cseq(n)
        self._drop_node_caches()
        return super(DeviceMarginalMixin, self)._ml_anc_marginal(sample_from_profile=sample_from_profile,
                                                               reconstruct_tip_states=reconstruct_tip_states,
                                                               debug=debug, **kwargs)

    def _ml_anc_joint(self, sample_from_profile=False, reconstruct_tip_states=False, debug=False, **kwargs):
        """Joint ML reconstruction (treeanc.py:934-1080) on the device, with the same fallbacks."""
        why = self._device_ok()
        if why is None and getattr(self.gtr, 'is_site_specific', False):
            why = 'site-specific model'
        if why is None and any(getattr(n, 'mask', None) is not None for n in self.tree.find_clades()):
            why = 'per-branch masks (ARG mode)'
        if why is None:
            try:
                prev_live = self._b200_live and self.sequence_reconstruction in ('joint', 'marginal')
                old = None
                if self.sequence_reconstruction and not prev_live:
                    old = {}
                    for n in self.tree.find_clades():
                        if n.up is not None and (reconstruct_tip_states or not n.is_terminal()):
                            try:
                                c = n.cseq
                            except ValueError:
                                c = None
                            if c is not None:
                                old[id(n)] = np.array(c)
                topo = self._refresh_topology()
                self._drop_node_caches()
                _install_hook(type(self.tree.root))
                for n in topo.nodes:
                    d = n.__dict__
                    for k in ('marginal_subtree_LH', 'marginal_outgroup_LH', 'marginal_profile', '_cseq', 'joint_Lx', 'joint_Cx',
                              'seq_idx', 'branch_state'):
                        d.pop(k, None)
                N_diff = DeviceMarginalMixin._ml_anc_joint(self, sample_from_profile=sample_from_profile,
                                                           reconstruct_tip_states=reconstruct_tip_states, debug=debug)
                self._b200_live = True
                if old is not None:
                    N_diff = 0
                    for n in topo.nodes[1:]:
                        if reconstruct_tip_states or not n.is_terminal():
                            N_diff += int((n.cseq != old[id(n)]).sum()) if id(n) in old else self.data.compressed_length
                return N_diff
            except Unsupported as e:
                why = str(e)
        self.logger('B200: falling back to the reference implementation (%s)' % why, 2)
        if self._b200_live and self._topo is not None:
            for n in self._topo.nodes:
                if '_cseq' not in n.__dict__ and (self.reconstructed_tip_sequences or not n.is_terminal()):
                    n.__dict__['_cseq'] = self._node_cseq(n)
        self._drop_node_caches()
        return super(DeviceMarginalMixin, self)._ml_anc_joint(sample_from_profile=sample_from_profile,
                                                            reconstruct_tip_states=reconstruct_tip_states, debug=debug, **kwargs)

    # accessors: the reference's own versions work through the lazy node attributes
    def sequence_LH(self, *args, **kwargs):
        return super(DeviceMarginalMixin, self).sequence_LH(*args, **kwargs)

    def marginal_branch_profile(self, node):
        return super(DeviceMarginalMixin, self).marginal_branch_profile(node)

    def get_branch_mutation_matrix(self, node, full_sequence=False):
        return super(DeviceMarginalMixin, self).get_branch_mutation_matrix(node, full_sequence=full_sequence)

    # N2 branch lengths: by default the reference's per-branch code runs on the lazily provided
    # node.branch_state (bit-identical results, no sequences cross PCIe); with
    # `batched_joint_branch_lengths = True` all branches are optimised in one lock-step Brent.
    batched_joint_branch_lengths = False

    def add_branch_state(self, node):
        if self._b200_live and getattr(node, 'mask', None) is None and not self.gtr.is_site_specific:
            return DeviceMarginalMixin.add_branch_state(self, node)
        return super(DeviceMarginalMixin, self).add_branch_state(node)

    def optimal_branch_length(self, node):
        return super(DeviceMarginalMixin, self).optimal_branch_length(node)

    def optimize_branch_lengths_joint(self, **kwargs):
        if (self.batched_joint_branch_lengths and self._b200_live and not self.gtr.is_site_specific
                and not any(getattr(n, 'mask', None) is not None for n in self.tree.find_clades())):
            try:
                return DeviceMarginalMixin.optimize_branch_lengths_joint(self, **kwargs)
            except Unsupported:
                pass
        return super(DeviceMarginalMixin, self).optimize_branch_lengths_joint(**kwargs)

    def optimize_tree_marginal(self, *args, **kwargs):
        if self._device_ok() is None and self._mask_problem() is None:
            try:
                return DeviceMarginalMixin.optimize_tree_marginal(self, *args, **kwargs)
            except Unsupported:
                pass
        return super(DeviceMarginalMixin, self).optimize_tree_marginal(*args, **kwargs)

    def optimal_marginal_branch_length(self, node, tol=1e-10):
        if self._b200_live:
            return DeviceMarginalMixin.optimal_marginal_branch_length(self, node, tol=tol)
        return super(DeviceMarginalMixin, self).optimal_marginal_branch_length(node, tol=tol)

    def infer_gtr(self, marginal=False, site_specific=False, **kwargs):
        if self._device_ok() is None and (marginal or (not site_specific and (self._b200_live or not self.sequence_reconstruction))):
            try:
                gtr = DeviceMarginalMixin.infer_gtr(self, marginal=marginal, site_specific=site_specific, **kwargs)
                return gtr
            except Unsupported:
                pass
        return super(DeviceMarginalMixin, self).infer_gtr(marginal=marginal, site_specific=site_specific, **kwargs)

    def _infer_gtr_from_counts(self, n_ij, T_i, root_state, fixed_pi, pc):
        """Use the reference's own GTR.infer so that the resulting object is a reference GTR."""
        from treetime.gtr import GTR
        return GTR.infer(n_ij, T_i, root_state, fixed_pi=fixed_pi, pc=pc, alphabet=self.gtr.alphabet,
                         logger=self.logger, prof_map=self.gtr.profile_map)

    def _infer_site_specific_gtr_from_counts(self, n_ija, T_ia, root_state, pc):
        """The reference's own GTR_site_specific.infer on the device's per-pattern statistics."""
        from treetime.gtr_site_specific import GTR_site_specific
        return GTR_site_specific.infer(n_ija, T_ia, pc=pc, root_state=root_state, logger=self.logger,
                                       alphabet=self.gtr.alphabet, prof_map=self.gtr.profile_map)

    def optimize_gtr_rate(self):
        if self._device_ok() is None and self._b200_live:
            return DeviceMarginalMixin.optimize_gtr_rate(self)
        return super(DeviceMarginalMixin, self).optimize_gtr_rate()


def branch_length_grid(mutation_length, one_mutation, n_grid_points, max_branch_length=4.0):
    """The t-grid on which BranchLenInterpolator tabulates a branch (restated from
    branch_len_interpolator.py:36-62 so that all branches can be tabulated in one device call)."""
    if mutation_length < np.min((1e-5, 0.1 * one_mutation)):  # zero-length branch
        short_range = 10 * one_mutation
        return np.concatenate([
            short_range * (np.linspace(0, 1.0, n_grid_points // 2)[:-1]),
            (short_range + (max_branch_length - short_range) * (np.linspace(0, 1.0, n_grid_points // 2 + 1) ** 2)),
        ])
    sigma = mutation_length
    grid_left = mutation_length * (1 - np.linspace(1, 0.0, n_grid_points // 3) ** 2.0)
    grid_zero = grid_left[1] * np.logspace(-20, 0, 6)[:5]
    grid_zero2 = grid_left[1] * np.linspace(0, 1, 10)[1:-1]
    grid_right = mutation_length + (3 * sigma * (np.linspace(0, 1, n_grid_points // 3) ** 2))
    far_grid = grid_right.max() + max_branch_length * np.linspace(0, 1, n_grid_points // 3) ** 2
    grid = np.concatenate((grid_zero, grid_zero2, grid_left, grid_right[1:], far_grid[1:]))
    grid.sort()
    return grid


class _PairProxy(object):
    """Stands in for `node.profile_pair = (pp, pc)` while ClockTree builds its branch-length
    interpolators: the tabulated log-likelihoods were computed on the device for all branches at
    once, so the two (L', q) arrays are only materialised if somebody really indexes the pair."""

    def __init__(self, tt, node, grid, values):
        self.tt, self.node, self.grid, self.values = tt, node, grid, values
        self._pair = None

    def materialize(self):
        if self._pair is None:
            self._pair = (self.node.marginal_outgroup_LH, self.node.marginal_subtree_LH)
        return self._pair

    def lookup(self, t):
        i = int(np.searchsorted(self.grid, t))
        for k in (i, i - 1, i + 1):
            if 0 <= k < self.grid.shape[0] and self.grid[k] == t:
                return self.values[k]
        return None

    def __getitem__(self, i):
        return self.materialize()[i]

    def __iter__(self):
        return iter(self.materialize())

    def __len__(self):
        return 2


class B200ClockMixin(object):
    """N1: batched branch-likelihood grids for BranchLenInterpolator in marginal mode
    (branch_len_interpolator.py:103-110, caller clock_tree.py:344-370).  All (branch x grid point)
    values of GTR.prob_t_profiles are evaluated in a few device calls from the resident messages;
    the reference's interpolator construction then runs unchanged on the tabulated numbers."""

    GRID_EVAL_CHUNK = 1 << 18

    def _tabulate_branch_grids(self):
        # every node of the live tree the device knows (nodes created since the last pass carry explicit arrays)
        nodes = [n for n in self.tree.find_clades() if n.up is not None and '_fid' in n.__dict__ and n._fid > 0]
        grids = [branch_length_grid(n.mutation_length, self.one_mutation, self.branch_grid_points) for n in nodes]
        fids = np.concatenate([np.full(g.shape[0], n._fid, dtype=np.int32) for n, g in zip(nodes, grids)])
        ts = np.concatenate(grids)
        vals = np.empty(ts.shape[0])
        for lo in range(0, ts.shape[0], self.GRID_EVAL_CHUNK):
            hi = min(ts.shape[0], lo + self.GRID_EVAL_CHUNK)
            f = self._engine.branch_objective(fids[lo:hi], ts[lo:hi])
            if self.comm.world_size > 1:
                f = self.comm.allreduce_sum(f)
            vals[lo:hi] = f
        out, k = {}, 0
        for n, g in zip(nodes, grids):
            out[id(n)] = _PairProxy(self, n, g, vals[k:k + g.shape[0]])
            k += g.shape[0]
        return out

    def marginal_branch_profile(self, node):
        tab = getattr(self, '_b200_grid_tab', None)
        if tab is not None and id(node) in tab:
            return tab[id(node)]
        return super(B200ClockMixin, self).marginal_branch_profile(node)

    def init_date_constraints(self, *args, **kwargs):
        use = (getattr(self, 'branch_length_mode', None) == 'marginal' and self._device_ok() is None and self.aln
               and not any(getattr(n, 'mask', None) is not None for n in self.tree.find_clades()))
        if not use:
            return super(B200ClockMixin, self).init_date_constraints(*args, **kwargs)
        if not self.sequence_reconstruction:     # clock_tree.py:335-338: everything but clock_rate is forwarded
            fwd = {k: v for k, v in kwargs.items() if k != 'clock_rate'}
            self.infer_ancestral_sequences('probabilistic', marginal=True, sample_from_profile='root', **fwd)
        if not self._b200_live:
            return super(B200ClockMixin, self).init_date_constraints(*args, **kwargs)
        self._b200_grid_tab = self._tabulate_branch_grids()
        gtr = self.gtr
        original = gtr.prob_t_profiles

        def prob_t_profiles(profile_pair, multiplicity, t, return_log=False, ignore_gaps=True):
            if isinstance(profile_pair, _PairProxy):
                v = profile_pair.lookup(t) if (return_log and ignore_gaps) else None
                if v is not None:
                    return v
                profile_pair = profile_pair.materialize()
            return original(profile_pair, multiplicity, t, return_log=return_log, ignore_gaps=ignore_gaps)

        gtr.prob_t_profiles = prob_t_profiles
        try:
            return super(B200ClockMixin, self).init_date_constraints(*args, **kwargs)
        finally:
            del gtr.prob_t_profiles
            self._b200_grid_tab = None


def accelerate(base):
    """Return `class B200<base>(B200MarginalMixin, base)`."""
    bases = (B200MarginalMixin, base)
    if hasattr(base, 'init_date_constraints'):        # ClockTree / TreeTime: also batch the branch grids
        bases = (B200ClockMixin, B200MarginalMixin, base)
    return type('B200' + base.__name__, bases, {'__doc__': B200MarginalMixin.__doc__})
