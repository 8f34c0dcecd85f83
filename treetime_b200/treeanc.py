"""TreeAnc: host-side mirror of the reference's TreeAnc for the marginal path.

Same constructor, method names, argument meaning, return values and error
behaviour as treetime.TreeAnc (treetime/treeanc.py) for
    infer_ancestral_sequences(marginal=True) / reconstruct_anc
    sequence_LH, optimize_tree(branch_length_mode='marginal'),
    optimize_tree_marginal, optimal_marginal_branch_length,
    marginal_branch_profile, get_branch_mutation_matrix,
    infer_gtr(marginal=True), optimize_gtr_rate, get_reconstructed_alignment
but every per-node numpy loop of the reference is one call into the CUDA engine
(libttb.so through treetime_b200.engine.Engine).  Per-node results
(`node.marginal_profile`, `node.marginal_subtree_LH`, `node.marginal_outgroup_LH`,
`node.cseq`) stay resident on the device and are fetched on first access.

With a communicator of world_size > 1 (treetime_b200.dist) every rank holds the
same tree/model and one contiguous block of the compressed patterns; scalars are
all-reduced, per-node arrays are all-gathered on access.

Not provided here (outside SURVEY.md §8): joint/Fitch reconstruction, masks
(ARG mode), sampling of non-root nodes from their profiles.  They raise
NotImplementedError; the drop-in mixin for the real TreeTime
(treetime_b200.dropin) falls back to the reference's own code for them.
"""
import numpy as np

from . import config as ttconf
from .brent import brent_lockstep
from .dist import SingleComm, default_comm, shard_bounds
from .flatten import FlatTopology, code_table, gtr_arrays
from .gtr import GTR, infer_gtr_from_counts
from .seq_utils import prof2seq, normalize_profile
from .sequence_data import SequenceData
from .tree import Node, Tree, read_newick

SUBTREE, OUTGROUP, PROFILE = 0, 1, 2


class TreeTimeError(Exception):
    """Base error (mirrors treetime.TreeTimeError)."""


class MissingDataError(TreeTimeError):
    """Tree or sequences are missing (mirrors treetime.MissingDataError)."""


class UnknownMethodError(TreeTimeError):
    """Unknown reconstruction / optimisation method (mirrors treetime.UnknownMethodError)."""


# -- lazy per-node views (the reference stores these as numpy attributes) --------------
def _lazy(which):
    def get(node):
        if getattr(node, 'tt', None) is None:
            raise AttributeError('node is not attached to a TreeAnc')
        return node.tt._node_array(node, which)
    return property(get)


def _cseq(node):
    """Compressed sequence of a node (treeanc.py:14-24)."""
    tt = node.tt
    if node.name in tt.data.compressed_alignment and not tt.reconstructed_tip_sequences:
        return tt.data.compressed_alignment[node.name]
    if node.is_terminal() and not tt.reconstructed_tip_sequences:
        return None
    return tt._node_cseq(node)


def _mutations(node):
    """(ancestral, position, derived) differences to the parent (treeanc.py:27-42)."""
    tt = node.tt
    if node.up is None:
        return []
    if node.is_terminal() and node.name not in tt.data.compressed_alignment:
        return []
    if (not tt.reconstructed_tip_sequences) and node.name in tt.data.compressed_alignment:
        child = tt.data.aln[node.name]
    else:
        child = tt.data.compressed_to_full_sequence(node.cseq)
    par = tt.data.compressed_to_full_sequence(node.up.cseq)
    L = min(par.shape[0], child.shape[0])
    pos = np.nonzero(par[:L] != child[:L])[0]
    return [(str(par[p]), int(p), str(child[p])) for p in pos]


Node.marginal_subtree_LH = _lazy(SUBTREE)
Node.marginal_outgroup_LH = _lazy(OUTGROUP)
Node.marginal_profile = _lazy(PROFILE)
Node.cseq = property(_cseq)
Node.mutations = property(_mutations)
Node.sequence = property(lambda n: n.tt.sequence(n, as_string=False))


def _default_engine_factory(n_states, device):
    from .engine import Engine          # raises if libttb.so is missing or there is no GPU
    return Engine(n_states, device=device)


class TreeAnc(object):
    def __init__(self, tree=None, aln=None, gtr=None, fill_overhangs=True, ref=None, verbose=0, ignore_gaps=True,
                 convert_upper=True, seq_multiplicity=None, log=None, compress=True, seq_len=None,
                 ignore_missing_alns=False, keep_node_order=False, rng_seed=None,
                 device=0, comm=None, engine_factory=None, **kwargs):
        if tree is None:
            raise TypeError('TreeAnc requires a tree!')
        self.verbose = verbose
        self.log_messages = set()
        self.ok = False
        self.data = None
        self.use_mutation_length = False
        self.ignore_gaps = ignore_gaps
        self.reconstructed_tip_sequences = False
        self.sequence_reconstruction = None
        self.ignore_missing_alns = ignore_missing_alns
        self.keep_node_order = keep_node_order
        self.rng = np.random.default_rng(seed=rng_seed)
        self.device = device
        self.comm = comm if comm is not None else default_comm()
        self._engine_factory = engine_factory or _default_engine_factory
        self._engine = None
        self._topo = None
        self._device_topology = None
        self._device_patterns = False
        self._cache = {}
        self._seq_cache = {}
        self._tree = None
        self.tree = tree
        self._gtr = None
        self.set_gtr(gtr or 'JC69', **kwargs)
        if ref is not None:
            raise NotImplementedError('sparse (VCF) alignments are read by the reference; pass a dense alignment')
        self.data = SequenceData(aln, compress=compress, convert_upper=convert_upper, fill_overhangs=fill_overhangs,
                                 ambiguous=self.gtr.ambiguous, sequence_length=seq_len, logger=self.logger)
        if self.gtr.is_site_specific and self.data.compress:
            raise TypeError('TreeAnc: sequence compression and site specific gtr models are incompatible!')
        self._check_alignment_tree_gtr_consistency()

    # -- logging -----------------------------------------------------------------------
    def logger(self, msg, level, warn=False, only_once=False):
        if only_once and msg in self.log_messages:
            return
        self.log_messages.add(msg)
        if level < self.verbose or (warn and level <= self.verbose):
            print(('  ' * level if level > 0 else '') + str(msg))

    # -- model / tree / alignment ---------------------------------------------------------
    @property
    def gtr(self):
        return self._gtr

    @gtr.setter
    def gtr(self, value):
        if not hasattr(value, 'eigenvals'):
            raise TypeError('TreeAnc.gtr setter: can not assign to GTR. GTR instance is required.')
        self._gtr = value

    def set_gtr(self, in_gtr, **kwargs):
        """treeanc.py:259-287."""
        if isinstance(in_gtr, str):
            if in_gtr.upper() not in ('JC69', 'JC', 'JUKES-CANTOR'):
                raise NotImplementedError("only 'JC69' can be built by name here; pass a GTR object "
                                          '(treetime_b200.gtr.GTR.custom or any reference GTR)')
            self._gtr = GTR.jc69(**kwargs)
        elif hasattr(in_gtr, 'eigenvals'):
            self._gtr = in_gtr
        else:
            raise TypeError('Cannot set GTR model in TreeAnc class: GTR or string expected')
        if getattr(self._gtr, 'ambiguous', None) is None:
            self.fill_overhangs = False

    @property
    def tree(self):
        return self._tree

    @tree.setter
    def tree(self, in_tree):
        """treeanc.py:315-371."""
        if isinstance(in_tree, Tree):
            self._tree = in_tree
        elif isinstance(in_tree, str):
            try:
                self._tree = read_newick(in_tree)
            except Exception:
                raise MissingDataError('TreeAnc: could not load tree! input was ' + str(in_tree)[:80])
        elif hasattr(in_tree, 'root') and hasattr(in_tree.root, 'clades'):
            self._tree = in_tree
        else:
            raise MissingDataError('TreeAnc: could not load tree! input was ' + str(in_tree)[:80])
        if self._tree.count_terminals() < 3:
            raise MissingDataError('TreeAnc: tree has only %d tips. Please check your tree!' % self._tree.count_terminals())
        for node in self._tree.find_clades():
            node.branch_length = node.branch_length if node.branch_length else 0.0
            node.original_length = node.branch_length
            node.mutation_length = node.branch_length
        self.prepare_tree()
        if self.data:
            self._check_alignment_tree_gtr_consistency()

    @property
    def aln(self):
        return self.data.aln

    @property
    def one_mutation(self):
        return 1.0 / self.data.full_length if self.data.full_length else np.nan

    @property
    def seq_len(self):
        return self.data.full_length

    sequence_length = seq_len

    def prepare_tree(self):
        """treeanc.py:446-493: root branch length, ladderize, node names, up-links."""
        self.sequence_reconstruction = False
        root = self.tree.root
        root.branch_length = 0.001
        root.mutation_length = root.branch_length
        root.mask = None
        if not self.keep_node_order:
            self.tree.ladderize()
        self._prepare_nodes()
        self._leaves_lookup = {n.name: n for n in self.tree.get_terminals()}

    def _prepare_nodes(self):
        root = self.tree.root
        root.up = None
        root.tt = self
        name_set = {n.name for n in self.tree.find_clades() if n.name}
        count = 0
        for clade in self.tree.get_nonterminals(order='preorder'):
            if clade.name is None:
                tmp = 'NODE_' + format(count, '07d')
                while tmp in name_set:
                    count += 1
                    tmp = 'NODE_' + format(count, '07d')
                clade.name = tmp
                name_set.add(tmp)
            count += 1
            for c in clade.clades:
                c.up = clade
                c.tt = self
        for clade in self.tree.find_clades():
            if not hasattr(clade, 'mask'):
                clade.mask = None
        root.dist2root = 0.0
        for clade in self.tree.get_nonterminals(order='preorder'):
            for c in clade.clades:
                c.dist2root = clade.dist2root + (c.mutation_length if hasattr(c, 'mutation_length') else c.branch_length)
        self._topo = None          # re-flatten lazily
        self._cache = {}
        self._seq_cache = {}

    @property
    def leaves_lookup(self):
        return self._leaves_lookup

    def _check_alignment_tree_gtr_consistency(self):
        """treeanc.py:395-444."""
        failed = 0
        n_tips = 0
        for l in self.tree.get_terminals():
            n_tips += 1
            if l.name not in self.data.compressed_alignment:
                self.logger("***WARNING: TreeAnc._check_alignment_tree_gtr_consistency: NO SEQUENCE FOR LEAF: '%s'" % l.name, 0, warn=True)
                failed += 1
                if not self.ignore_missing_alns and failed > n_tips / 3 and failed > self.tree.count_terminals() / 3:
                    raise MissingDataError('TreeAnc._check_alignment_tree_gtr_consistency: At least 30\\% terminal nodes '
                                           'cannot be assigned a sequence!\nAre you sure the alignment belongs to the tree?')
        # extend_profile (seq_utils.py:126-136): unknown characters are missing data
        present = np.unique(self.data.matrix)
        for b in present:
            c = chr(int(b))
            if c not in self.gtr.profile_map:
                self.gtr.profile_map[c] = np.ones(self.gtr.n_states)
                self.logger('WARNING: character %s is unknown. Treating it as missing information' % c, 1, warn=True)
        self._device_patterns = False
        self.ok = True

    def _branch_length_to_gtr(self, node):
        """treeanc.py:752-760."""
        if self.use_mutation_length:
            return max(ttconf.MIN_BRANCH_LENGTH * self.one_mutation, node.mutation_length)
        return max(ttconf.MIN_BRANCH_LENGTH * self.one_mutation, node.branch_length)

    # -- device synchronisation -------------------------------------------------------------
    def _flat(self):
        if self._topo is None:
            self._topo = FlatTopology(self.tree.root)
            for i, n in enumerate(self._topo.nodes):
                n._fid = i
        return self._topo

    def _shard(self):
        return shard_bounds(self.data.compressed_length, self.comm.rank, self.comm.world_size)

    def _tip_codes(self):
        """uint8 codes [n_tips, L'] + (n_codes, q) table from the compressed ASCII matrix."""
        topo = self._flat()
        chars, lut, table = code_table(self.gtr.profile_map, self.gtr.n_states)
        lut8 = np.full(256, 255, dtype=np.uint8)
        for c, i in lut.items():
            lut8[ord(c)] = i
        lo, hi = self._shard()
        rows = np.array([self.data._row.get(topo.nodes[n].name, -1) for n in topo.tip_nodes])
        codes = np.full((topo.n_tips, hi - lo), len(chars), dtype=np.uint8)     # default: missing = all ones
        have = rows >= 0
        codes[have] = lut8[self.data.compressed_matrix[rows[have], lo:hi]]
        if (codes == 255).any():
            raise KeyError('alignment contains characters that are not in the profile map')
        return codes, table

    def _sync_device(self):
        """Bring the engine up to date with tree topology, patterns, model and branch lengths."""
        topo = self._flat()
        if self._engine is None:
            self._engine = self._engine_factory(self.gtr.n_states, self.device)
        eng = self._engine
        sig = topo.signature()
        if sig != self._device_topology:
            eng.set_tree(topo.parent, topo.child_ptr, topo.child_idx, topo.tip_row)
            self._device_topology = sig
            self._device_patterns = False
        if not self._device_patterns:
            codes, table = self._tip_codes()
            lo, hi = self._shard()
            eng.set_patterns(codes, table, self.data.multiplicity()[lo:hi])
            self._device_patterns = True
        g = gtr_arrays(self.gtr)
        if g['site_specific']:
            lo, hi = self._shard()
            g = dict(g, eigenvals=g['eigenvals'][:, lo:hi], v=g['v'][:, :, lo:hi], v_inv=g['v_inv'][:, :, lo:hi],
                     Pi=g['Pi'][:, lo:hi], mu=g['mu'][lo:hi], t_grid=self._t_grid())
        tvec = np.array([self._branch_length_to_gtr(n) for n in topo.nodes], dtype=np.float64)
        lam = np.max(g['eigenvals']) * np.max(g['mu'])
        if lam * tvec[1:].max() > 10:
            raise ValueError('Error in computing exp(Q * t): Q has positive eigenvalues or the branch length t is too large. '
                             'This is most likely caused by incorrect input data.')     # gtr.py:1041-1047
        eng.set_gtr(g)
        eng.set_branch_lengths(tvec)
        self._t_last = tvec
        return eng

    def _t_grid(self):
        from .gtr import GTRSiteSpecific  # noqa: F401
        rs = self.gtr.rate_scale
        return (1.0 / rs) * np.concatenate((np.linspace(0, 0.1, 11)[:-1], np.linspace(0.1, 1, 21)[:-1],
                                            np.linspace(1, 5, 21)[:-1], np.linspace(5, 10, 11)))

    def _gather_patterns(self, x, axis=0):
        return x if self.comm.world_size == 1 else self.comm.allgather(x, axis=axis)

    def _node_array(self, node, which):
        key = (node._fid, which)
        if key not in self._cache:
            if not self.sequence_reconstruction and not (which == SUBTREE and self._engine is not None):
                raise AttributeError('marginal ancestral inference needs to be performed first!')
            if which == PROFILE and node.is_terminal() and not self.reconstructed_tip_sequences:
                raise AttributeError('tip profiles exist only after reconstruct_tip_states=True')
            self._cache[key] = self._gather_patterns(self._engine.node_array(node._fid, which), axis=0)
        return self._cache[key]

    def _node_cseq(self, node):
        if not self.sequence_reconstruction:
            raise ValueError('Ancestral sequences are not yet inferred')
        k = node._fid
        if k not in self._seq_cache:
            override = getattr(node, '_cseq_override', None)
            if override is not None:
                self._seq_cache[k] = override
            else:
                idx = self._gather_patterns(self._engine.seq_idx([k])[0], axis=0)
                self._seq_cache[k] = self.gtr.alphabet[idx]
        return self._seq_cache[k]

    # -- ancestral reconstruction ---------------------------------------------------------
    def reconstruct_anc(self, *args, **kwargs):
        return self.infer_ancestral_sequences(*args, **kwargs)

    def infer_ancestral_sequences(self, method='probabilistic', infer_gtr=False, marginal=False,
                                  reconstruct_tip_states=False, **kwargs):
        """treeanc.py:516-570.  Returns N_diff."""
        if not self.ok:
            raise MissingDataError('TreeAnc.infer_ancestral_sequences: ERROR, sequences or tree are missing')
        self.logger('TreeAnc.infer_ancestral_sequences with method: %s, %s' % (method, 'marginal' if marginal else 'joint'), 1)
        if method.lower() in ['ml', 'probabilistic']:
            if not marginal:
                raise NotImplementedError('joint ML reconstruction is outside the B200 hot path (SURVEY.md §8f N2); '
                                          'use marginal=True or the reference implementation')
        elif method.lower() in ['fitch', 'parsimony']:
            raise NotImplementedError('Fitch reconstruction is outside the B200 hot path; use the reference implementation')
        else:
            raise UnknownMethodError("Reconstruction method needs to be in ['ml', 'probabilistic', 'fitch', 'parsimony'], "
                                     "got '{}'".format(method))
        if infer_gtr:
            self.infer_gtr(marginal=marginal, **kwargs)
        return self._ml_anc_marginal(reconstruct_tip_states=reconstruct_tip_states, **kwargs)

    def _ml_anc_marginal(self, sample_from_profile=False, reconstruct_tip_states=False, debug=False, **kwargs):
        """treeanc.py:762-812: postorder, root, preorder -- one graph launch on the device."""
        self.logger('TreeAnc._ml_anc_marginal: type of reconstruction: Marginal', 2)
        if sample_from_profile == 'root':
            root_sample = True
        elif isinstance(sample_from_profile, bool):
            root_sample = sample_from_profile
            if sample_from_profile:
                raise NotImplementedError('sampling every node from its profile is not provided; '
                                          "sample_from_profile='root' is")
        else:
            raise ValueError("sample_from_profile must be a bool or 'root'")
        if any(getattr(n, 'mask', None) is not None for n in self._flat().nodes):
            raise NotImplementedError('per-branch masks (ARG mode) are not supported on the device path')
        eng = self._sync_device()
        topo = self._flat()
        eng.marginal(reconstruct_tips=reconstruct_tip_states)
        tot, nd = eng.results()
        if self.comm.world_size > 1:
            tot, nd = self.comm.allreduce_sum(np.array([tot, float(nd)]))
        self._cache = {}
        self._seq_cache = {}
        self.tree.sequence_LH = self._gather_patterns(eng.site_lh())
        self.tree.total_sequence_LH = float(tot)
        self.tree.sequence_marginal_LH = self.tree.total_sequence_LH
        n_rec = (topo.n_nodes - 1) if reconstruct_tip_states else (topo.n_nodes - topo.n_tips - 1)
        if self.sequence_reconstruction:
            N_diff = int(round(nd))
        else:
            N_diff = n_rec * self.data.compressed_length               # treeanc.py:927-928
        root = self.tree.root
        root._cseq_override = None
        self.reconstructed_tip_sequences = reconstruct_tip_states
        self.sequence_reconstruction = 'marginal'
        if root_sample:                                                 # treeanc.py:831-838, host RNG
            seq, _, _ = prof2seq(self._node_array(root, PROFILE), self.gtr, sample_from_prof=True, normalize=False, rng=self.rng)
            root._cseq_override = seq
        self.logger('TreeAnc._ml_anc_marginal: ...done', 3)
        return N_diff

    def sequence_LH(self, pos=None, full_sequence=False):
        """treeanc.py:691-716."""
        if not hasattr(self.tree, 'total_sequence_LH'):
            self.logger('TreeAnc.sequence_LH: you need to run marginal ancestral inference first!', 1)
            self.infer_ancestral_sequences(marginal=True)
        if pos is not None:
            cpos = self.data.full_to_compressed_sequence_map[pos] if full_sequence else pos
            return self.tree.sequence_LH[cpos]
        return self.tree.total_sequence_LH

    # -- branch profiles / lengths ---------------------------------------------------------------
    def marginal_branch_profile(self, node):
        """treeanc.py:1122-1146: (pp, pc) = (outgroup_LH, subtree_LH) of the branch above `node`."""
        if node.up is None:
            raise Exception("Branch profiles can't be calculated for the root!")
        if not self.sequence_reconstruction:
            raise Exception('marginal ancestral inference needs to be performed first!')
        return node.marginal_outgroup_LH, node.marginal_subtree_LH

    def get_branch_mutation_matrix(self, node, full_sequence=False):
        """treeanc.py:1085-1120 (host einsum on two fetched profiles; the summed statistics
        used by infer_gtr are accumulated on the device instead)."""
        pp, pc = self.marginal_branch_profile(node)
        expQt = self.gtr.expQt(self._t_last[node._fid]) + ttconf.SUPERTINY_NUMBER
        stack = np.einsum('ai,aj,ij->aij', pc, pp, expQt)
        stack = stack / stack.sum(axis=2).sum(axis=1)[:, None, None]
        return stack[self.data.full_to_compressed_sequence_map] if full_sequence else stack

    def _optimal_branch_lengths(self, fids, kinds, tol):
        """Batched GTR.optimal_t_compressed(profiles=True) (gtr.py:816-920) for many branches:
        one lock-step Brent over s = sqrt(t) with the reference's bracket and penalty."""
        eng = self._engine
        fids = np.asarray(fids, dtype=np.int32)
        kinds = np.asarray(kinds, dtype=np.int32)
        num, _ = eng.branch_hamming(fids, kinds)
        if self.comm.world_size > 1:
            num = self.comm.allreduce_sum(num)
        den = self.data.multiplicity().sum()
        hamming = 1 - num / den

        def neg_prob(idx, s):
            f = eng.branch_objective(fids[idx], s ** 2, kinds[idx])
            if self.comm.world_size > 1:
                f = self.comm.allreduce_sum(f)
            return -1.0 * f + np.exp(s ** 4 / 10000)

        n = fids.shape[0]
        smax = np.sqrt(ttconf.MAX_BRANCH_LENGTH)
        with np.errstate(invalid='ignore'):
            xb = np.sqrt(hamming)
        opt = brent_lockstep(neg_prob, np.full(n, -smax), xb, np.full(n, smax), tol=tol)
        new_len = opt['x'] ** 2
        if (new_len > 0.9 * ttconf.MAX_BRANCH_LENGTH).any():
            self.logger('WARNING: GTR.optimal_t_compressed -- The branch length seems to be very long!', 4, warn=True)
        new_len = np.where(opt['success'], new_len, hamming)           # gtr.py:916-918
        self._last_brent = opt
        return new_len

    def optimal_marginal_branch_length(self, node, tol=1e-10):
        """treeanc.py:1272-1295."""
        if node.up is None:
            return self.one_mutation
        if not self.sequence_reconstruction:
            raise Exception('marginal ancestral inference needs to be performed first!')
        return float(self._optimal_branch_lengths([node._fid], [0], tol)[0])

    def optimize_tree_marginal(self, max_iter=10, infer_gtr=False, pc=1.0, damping=0.75, LHtol=0.1,
                               site_specific_gtr=False, **kwargs):
        """treeanc.py:1297-1360 with all branches of one sweep optimised in one batched Brent."""
        self.infer_ancestral_sequences(marginal=True, **kwargs)
        oldLH = self.sequence_LH()
        self.logger('TreeAnc.optimize_tree_marginal: initial, LH=%1.2f, total branch_length %1.4f'
                    % (oldLH, self.tree.total_branch_length()), 2)
        for i in range(max_iter):
            if infer_gtr:
                self.infer_gtr(site_specific=site_specific_gtr, marginal=True, normalized_rate=True, pc=pc)
                self.infer_ancestral_sequences(marginal=True, **kwargs)
            old_bl = self.tree.total_branch_length()
            tol = 1e-8 + 0.01 ** (i + 1)
            topo = self._flat()
            root = self.tree.root
            root_bif = len(root.clades) == 2
            fids, kinds = [], []
            for n in topo.nodes[1:]:
                if n.up is root and root_bif:
                    continue
                fids.append(n._fid)
                kinds.append(0)
            if root_bif:
                fids.append(root.clades[0]._fid)
                kinds.append(1)
            new = self._optimal_branch_lengths(fids, kinds, tol)
            d = damping ** (i + 1)
            for k, fid in enumerate(fids[:len(fids) - (1 if root_bif else 0)]):
                n = topo.nodes[fid]
                n.branch_length = new[k] * (1 - d) + n.branch_length * d
                n.mutation_length = n.branch_length
            if root_bif:
                # the reference runs this block once per root child (treeanc.py:1317-1339)
                n1, n2 = root.clades
                for _ in range(2):
                    total_bl = n1.branch_length + n2.branch_length
                    bl_ratio = n1.branch_length / total_bl
                    update_val = new[-1] * (1 - d) + total_bl * d
                    n1.branch_length = update_val * bl_ratio
                    n2.branch_length = update_val * (1 - bl_ratio)
                    n1.mutation_length = n1.branch_length
                    n2.mutation_length = n2.branch_length
            self.infer_ancestral_sequences(marginal=True, **kwargs)
            LH = self.sequence_LH()
            deltaLH = LH - oldLH
            oldLH = LH
            dbl = self.tree.total_branch_length() - old_bl
            self.logger('TreeAnc.optimize_tree_marginal: iteration %d, LH=%1.2f (%1.2f), delta branch_length=%1.4f, '
                        'total branch_length %1.4f' % (i, LH, deltaLH, dbl, self.tree.total_branch_length()), 2)
            if deltaLH < LHtol:
                self.logger('TreeAnc.optimize_tree_marginal: deltaLH=%f, stopping iteration.' % deltaLH, 1)
                break
        return ttconf.SUCCESS

    def optimize_tree(self, prune_short=True, marginal_sequences=False, branch_length_mode='joint', max_iter=5,
                      infer_gtr=False, pc=1.0, method_anc='probabilistic', **kwargs):
        """treeanc.py:1384-1473; only the marginal and input modes run on the device path."""
        if branch_length_mode == 'marginal':
            self.optimize_tree_marginal(max_iter=max_iter, infer_gtr=infer_gtr, pc=pc, **kwargs)
            if prune_short:
                self.prune_short_branches()
            return ttconf.SUCCESS
        elif branch_length_mode == 'input':
            self.reconstruct_anc(method=method_anc, infer_gtr=infer_gtr, pc=pc, marginal=marginal_sequences, **kwargs)
            if prune_short:
                self.prune_short_branches()
            return ttconf.SUCCESS
        elif branch_length_mode != 'joint':
            raise UnknownMethodError("TreeAnc.optimize_tree: `branch_length_mode` should be in ['marginal', 'joint', 'input']")
        raise NotImplementedError("branch_length_mode='joint' is outside the B200 hot path; use the reference implementation")

    def optimize_branch_lengths(self, **kwargs):
        """Branch-length optimisation in marginal mode (north-star surface name)."""
        return self.optimize_tree_marginal(**kwargs)

    optimize_branch_len = optimize_branch_lengths

    def prune_short_branches(self):
        """treeanc.py:1475-1495: remove internal branches shorter than 0.1/L whose two ends carry
        identical sequences (prob_t(.., t=0) > 0.1  <=>  no non-gap mismatch)."""
        self.logger('TreeAnc.prune_short_branches: pruning short branches (max prob at zero)...', 1)
        pruned = False
        gap = self.gtr.alphabet[self.gtr.gap_index] if self.gtr.gap_index is not None else None
        for node in list(self.tree.find_clades()):
            if node.up is None or node.is_terminal():
                continue
            if node.branch_length < 0.1 * self.one_mutation:
                a, b = node.up.cseq, node.cseq
                diff = a != b
                if gap is not None:
                    diff &= (a != gap) & (b != gap)
                if not diff.any():
                    node.up.clades = [k for k in node.up.clades if k is not node] + node.clades
                    for clade in node.clades:
                        clade.up = node.up
                    pruned = True
        if pruned:
            self._topo = None

    # -- model inference -------------------------------------------------------------------------
    def infer_gtr(self, marginal=False, site_specific=False, normalized_rate=True, fixed_pi=None, pc=5.0, **kwargs):
        """treeanc.py:1500-1632, marginal branch: the n_ij / T_i accumulation over all branches
        (:1556-1572) is one device kernel; GTR.infer (gtr.py:491-599) stays on the host."""
        if site_specific:
            raise NotImplementedError('site-specific GTR inference is not provided on the device path')
        if not marginal:
            raise NotImplementedError('joint-mode GTR inference is outside the B200 hot path')
        if not self.ok:
            raise MissingDataError('TreeAnc.infer_gtr: ERROR, sequences or tree are missing')
        if self.sequence_reconstruction != 'marginal':
            self._ml_anc_marginal(**kwargs)
        n_ij, T_i = self._engine.mutation_counts()
        if self.comm.world_size > 1:
            red = self.comm.allreduce_sum(np.concatenate([n_ij.ravel(), T_i]))
            q = self.gtr.n_states
            n_ij, T_i = red[:q * q].reshape(q, q), red[q * q:]
        root_cseq = self.tree.root.cseq
        m = self.data.multiplicity()
        root_state = np.array([np.sum((root_cseq == nuc) * m) for nuc in self.gtr.alphabet])
        self._gtr = infer_gtr_from_counts(n_ij, T_i, root_state, fixed_pi=fixed_pi, pc=pc, alphabet=self.gtr.alphabet,
                                          prof_map=self.gtr.profile_map, logger=self.logger)
        if normalized_rate:
            self.logger('TreeAnc.infer_gtr: setting overall rate to 1.0...', 2)
            self._gtr.mu = 1.0
        return self._gtr

    def optimize_gtr_rate(self):
        """treeanc.py:1679-1708: Brent over sqrt(mu); each evaluation is the LH-only device pass."""
        from scipy.optimize import minimize_scalar

        def cost_func(sqrt_mu):
            self.gtr.mu = sqrt_mu ** 2
            eng = self._sync_device()
            eng.marginal(lh_only=True)
            tot, _ = eng.results()
            if self.comm.world_size > 1:
                tot = self.comm.allreduce_sum(np.array([tot]))[0]
            self.tree.total_sequence_LH = float(tot)
            return -float(tot)

        old_mu = self.gtr.mu
        try:
            sol = minimize_scalar(cost_func, bracket=[0.01 * np.sqrt(old_mu), np.sqrt(old_mu), 100 * np.sqrt(old_mu)],
                                  method='brent')
        except Exception:
            self.gtr.mu = old_mu
            self.logger('treeanc:optimize_gtr_rate: optimization failed, continuing with previous mu', 1, warn=True)
            return
        if sol['success']:
            self.gtr.mu = sol['x'] ** 2
            self.logger('treeanc:optimize_gtr_rate: optimization successful. Overall rate estimated to be %f' % self.gtr.mu, 1)
        else:
            self.gtr.mu = old_mu
            self.logger('treeanc:optimize_gtr_rate: optimization failed, continuing with previous mu', 1, warn=True)

    # -- sequences out -------------------------------------------------------------------------
    def sequence(self, node, reconstructed=False, as_string=True, compressed=False):
        """treeanc.py:1762-1809."""
        if isinstance(node, str):
            node = self.leaves_lookup[node]
        if reconstructed and not self.reconstructed_tip_sequences:
            raise ValueError('TreeAnc.sequence can only return reconstructed terminal nodes if '
                             'TreeAnc.infer_ancestral_sequences was run with this the flag `reconstruct_tip_states`.')
        if compressed:
            if (not reconstructed) and (node.name in self.data.compressed_alignment):
                tmp = self.data.compressed_alignment[node.name]
            else:
                tmp = node.cseq
        else:
            if (not reconstructed) and (node.name in self.data.aln):
                tmp = self.data.aln[node.name]
            elif node.cseq is not None:
                tmp = self.data.compressed_to_full_sequence(node.cseq, as_string=False)
            else:
                tmp = np.array([self.gtr.ambiguous or 'N'] * self.sequence_length)
        return ''.join(tmp) if as_string else np.copy(tmp)

    def get_reconstructed_alignment(self, reconstruct_tip_states=False):
        """treeanc.py:1713-1760: dict node name -> full-length sequence string."""
        if (not self.sequence_reconstruction) or (reconstruct_tip_states != self.reconstructed_tip_sequences):
            self.infer_ancestral_sequences(marginal=True, reconstruct_tip_states=reconstruct_tip_states)
        return {n.name: self.sequence(n, reconstructed=reconstruct_tip_states, as_string=True, compressed=False)
                for n in self.tree.find_clades()}
