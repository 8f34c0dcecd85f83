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

Not provided here (outside SURVEY.md §8): Fitch reconstruction, masks
(ARG mode).  They raise
NotImplementedError; the drop-in mixin for the real TreeTime
(treetime_b200.dropin) falls back to the reference's own code for them.
"""
import numpy as np

from . import config as ttconf
from .device_mixin import DeviceMarginalMixin, Unsupported, SUBTREE, OUTGROUP, PROFILE
from .gtr import GTR
from .sequence_data import SequenceData
from .tree import Node, Tree, read_newick


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


class TreeAnc(DeviceMarginalMixin):
    _missing_data_error = MissingDataError

    def __init__(self, tree=None, aln=None, gtr=None, fill_overhangs=True, ref=None, verbose=0, ignore_gaps=True,
                 convert_upper=True, seq_multiplicity=None, log=None, compress=True, seq_len=None,
                 ignore_missing_alns=False, keep_node_order=False, rng_seed=None,
                 device=0, comm=None, engine_factory=None, device_compress=False, sparse_io=False, **kwargs):
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
        self._init_device(device=device, comm=comm, engine_factory=engine_factory)
        self.sparse_io = bool(sparse_io)
        self._tree = None
        self.tree = tree
        self._gtr = None
        self.set_gtr(gtr or 'JC69', **kwargs)
        if ref is not None:
            raise NotImplementedError('sparse (VCF) alignments are read by the reference; pass a dense alignment')
        device_stats = None
        if device_compress and compress:
            # N3: pattern compression on the device -- the engine is created now, the raw alignment is
            # uploaded once and stays resident; the host only numbers the patterns
            self._engine = self._engine_factory(self.gtr.n_states, self.device)
            device_stats = self._engine.alignment_stats
        self.data = SequenceData(aln, compress=compress, convert_upper=convert_upper, fill_overhangs=fill_overhangs,
                                 ambiguous=self.gtr.ambiguous, sequence_length=seq_len, logger=self.logger,
                                 device_stats=device_stats)
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
        self._topo_dirty = True    # re-flatten at the next pass

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
        # characters in the alignment: histogram per block of rows on the host's threads (np.bincount widens its input to
        # 8-byte integers, so one call over the whole matrix allocates eight times the alignment)
        from .sequence_data import _blocks, _pool_map
        M = self.data._matrix
        hist = _pool_map(lambda b: np.bincount(M[b[0]:b[1]].ravel(), minlength=256), _blocks(M.shape[0], 256))
        present = np.flatnonzero(np.sum(hist, axis=0))
        for b in present:
            c = chr(int(b))
            if c not in self.gtr.profile_map:
                self.gtr.profile_map[c] = np.ones(self.gtr.n_states)
                self.logger('WARNING: character %s is unknown. Treating it as missing information' % c, 1, warn=True)
        self._device_patterns = False
        self.ok = True

    def reconstruct_anc(self, *args, **kwargs):
        return self.infer_ancestral_sequences(*args, **kwargs)

    def infer_ancestral_sequences(self, method='probabilistic', infer_gtr=False, marginal=False,
                                  reconstruct_tip_states=False, **kwargs):
        """treeanc.py:516-570.  Returns N_diff."""
        if not self.ok:
            raise MissingDataError('TreeAnc.infer_ancestral_sequences: ERROR, sequences or tree are missing')
        self.logger('TreeAnc.infer_ancestral_sequences with method: %s, %s' % (method, 'marginal' if marginal else 'joint'), 1)
        if method.lower() in ['ml', 'probabilistic']:
            _ml_anc = self._ml_anc_marginal if marginal else self._ml_anc_joint
        elif method.lower() in ['fitch', 'parsimony']:
            raise NotImplementedError('Fitch reconstruction is outside the B200 hot path; use the reference implementation')
        else:
            raise UnknownMethodError("Reconstruction method needs to be in ['ml', 'probabilistic', 'fitch', 'parsimony'], "
                                     "got '{}'".format(method))
        if infer_gtr:
            self.infer_gtr(marginal=marginal, **kwargs)
        return _ml_anc(reconstruct_tip_states=reconstruct_tip_states, **kwargs)

    def _branch_length_to_gtr(self, node):
        """treeanc.py:752-760."""
        if self.use_mutation_length:
            return max(ttconf.MIN_BRANCH_LENGTH * self.one_mutation, node.mutation_length)
        return max(ttconf.MIN_BRANCH_LENGTH * self.one_mutation, node.branch_length)

    def optimize_tree(self, prune_short=True, marginal_sequences=False, branch_length_mode='joint', max_iter=5,
                      infer_gtr=False, pc=1.0, method_anc='probabilistic', **kwargs):
        """treeanc.py:1384-1473."""
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
        # joint mode (treeanc.py:1449-1473): reconstruct, optimise every branch on the device's pair counts, repeat
        self.logger('TreeAnc.optimize_tree: sequences...', 1)
        self.reconstruct_anc(method=method_anc, infer_gtr=infer_gtr, pc=pc, marginal=marginal_sequences, **kwargs)
        self.optimize_branch_lengths_joint(store_old=False)
        n = 0
        while n < max_iter:
            n += 1
            if prune_short:
                self.prune_short_branches()
            N_diff = self.reconstruct_anc(method=method_anc, infer_gtr=False, marginal=marginal_sequences, **kwargs)
            self.logger('TreeAnc.optimize_tree: Iteration %d. #Nuc changed since prev reconstructions: %d' % (n, N_diff), 2)
            if N_diff < 1:
                break
            self.optimize_branch_lengths_joint(store_old=False)
        self.tree.unconstrained_sequence_LH = (self.tree.sequence_LH * self.data.multiplicity()).sum()
        self._prepare_nodes()
        self.logger('TreeAnc.optimize_tree: Unconstrained sequence LH:%f' % self.tree.unconstrained_sequence_LH, 2)
        return ttconf.SUCCESS

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
            self._topo_dirty = True

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
