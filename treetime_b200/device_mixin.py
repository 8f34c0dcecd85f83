"""Device-side implementation of the marginal path, shared by the standalone TreeAnc mirror
(treetime_b200.treeanc.TreeAnc) and the drop-in mixin for the real TreeTime
(treetime_b200.dropin).  It only relies on the TreeAnc attribute surface both have:
`tree`, `data` (compressed_alignment, multiplicity(), compressed_length, full_length), `gtr`,
`_branch_length_to_gtr`, `logger`, `rng`, `one_mutation`, `sequence_reconstruction`,
`reconstructed_tip_sequences`.
"""
import operator
import os

import numpy as np

from . import _fastscan
from . import config as ttconf
from .brent import brent_lockstep
from .dist import default_comm, shard_bounds
from .flatten import FlatTopology, code_table, encode_chars, gtr_arrays
from .gtr import infer_gtr_from_counts
from .pairs import NONE as PAIR_NONE, count_matrices, fold_state_pairs, optimal_t_from_counts
from .seq_utils import prof2seq

SUBTREE, OUTGROUP, PROFILE = 0, 1, 2


class Unsupported(NotImplementedError):
    """Raised for inputs the device path does not cover (masks, full sampling, ...)."""


def _default_engine_factory(n_states, device):
    from .engine import Engine          # raises if libttb.so is missing or there is no GPU
    return Engine(n_states, device=device)


class DeviceMarginalMixin(object):
    def _init_device(self, device=0, comm=None, engine_factory=None):
        self.device = device
        self.comm = comm if comm is not None else default_comm()
        self._engine_factory = engine_factory or _default_engine_factory
        self._engine = None
        self._topo = None
        self._device_topology = None
        self._device_patterns = False
        self._device_data_id = None
        self._device_masks = None
        self._cache = {}
        self._seq_cache = {}
        self._stale_states = None
        self._device_tips = None
        self._tips_of_topo = (None, None)
        # sparse host interface: the tip codes cross PCIe as (reference row + differences) -- the analogue of the
        # reference's dict-of-differences alignments (sequence_data.py:363-383) -- and sequence_differences() returns
        # the reconstruction as (root row + states differing from the parent), the content of node.mutations
        self.sparse_io = False
        self._sparse_codes = None

    def _unsupported(self, why):
        raise Unsupported(why)

    _missing_data_error = RuntimeError      # overridden with the host package's MissingDataError

    # -- device synchronisation -------------------------------------------------------------
    def _flat(self):
        """The flattening the DEVICE currently knows (node._fid = device node id).  It is rebuilt only
        at the start of a pass (_refresh_topology): between passes the per-node results on the device
        stay addressable even if the host tree was edited, exactly like the reference's stale per-node
        attributes after prune_short_branches / polytomy resolution."""
        if self._topo is None:
            self._refresh_topology()
        return self._topo

    def _topology_changed(self):
        return self._topo is None or getattr(self, '_topo_dirty', False)

    def _snapshot_states(self):
        """State indices of every reconstructed node of the CURRENT device topology, keyed by node object:
        {id(node): uint8[L' shard]}.  Taken just before the flattening is rebuilt: ttb_set_tree drops the
        device's previous states, but the reference compares every surviving node with its own old cseq
        (treeanc.py:925-926) -- so after prune_short_branches / resolve_polytomies / reroot N_diff is counted
        on the host against this snapshot.  None when there is nothing to compare with."""
        eng, topo = self._engine, self._topo
        if eng is None or topo is None or not self.sequence_reconstruction or not getattr(self, '_b200_live', True):
            return None
        try:
            rows = eng.all_seq_idx()
            old = {id(topo.nodes[n]): rows[k] for k, n in enumerate(topo.internal_nodes)}
            if self.reconstructed_tip_sequences:
                tips = eng.seq_idx(topo.tip_nodes)
                old.update({id(topo.nodes[n]): tips[k] for k, n in enumerate(topo.tip_nodes)})
        except RuntimeError:            # TTBError: the engine holds no reconstruction
            return None
        return old

    def _refresh_topology(self):
        if self._topology_changed():
            self._stale_states = self._snapshot_states()
            self._topo = FlatTopology(self.tree.root)
            for i, n in enumerate(self._topo.nodes):
                n._fid = i
            self._topo_dirty = False
            self._cache = {}
            self._seq_cache = {}
        return self._topo

    def _shard(self):
        return shard_bounds(self.data.compressed_length, self.comm.rank, self.comm.world_size)

    def _tip_codes(self):
        """uint8 codes [n_tips, L'] + (n_codes, q) table.  Works on treetime_b200.SequenceData
        (compressed ASCII matrix, vectorised) and on the reference's SequenceData (dict name ->
        numpy char array)."""
        topo = self._flat()
        chars, lut, table = code_table(self.gtr.profile_map, self.gtr.n_states)
        lo, hi = self._shard()
        if hasattr(self.data, 'compressed_matrix'):
            codes = np.empty((topo.n_tips, hi - lo), dtype=np.uint8)
            lut8 = np.full(256, 255, dtype=np.uint8)
            for c, i in lut.items():
                lut8[ord(c)] = i
            rows = np.array([self.data._row.get(topo.nodes[n].name, -1) for n in topo.tip_nodes])
            have = np.flatnonzero(rows >= 0)
            cm = self.data.compressed_matrix
            from .sequence_data import _blocks, _pool_map

            def encode(b):          # row blocks on the host's threads: gather the tips' rows, map characters to codes
                k = have[b[0]:b[1]]
                codes[k] = np.take(lut8, cm[rows[k], lo:hi])
            _pool_map(encode, _blocks(have.shape[0], 256))
            codes[rows < 0] = len(chars)                                      # tips without sequence: missing = all ones
        else:
            codes = np.full((topo.n_tips, hi - lo), len(chars), dtype=np.uint8)
            ca = self.data.compressed_alignment
            for n in topo.tip_nodes:
                name = topo.nodes[n].name
                if name in ca:
                    codes[topo.tip_row[n]] = encode_chars(np.asarray(ca[name])[lo:hi], chars)
        if codes.size and codes.max() == 255:             # one pass, no temporary
            raise KeyError('alignment contains characters that are not in the profile map')
        return codes, table

    def _sync_device(self):
        """Bring the engine up to date with tree topology, patterns, model and branch lengths."""
        topo = self._refresh_topology()
        if self._engine is None:
            self._engine = self._engine_factory(self.gtr.n_states, self.device)
        eng = self._engine
        sig = topo.signature()
        if sig != self._device_topology:
            eng.set_tree(topo.parent, topo.child_ptr, topo.child_idx, topo.tip_row)
            self._device_topology = sig
            self._device_patterns = False
            self._device_masks = None
        data_id = (id(self.data), self.data.compressed_length, len(self.gtr.profile_map))
        if data_id != self._device_data_id:
            self._device_patterns = False
        # the same shape with the tips in other positions (swapped leaves, reroot + ladderize of a symmetric tree,
        # `tt.tree = other`) leaves parent / child_idx unchanged: the row -> tip assignment is part of the identity
        if self._tips_of_topo[0] is not topo:                  # once per flattening, not per pass
            self._tips_of_topo = (topo, hash(tuple(topo.nodes[n].name for n in topo.tip_nodes)))
        tips = self._tips_of_topo[1]
        if tips != self._device_tips:
            self._device_patterns = False
            self._device_tips = tips
        if not self._device_patterns:
            self._device_data_id = data_id
            lo, hi = self._shard()
            if getattr(self.data, 'device_resident', False):
                # the raw alignment already lives on the device (N3): gather the code matrix there
                chars, lut, table = code_table(self.gtr.profile_map, self.gtr.n_states)
                lut8 = np.full(256, 255, dtype=np.uint8)
                for c, i in lut.items():
                    lut8[ord(c)] = i
                rows = np.array([self.data._row.get(topo.nodes[n].name, -1) for n in topo.tip_nodes], dtype=np.int32)
                eng.set_patterns_from_alignment(self.data.pattern_first_position[lo:hi], self.data.pattern_const_letter[lo:hi],
                                                rows, lut8, len(chars), table, self.data.multiplicity()[lo:hi])
            elif self.sparse_io:
                key = (self._device_data_id, self._device_tips, lo, hi)
                if self._sparse_codes is None or self._sparse_codes[0] != key:
                    from .sparse import sparse_from_dense
                    codes, table = self._tip_codes()
                    self._sparse_codes = (key, sparse_from_dense(codes), table, np.ascontiguousarray(self.data.multiplicity()[lo:hi]))
                _, (ref, e_row, e_pos, e_code), table, mult = self._sparse_codes
                eng.set_patterns_sparse(ref, e_row, e_pos, e_code, table, mult)
            else:
                codes, table = self._tip_codes()
                eng.set_patterns(codes, table, self.data.multiplicity()[lo:hi], validate=False)   # codes built by _tip_codes
            self._device_patterns = True
            self._device_masks = None
        scan = self._scan_nodes(topo.nodes)           # branch lengths + mask presence in one walk over the node objects
        self._sync_masks(eng, topo, masked=None if scan is None else scan[1])
        g = gtr_arrays(self.gtr)
        upload_model = True
        if g['site_specific']:
            lo, hi = self._shard()
            g = dict(g, eigenvals=g['eigenvals'][:, lo:hi], v=g['v'][:, :, lo:hi], v_inv=g['v_inv'][:, :, lo:hi],
                     Pi=g['Pi'][:, lo:hi], mu=g['mu'][lo:hi])
            # the per-pattern model is megabytes: re-upload only when it changed
            fp = (id(self.gtr), float(np.sum(g['mu'])), float(np.sum(g['eigenvals'])), float(np.sum(g['Pi'][0])), lo, hi,
                  bool(g['approximate']), self._device_data_id)
            upload_model = fp != getattr(self, '_device_model_fp', None)
            self._device_model_fp = fp
        else:
            self._device_model_fp = None
        tvec = self._branch_lengths_to_gtr(topo.nodes, scanned=None if scan is None else scan[0])
        lam = np.max(g['eigenvals']) * np.max(g['mu'])
        if lam * tvec[1:].max() > 10:
            raise ValueError('Error in computing exp(Q * t): Q has positive eigenvalues or the branch length t is too large. '
                             'This is most likely caused by incorrect input data.')     # gtr.py:1041-1047
        if upload_model:
            eng.set_gtr(g)
        eng.set_branch_lengths(tvec)
        self._t_last = tvec
        return eng

    def _scan_nodes(self, nodes):
        """(raw branch / mutation lengths with entry 0 unset, any node has a mask) from ONE cache-friendly C walk over the
        nodes' dicts (csrc/ttb_fastscan.c: 200 000 nodes 45 -> ~10 ms against two separate walks), or None when the helper is
        missing, the per-node method is overridden or an attribute is not a plain number -- the callers then scan themselves."""
        own = getattr(type(self)._branch_length_to_gtr, '__qualname__', '').split('.')[0] in ('TreeAnc',)
        if not own:
            return None
        attr = 'mutation_length' if self.use_mutation_length else 'branch_length'
        vals = np.empty(len(nodes), dtype=np.float64)
        masked = _fastscan.scan_nodes(self._node_dicts(nodes), attr, 'mask', vals, 1)
        return None if masked is None else (vals, masked)

    def _branch_lengths_to_gtr(self, nodes, scanned=None):
        """_branch_length_to_gtr (treeanc.py:752-760) for all nodes at once: max(MIN_BRANCH_LENGTH * one_mutation,
        branch or mutation length).  One Python call per node costs more than the device pass on large trees (40 000
        nodes: 20 ms vs 14 ms), so the floor is applied vectorised; a subclass that overrides the per-node method is
        still honoured."""
        own = getattr(type(self)._branch_length_to_gtr, '__qualname__', '').split('.')[0] in ('TreeAnc',)
        if not own:
            return np.array([self._branch_length_to_gtr(n) for n in nodes], dtype=np.float64)
        attr = 'mutation_length' if self.use_mutation_length else 'branch_length'
        floor = ttconf.MIN_BRANCH_LENGTH * self.one_mutation
        n = len(nodes)
        vals = scanned if scanned is not None else np.empty(n, dtype=np.float64)
        try:        # instance attributes straight from the nodes' dicts, no list of boxed floats in between
            dicts = self._node_dicts(nodes)
            if scanned is None and not _fastscan.scan_float_attr(dicts, attr, vals, 1):        # C loop over the dicts (3.4 -> 0.5 ms at 40 000 nodes)
                vals[1:] = np.fromiter(map(operator.itemgetter(attr), dicts[1:]), dtype=np.float64, count=n - 1)
            root_val = nodes[0].__dict__.get(attr)
        except (KeyError, TypeError, ValueError):           # properties, slots, None on a non-root node
            vals = np.array(list(map(operator.attrgetter(attr), nodes)), dtype=np.float64)      # None -> nan
            root_val = vals[0]
        vals[0] = floor if (root_val is None or not np.isfinite(root_val)) else root_val
        return np.maximum(floor, vals)

    def _node_dicts(self, nodes):
        """The nodes' instance dicts, cached per flattening (the per-pass scans read attributes from them at C speed)."""
        cached = getattr(self, '_dicts_of', None)
        if cached is None or cached[0] is not nodes:
            cached = self._dicts_of = (nodes, [n.__dict__ for n in nodes])
        return cached[1]

    def _has_masks(self):
        return any(getattr(n, 'mask', None) is not None for n in self.tree.find_clades())

    def _mask_problem(self):
        """None if every node.mask has a device form (0/1 over the patterns), else the reason."""
        L = self.data.compressed_length
        seen = set()
        for n in self.tree.find_clades():
            m = getattr(n, 'mask', None)
            if m is None or id(m) in seen:
                continue
            seen.add(id(m))
            arr = np.asarray(m, dtype=float)
            if arr.shape != (L,):
                return 'a branch mask must have one entry per alignment pattern'
            if not np.all((arr == 0) | (arr == 1)):
                return 'fractional branch masks are not supported on the device path'
        return None

    def _sync_masks(self, eng, topo, masked=None):
        """Per-branch masks (node.mask, set by arg.py:128-133): distinct 0/1 vectors over the patterns + one index per
        node.  Fractional masks have no device form (a masked message is dropped, not scaled)."""
        from itertools import repeat
        dicts = self._node_dicts(topo.nodes)
        # one lazy C-speed pass in the common case (no node has a mask); the list is only built otherwise
        if masked is None:
            masked = _fastscan.any_not_none(dicts, 'mask')
        if masked is None:
            masked = any(map(operator.is_not, map(dict.get, dicts, repeat('mask')), repeat(None)))     # identity, not ==: masks are arrays
        node_masks = list(map(dict.get, dicts, repeat('mask'))) if masked else None
        if not masked:
            if self._device_masks is not None:
                eng.set_branch_masks(None, None)
                self._device_masks = None
            return
        L = self.data.compressed_length
        rows, index, node_mask = [], {}, np.full(topo.n_nodes, -1, dtype=np.int32)
        for i, m in enumerate(node_masks):
            if m is None:
                continue
            arr = np.asarray(m, dtype=float)
            if arr.shape != (L,):
                self._unsupported('a branch mask must have one entry per alignment pattern')
            if not np.all((arr == 0) | (arr == 1)):
                self._unsupported('fractional branch masks are not supported on the device path')
            key = arr.tobytes()
            if key not in index:
                index[key] = len(rows)
                rows.append(arr.astype(np.uint8))
            node_mask[i] = index[key]
        sig = (node_mask.tobytes(), tuple(index))
        if sig != self._device_masks:
            lo, hi = self._shard()
            eng.set_branch_masks(np.array(rows)[:, lo:hi], node_mask)
            self._device_masks = sig

    def _branch_mask(self, node, kind=0):
        """The mask data.multiplicity() gets for the branch above `node` (treeanc.py:1294) or, kind 1, for the merged
        branch across a bifurcating root (:1326-1333); None = no mask."""
        if kind == 1:
            n1, n2 = self.tree.root.clades
            m1, m2 = getattr(n1, 'mask', None), getattr(n2, 'mask', None)
            return None if m1 is None or m2 is None else np.asarray(m1) * np.asarray(m2)
        return getattr(node, 'mask', None)

    def _shard_sizes(self):
        """Pattern count of every rank's shard (shard_bounds is deterministic: no need to exchange them)."""
        L, W = self.data.compressed_length, self.comm.world_size
        return [hi - lo for lo, hi in (shard_bounds(L, r, W) for r in range(W))]

    def _gather_patterns(self, x, axis=0):
        if self.comm.world_size == 1:
            return x
        sizes = self._shard_sizes()
        return self.comm.allgather(x, axis=axis, sizes=sizes if np.shape(x)[axis] == sizes[self.comm.rank] else None)

    def _gather_pass_results(self, site_lh, tot, nd):
        """tree.sequence_LH of all shards plus the summed {total LH, N_diff} of a pass in ONE collective: every rank appends
        its two scalars to its per-pattern likelihoods (the scalars are then added up in rank order on every rank: same
        bits everywhere)."""
        if self.comm.world_size == 1:
            return site_lh, tot, nd
        if os.environ.get('TTB_TWO_COLLECTIVES'):      # measurement only: the earlier form (all-reduce + gather with size exchange)
            tot, nd = self.comm.allreduce_sum(np.array([tot, float(nd)]))
            return self.comm.allgather(site_lh), tot, nd
        sizes = [k + 2 for k in self._shard_sizes()]
        g = self.comm.allgather(np.concatenate([np.asarray(site_lh, dtype=np.float64), [float(tot), float(nd)]]), sizes=sizes)
        parts, off, tot, nd = [], 0, 0.0, 0.0
        for k in sizes:
            parts.append(g[off:off + k - 2])
            tot += g[off + k - 2]
            nd += g[off + k - 1]
            off += k
        return np.concatenate(parts), tot, nd

    def _node_array(self, node, which):
        key = (node._fid, which)
        if key not in self._cache:
            if self.sequence_reconstruction != 'marginal' and not (which == SUBTREE and self._engine is not None
                                                                   and not self.sequence_reconstruction):
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

    def reload_alignment(self):
        """Mark the device copy of the alignment stale: the next pass uploads the tip codes again (dense, or as
        reference row + differences with sparse_io).  The reference has no such call -- its alignment lives in host
        memory -- this is the host->device half of a pass for callers that stream alignments through one object."""
        self._device_patterns = False

    def sequence_differences(self, gather=True):
        """Every reconstructed internal sequence in sparse form: (root state indices [L'], node, pos, state) with one
        entry per (internal node, compressed position) whose state differs from the parent's, sorted by (node, pos);
        `node` indexes tree.find_clades() order.  This is what `node.mutations` enumerates (treeanc.py:27-42), for the
        whole tree in one device call instead of one dense sequence per node."""
        if not self.sequence_reconstruction:
            raise ValueError('Ancestral sequences are not yet inferred')
        root, node, pos, state = self._engine.mutations()
        if gather and self.comm.world_size > 1:       # gather=False: this rank's pattern shard only, shard-local positions
            lo, _ = self._shard()
            root = self.comm.allgather(root, axis=0)
            node, pos, state = (self.comm.allgather(x, axis=0) for x in (node, pos + lo, state))
            order = np.lexsort((pos, node))
            node, pos, state = node[order], pos[order], state[order]
        return root, node, pos, state

    # -- ancestral reconstruction ---------------------------------------------------------
    def _ml_anc_marginal(self, sample_from_profile=False, reconstruct_tip_states=False, debug=False, **kwargs):
        """treeanc.py:762-812: postorder, root, preorder -- one graph launch on the device."""
        self.logger('TreeAnc._ml_anc_marginal: type of reconstruction: Marginal', 2)
        if sample_from_profile == 'root':
            root_sample, other_sample = True, False
        elif isinstance(sample_from_profile, bool):
            root_sample = other_sample = sample_from_profile
        else:
            raise ValueError("sample_from_profile must be a bool or 'root'")
        eng = self._sync_device()
        topo = self._flat()
        stale, self._stale_states = self._stale_states, None
        eng.marginal(reconstruct_tips=reconstruct_tip_states, keep_prev=other_sample)
        tot, nd = eng.results()
        self._cache = {}
        self._seq_cache = {}
        self.tree.sequence_LH, tot, nd = self._gather_pass_results(eng.site_lh(), tot, nd)
        self.tree.total_sequence_LH = float(tot)
        self.tree.sequence_marginal_LH = self.tree.total_sequence_LH
        had_reconstruction, prev_tips = self.sequence_reconstruction, self.reconstructed_tip_sequences
        root = self.tree.root
        root._cseq_override = None
        self.reconstructed_tip_sequences = reconstruct_tip_states
        self.sequence_reconstruction = 'marginal'
        if root_sample:                                                 # treeanc.py:831-838, host RNG
            seq, _, _ = prof2seq(self._node_array(root, PROFILE), self.gtr, sample_from_prof=True, normalize=False, rng=self.rng)
            root._cseq_override = seq
        nd_tips = None
        if other_sample:
            nd, nd_tips = self._sample_states(eng, topo, reconstruct_tip_states)
        N_diff = self._n_diff(eng, topo, nd, reconstruct_tip_states, prev_tips, had_reconstruction, nd_tips, stale=stale)
        self.logger('TreeAnc._ml_anc_marginal: ...done', 3)
        return N_diff

    def _sample_states(self, eng, topo, reconstruct_tip_states):
        """sample_from_profile=True (treeanc.py:919-923): every reconstructed non-root node draws its sequence from
        its marginal profile.  The uniforms come from the caller's generator in the reference's order -- one
        rng.random(L') per node in preorder (seq_utils.py:268), after the root's draw -- and go to the device in
        blocks; a block of k rows of rng.random((k, L')) is the same stream as k successive rng.random(L') calls."""
        L = self.data.compressed_length
        lo, hi = self._shard()
        ids = [n._fid for n in topo.nodes[1:] if reconstruct_tip_states or not n.is_terminal()]
        blk = max(1, (1 << 24) // max(L, 1))                             # <= 128 MB of uniforms per block
        nd = nd_tips = 0
        for b in range(0, len(ids), blk):
            part = ids[b:b + blk]
            u = self.rng.random(size=(len(part), L))
            a, t = eng.sample_states(part, u[:, lo:hi])
            nd += a
            nd_tips += t
        if self.comm.world_size > 1:
            nd, nd_tips = (int(round(x)) for x in self.comm.allreduce_sum(np.array([float(nd), float(nd_tips)])))
        return nd + nd_tips, nd_tips

    def _n_diff(self, eng, topo, nd, reconstruct_tip_states, prev_tips, had_reconstruction=Ellipsis, nd_tips=None, stale=None):
        """N_diff of a pass (treeanc.py:925-928 / 1042-1045).  The device counts changed states against
        the previous device states; host-side corrections reproduce the reference:
          * no previous reconstruction -> every reconstructed position counts;
          * the topology was rebuilt since the previous pass (`stale` = _snapshot_states() of the old numbering)
            -> every surviving node is compared with its own old states, new nodes count fully;
          * tips reconstructed now but not before -> the reference compares them with the alignment's own
            (possibly ambiguous) characters, the device compared them with stale / unset states."""
        L = self.data.compressed_length
        if had_reconstruction is Ellipsis:
            had_reconstruction = self.sequence_reconstruction
        if not had_reconstruction:
            n_rec = (topo.n_nodes - 1) if reconstruct_tip_states else (topo.n_nodes - topo.n_tips - 1)
            return n_rec * L
        if stale is not None:
            lo, hi = self._shard()
            rows = eng.all_seq_idx()
            nd = 0
            for k, n in enumerate(topo.internal_nodes):
                if n:                                                # the root's sequence is not compared
                    old = stale.get(id(topo.nodes[n]))
                    nd += (hi - lo) if old is None else int((rows[k] != old).sum())
            nd_tips = 0
            if reconstruct_tip_states:
                for b in range(0, topo.n_tips, 256):
                    blk = topo.tip_nodes[b:b + 256]
                    idx = eng.seq_idx(blk)
                    for k, n in enumerate(blk):
                        old = stale.get(id(topo.nodes[n]))
                        nd_tips += (hi - lo) if old is None else int((idx[k] != old).sum())
            if self.comm.world_size > 1:
                nd, nd_tips = (int(round(x)) for x in self.comm.allreduce_sum(np.array([float(nd), float(nd_tips)])))
            nd += nd_tips
        nd = int(round(nd))
        if reconstruct_tip_states and not prev_tips:
            if nd_tips is None:
                nd_tips = eng.results_tips()
                if self.comm.world_size > 1:
                    nd_tips = int(round(self.comm.allreduce_sum(np.array([float(nd_tips)]))[0]))
            fresh = 0
            ca = self.data.compressed_alignment
            tips = list(topo.tip_nodes)
            for b in range(0, len(tips), 256):                        # batched D2H, one transition per run
                blk = tips[b:b + 256]
                idx = self._gather_patterns(eng.seq_idx(blk), axis=1)
                for k, n in enumerate(blk):
                    name = topo.nodes[n].name
                    if name in ca:
                        fresh += int((self.gtr.alphabet[idx[k]] != np.asarray(ca[name])).sum())
                    else:
                        fresh += L      # a tip without sequence has no cseq: the reference's comparison is all-True
            nd = nd - nd_tips + fresh
        return nd

    def _ml_anc_joint(self, sample_from_profile=False, reconstruct_tip_states=False, debug=False, **kwargs):
        """treeanc.py:934-1080 (N2): joint ML reconstruction -- log-space postorder with back-pointers,
        root state (argmax, or sampled on the host with the caller's RNG), back-trace."""
        self.logger('TreeAnc._ml_anc_joint: type of reconstruction: Joint', 2)
        if sample_from_profile == 'root':
            root_sample = True
        elif isinstance(sample_from_profile, bool):
            root_sample = sample_from_profile           # treeanc.py:1008-1011: a bool only affects the root
        else:
            raise ValueError("sample_from_profile must be a bool or 'root'")
        if any(getattr(n, 'mask', None) is not None for n in self.tree.find_clades()):
            self._unsupported('joint reconstruction with per-branch masks (ARG mode) runs in the reference')
        if getattr(self.gtr, 'is_site_specific', False):
            self._unsupported('joint reconstruction with site-specific models runs in the reference')
        eng = self._sync_device()
        topo = self._flat()
        stale, self._stale_states = self._stale_states, None
        if root_sample:
            eng.joint(reconstruct_tips=reconstruct_tip_states, trace=False)
            eng.results()
            lx = self._gather_patterns(eng.node_array(0, 3), axis=0)              # root.joint_Lx
            normalized = (lx.T - lx.max(axis=1)).T
            seq, _, idxs = prof2seq(np.exp(normalized), self.gtr, sample_from_prof=True, rng=self.rng)
            lo, hi = self._shard()
            eng.joint_retrace(idxs[lo:hi].astype(np.uint8), reconstruct_tips=reconstruct_tip_states)
        else:
            eng.joint(reconstruct_tips=reconstruct_tip_states)
        tot, nd = eng.results()
        self._cache = {}
        self._seq_cache = {}
        self.tree.sequence_LH, tot, nd = self._gather_pass_results(eng.site_lh(), tot, nd)
        self.tree.sequence_joint_LH = float(tot)
        N_diff = self._n_diff(eng, topo, nd, reconstruct_tip_states, self.reconstructed_tip_sequences, stale=stale)
        self.tree.root._cseq_override = None
        self.reconstructed_tip_sequences = reconstruct_tip_states
        self.sequence_reconstruction = 'joint'
        if debug:       # the reference keeps joint_Lx / joint_Cx of every node in debug mode; here only the root's is resident
            self.tree.root.joint_Lx = self._gather_patterns(eng.node_array(0, 3), axis=0)
        elif 'joint_Lx' in self.tree.root.__dict__:
            del self.tree.root.joint_Lx
        self.logger('TreeAnc._ml_anc_joint: ...done', 3)
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
        if np.ndim(expQt) == 3:                                                  # site-specific: treeanc.py:1107-1108
            stack = np.einsum('ai,aj,ija->aij', pc, pp, expQt)
        else:
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
        if self._device_masks is not None:          # data.multiplicity(mask=...) per branch
            topo = self._flat()
            den = np.array([self.data.multiplicity(mask=self._branch_mask(topo.nodes[f], k)).sum() for f, k in zip(fids, kinds)])
        hamming = 1 - num / den

        n = fids.shape[0]
        smax = np.sqrt(ttconf.MAX_BRANCH_LENGTH)
        with np.errstate(invalid='ignore'):
            xb = np.sqrt(hamming)
        reduce_dev = None
        on_device = hasattr(eng, 'brent_minimize')
        if on_device and self.comm.world_size > 1:
            reduce_dev = getattr(self.comm, 'allreduce_sum_device', lambda e: None)(eng)
            on_device = reduce_dev is not None
        if on_device:
            # the whole state machine on the device: no per-iteration host round trip (ttb_brent_*)
            opt = eng.brent_minimize(fids, kinds, np.full(n, -smax), xb, np.full(n, smax), tol, allreduce=reduce_dev)
        else:
            def neg_prob(idx, s):
                f = eng.branch_objective(fids[idx], s ** 2, kinds[idx])
                if self.comm.world_size > 1:
                    f = self.comm.allreduce_sum(f)
                return -1.0 * f + np.exp(s ** 4 / 10000)

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
        last_tbl = 0.0
        if self.verbose > 2:
            last_tbl = self.tree.total_branch_length()
            self.logger('TreeAnc.optimize_tree_marginal: initial, LH=%1.2f, total branch_length %1.4f' % (oldLH, last_tbl), 2)
        for i in range(max_iter):
            if infer_gtr:
                self.infer_gtr(site_specific=site_specific_gtr, marginal=True, normalized_rate=True, pc=pc)
                self.infer_ancestral_sequences(marginal=True, **kwargs)
            tol = 1e-8 + 0.01 ** (i + 1)
            topo = self._flat()
            root = self.tree.root
            root_bif = len(root.clades) == 2
            # every branch but the two below a bifurcating root, which are optimised as one merged branch (kind 1)
            ids = np.arange(1, topo.n_nodes, dtype=np.int32)
            if root_bif:
                ids = ids[topo.parent[1:] != 0]
                fids = np.concatenate([ids, np.array([root.clades[0]._fid], dtype=np.int32)])
                kinds = np.zeros(fids.shape[0], dtype=np.int32)
                kinds[-1] = 1
            else:
                fids, kinds = ids, np.zeros(ids.shape[0], dtype=np.int32)
            new = self._optimal_branch_lengths(fids, kinds, tol)
            d = damping ** (i + 1)
            # damped update (treeanc.py:1340-1343) for all branches at once; the lengths are read and written through the
            # nodes' instance dicts (one C-level pass each) instead of three attribute operations per node in Python
            dicts = self._node_dicts(topo.nodes)
            try:
                cur = np.fromiter(map(operator.itemgetter('branch_length'), map(dicts.__getitem__, ids)), dtype=np.float64, count=ids.shape[0])
                upd = (new[:ids.shape[0]] * (1 - d) + cur * d).tolist()
                for k, v in zip(ids.tolist(), upd):
                    dk = dicts[k]
                    dk['branch_length'] = v
                    dk['mutation_length'] = v
            except (KeyError, TypeError, ValueError):        # branch_length is not a plain instance attribute
                for k, fid in enumerate(ids.tolist()):
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
            if self.verbose > 2:            # two tree walks per iteration only when somebody reads the message
                tbl = self.tree.total_branch_length()
                self.logger('TreeAnc.optimize_tree_marginal: iteration %d, LH=%1.2f (%1.2f), delta branch_length=%1.4f, '
                            'total branch_length %1.4f' % (i, LH, deltaLH, tbl - last_tbl, tbl), 2)
                last_tbl = tbl
            if deltaLH < LHtol:
                self.logger('TreeAnc.optimize_tree_marginal: deltaLH=%f, stopping iteration.' % deltaLH, 1)
                break
        return ttconf.SUCCESS

    # -- joint branch lengths (N2) ---------------------------------------------------------------
    def _pair_tables(self):
        """Parent/child pair counts of every branch from the device (ttb_branch_state_pairs), cached until
        the next reconstruction.  Returns (C[n_nodes-1, q, W], F, tip_states, code characters)."""
        if 'pairs' not in self._cache:
            if not self.sequence_reconstruction or self._engine is None:
                raise Exception('ancestral sequences need to be reconstructed first!')
            if self._has_masks():
                self._unsupported('pair counts under per-branch masks (ARG mode) run in the reference')
            eng, topo = self._engine, self._flat()
            tip_states = bool(self.reconstructed_tip_sequences)
            C, F = eng.branch_state_pairs(np.arange(1, topo.n_nodes, dtype=np.int32), tip_states=tip_states)
            if self.comm.world_size > 1:
                lo, _ = self._shard()
                F = np.where(F != PAIR_NONE, F + lo, F)              # global pattern index of the first occurrence
                C = self.comm.allreduce_sum(C)
                F = self.comm.allgather(F[None], axis=0).min(axis=0)
            chars, _, _ = code_table(self.gtr.profile_map, self.gtr.n_states)
            self._cache['pairs'] = (C, F, tip_states, list(chars) + [None])   # last code = tip without sequence
        return self._cache['pairs']

    def _branch_state(self, node):
        """The reference's node.branch_state = {'pair', 'multiplicity'} (treeanc.py:1148-1163) from device counts."""
        C, F, tip_states, code_chars = self._pair_tables()
        alphabet = [str(c) for c in self.gtr.alphabet]
        if node.is_terminal() and not tip_states:
            if node.name not in self.data.compressed_alignment:
                raise self._missing_data_error(
                    "TreeAnc.optimal_branch_length: terminal node alignments required; sequence is missing for leaf: '%s'. "
                    'Missing terminal sequences can be inferred from sister nodes by rerunning with `reconstruct_tip_states=True` '
                    'or `--reconstruct-tip-states`' % node.name)
            cols = code_chars
        else:
            cols = alphabet
        pairs, mult = fold_state_pairs(C[node._fid - 1], F[node._fid - 1], alphabet, cols, self.gtr.gap_index, self.ignore_gaps)
        return {'pair': pairs, 'multiplicity': mult}

    def add_branch_state(self, node):
        """treeanc.py:1148-1163."""
        node.branch_state = self._branch_state(node)

    def optimal_branch_length(self, node):
        """treeanc.py:1245-1270: optimal length of the branch above `node` given the reconstructed sequences."""
        if node.up is None:
            return self.one_mutation
        bs = self._branch_state(node)
        q = self.gtr.n_states
        M = np.zeros((1, q, q))
        M[0, bs['pair'][:, 0], bs['pair'][:, 1]] = bs['multiplicity']
        return float(optimal_t_from_counts(self.gtr, M)[0][0])

    def optimize_branch_lengths_joint(self, **kwargs):
        """treeanc.py:1176-1243 with every branch optimised in one lock-step Brent on the device's pair counts."""
        self.logger('TreeAnc.optimize_branch_length: running branch length optimization using jointML ancestral sequences', 1)
        if getattr(self.gtr, 'is_site_specific', False):
            self._unsupported('joint branch-length optimisation with site-specific models runs in the reference')
        store_old = kwargs.get('store_old', False)
        topo = self._flat()
        C, F, tip_states, code_chars = self._pair_tables()
        alphabet = [str(c) for c in self.gtr.alphabet]
        is_tip = np.array([n.is_terminal() for n in topo.nodes[1:]])
        if not tip_states:
            for n in topo.nodes[1:]:
                if n.is_terminal() and n.name not in self.data.compressed_alignment:
                    self._branch_state(n)                                   # raises the reference's MissingDataError
        M = count_matrices(C[:, :, :len(alphabet)], alphabet, alphabet, self.gtr.gap_index, self.ignore_gaps)
        if not tip_states and is_tip.any():
            M[is_tip] = count_matrices(C[is_tip], alphabet, code_chars, self.gtr.gap_index, self.ignore_gaps)
        new, self._last_brent = optimal_t_from_counts(self.gtr, M)
        max_bl = 0
        for n, t in zip(topo.nodes[1:], new):
            if store_old:
                n._old_length = n.branch_length
            t = max(0, float(t))
            n.branch_length = t
            n.mutation_length = t
            max_bl = max(max_bl, t)
        if max_bl > 0.15:
            self.logger("TreeAnc.optimize_branch_lengths_joint: THIS TREE HAS LONG BRANCHES. TreeTime's JOINT IS NOT DESIGNED TO "
                        "OPTIMIZE LONG BRANCHES; use branch_length_mode='input' or 'marginal'", 0, warn=True)
        self.tree.root.up = None
        self.tree.root.dist2root = 0.0
        self._prepare_nodes()
        return ttconf.SUCCESS

    # -- model inference -------------------------------------------------------------------------
    def infer_gtr(self, marginal=False, site_specific=False, normalized_rate=True, fixed_pi=None, pc=5.0, **kwargs):
        """treeanc.py:1500-1632, marginal branch: the n_ij / T_i accumulation over all branches
        (:1556-1572) is one device kernel; GTR.infer (gtr.py:491-599) stays on the host."""
        if site_specific and self.data.compress:
            raise TypeError('TreeAnc.infer_gtr(): sequence compression and site specific GTR models are incompatible!')
        if not self.ok:
            raise self._missing_data_error('TreeAnc.infer_gtr: ERROR, sequences or tree are missing')
        if not marginal:
            return self._infer_gtr_from_sequences(site_specific, normalized_rate, fixed_pi, pc, **kwargs)
        if self.sequence_reconstruction != 'marginal':
            self._ml_anc_marginal(**kwargs)
        q = self.gtr.n_states
        if site_specific or getattr(self.gtr, 'is_site_specific', False):
            # per-pattern statistics (treeanc.py:1551-1572 before the sum over patterns), sharded like the patterns
            n_ija, T_ia = self._engine.mutation_counts_per_site()
            if self.comm.world_size > 1:
                n_ija = self.comm.allgather(n_ija, axis=2)
                T_ia = self.comm.allgather(T_ia, axis=1)
            n_ij, T_i = n_ija.sum(axis=-1), T_ia.sum(axis=-1)
        else:
            n_ij, T_i = self._engine.mutation_counts()
            if self.comm.world_size > 1:
                red = self.comm.allreduce_sum(np.concatenate([n_ij.ravel(), T_i]))
                n_ij, T_i = red[:q * q].reshape(q, q), red[q * q:]
        if site_specific:
            root_state = self.tree.root.marginal_profile.T                      # treeanc.py:1594-1595
            self._gtr = self._infer_site_specific_gtr_from_counts(n_ija, T_ia, root_state, pc)
        else:
            root_cseq = self.tree.root.cseq
            m = self.data.multiplicity(mask=getattr(self.tree.root, 'mask', None))      # treeanc.py:1610
            root_state = np.array([np.sum((root_cseq == nuc) * m) for nuc in self.gtr.alphabet])
            self._gtr = self._infer_gtr_from_counts(n_ij, T_i, root_state, fixed_pi, pc)
        if normalized_rate:
            self.logger('TreeAnc.infer_gtr: setting overall rate to 1.0...', 2)
            if site_specific:
                self._gtr.mu /= self._gtr.average_rate().mean()                 # treeanc.py:1626-1627
            else:
                self._gtr.mu = 1.0
        return self._gtr

    def _infer_gtr_from_sequences(self, site_specific, normalized_rate, fixed_pi, pc, **kwargs):
        """infer_gtr(marginal=False) (treeanc.py:1573-1589): substitutions counted on the reconstructed sequences.
        Everything the reference collects from node.mutations and node.cseq is in the per-branch parent/child pair
        counts the device already provides (ttb_branch_state_pairs, N2): a mutation j -> i adds its multiplicity to
        n_ij and moves half the branch from T_i to T_j, every position adds the branch length to T_(child state);
        pairs with an ambiguous character are skipped like the reference's `except: continue` / `cseq == nuc`."""
        if site_specific:
            self._unsupported('site-specific GTR inference from reconstructed sequences runs in the reference')
        if not self.sequence_reconstruction:
            self._ml_anc_joint(**kwargs)
        C, F, tip_states, code_chars = self._pair_tables()               # raises Unsupported under masks
        topo = self._flat()
        q = self.gtr.n_states
        alphabet = [str(c) for c in self.gtr.alphabet]
        t = self._branch_lengths_to_gtr(topo.nodes)[1:]
        is_tip = np.array([n.is_terminal() for n in topo.nodes[1:]])
        M = np.array(C[:, :, :q])                                        # [branch, parent j, child i]
        if not tip_states and is_tip.any():
            # tips show their alignment characters: keep the columns that are plain alphabet letters
            col = {ch: k for k, ch in enumerate(code_chars) if ch is not None and k < C.shape[2]}
            Mt = np.zeros((int(is_tip.sum()), q, q))
            for i, ch in enumerate(alphabet):
                if ch in col:
                    Mt[:, :, i] = C[is_tip][:, :, col[ch]]
            M[is_tip] = Mt
        child = M.sum(axis=1)                                            # positions per child state
        off = M.copy()
        ar = np.arange(q)
        off[:, ar, ar] = 0
        n_ij = off.sum(axis=0).T                                         # i = derived, j = ancestral state
        T_i = (t[:, None] * (child + 0.5 * off.sum(axis=2) - 0.5 * off.sum(axis=1))).sum(axis=0)
        root_cseq = self.tree.root.cseq
        m = self.data.multiplicity(mask=getattr(self.tree.root, 'mask', None))
        root_state = np.array([np.sum((root_cseq == nuc) * m) for nuc in self.gtr.alphabet])
        self._gtr = self._infer_gtr_from_counts(n_ij, T_i, root_state, fixed_pi, pc)
        if normalized_rate:
            self.logger('TreeAnc.infer_gtr: setting overall rate to 1.0...', 2)
            self._gtr.mu = 1.0
        return self._gtr

    def _infer_site_specific_gtr_from_counts(self, n_ija, T_ia, root_state, pc):
        from .gtr import infer_site_specific_gtr_from_counts
        return infer_site_specific_gtr_from_counts(n_ija, T_ia, root_state, pc=pc, alphabet=self.gtr.alphabet,
                                                   prof_map=self.gtr.profile_map, logger=self.logger)

    def _infer_gtr_from_counts(self, n_ij, T_i, root_state, fixed_pi, pc):
        return infer_gtr_from_counts(n_ij, T_i, root_state, fixed_pi=fixed_pi, pc=pc, alphabet=self.gtr.alphabet,
                                     prof_map=self.gtr.profile_map, logger=self.logger)

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
        self._invalidate_after_lh_only()

    def _invalidate_after_lh_only(self):
        """The LH-only trial passes overwrote S and P on the device but not M: outgroup messages, branch objectives and
        substitution counts recomputed from M_old / (S_new P_new) would mix two passes, and arrays cached on the host come
        from yet another state.  The reference's stored arrays stay self-consistent (they belong to the last trial rate);
        here one full pass at the final rate makes everything consistent again (same options as the last reconstruction)."""
        if self.sequence_reconstruction == 'marginal' and self._engine is not None:
            eng = self._sync_device()
            eng.marginal(reconstruct_tips=bool(self.reconstructed_tip_sequences))
            tot, _ = eng.results()
            if self.comm.world_size > 1:
                tot = self.comm.allreduce_sum(np.array([tot]))[0]
            self.tree.sequence_LH = self._gather_patterns(eng.site_lh())
            self.tree.total_sequence_LH = float(tot)
            self.tree.sequence_marginal_LH = self.tree.total_sequence_LH
        self._cache = {}
        self._seq_cache = {}
        drop = getattr(self, '_drop_node_caches', None)
        if drop is not None and getattr(self, '_b200_live', False):
            drop()
            self._b200_live = True

