"""SeqGen on the device: evolve sequences down a tree under a GTR model.

Mirrors treetime.seqgen.SeqGen (seqgen.py:9-95): `SeqGen(L, tree=..., gtr=...)`, `evolve(root_seq=None)`,
`get_aln(internal=False)`.  With `reference_rng=True` the uniform numbers are drawn on the host from the
object's numpy Generator in the reference's order (root, then the children of every internal node in
preorder), which reproduces the reference's sequences bit for bit; otherwise a counter-based Philox stream
runs on the device (same distribution, different numbers) and nothing but the result crosses PCIe."""
import numpy as np

from .flatten import FlatTopology, code_table, gtr_arrays
from .tree import Tree, read_newick


class SeqGen(object):
    def __init__(self, L, tree=None, gtr=None, rng_seed=None, device=0, engine_factory=None, verbose=0, **kwargs):
        if tree is None or gtr is None:
            raise ValueError('SeqGen needs a tree and a GTR model')
        self.seq_len = int(L)
        self.gtr = gtr
        self.tree = tree if isinstance(tree, Tree) or hasattr(tree, 'root') else read_newick(tree)
        for n in self.tree.find_clades():
            n.branch_length = n.branch_length if n.branch_length else 0.0
        self.tree.ladderize()                   # TreeAnc.prepare_tree (treeanc.py:373-386): fixes the child order = draw order
        self.rng = np.random.default_rng(seed=rng_seed)
        self.device = device
        if engine_factory is None:
            from .engine import Engine
            engine_factory = Engine
        self._engine_factory = engine_factory
        self._engine = None
        self.topo = FlatTopology(self.tree.root)
        for i, n in enumerate(self.topo.nodes):
            n._fid = i
        self.aln = None

    def _prepare(self):
        topo, q = self.topo, self.gtr.n_states
        if self._engine is None:
            self._engine = self._engine_factory(q, self.device)
            self._engine.set_tree(topo.parent, topo.child_ptr, topo.child_idx, topo.tip_row)
        eng = self._engine
        chars, lut, table = code_table(self.gtr.profile_map, q)
        eng.set_patterns(np.zeros((topo.n_tips, self.seq_len), dtype=np.uint8), table, np.ones(self.seq_len), validate=False)
        eng.set_gtr(gtr_arrays(self.gtr))
        # the reference evolves over the raw branch lengths (seqgen.py:64), no flooring
        eng.set_branch_lengths(np.array([0.0] + [float(n.branch_length or 0.0) for n in topo.nodes[1:]]))
        return eng, np.array([lut[str(c)] for c in self.gtr.alphabet], dtype=np.uint8)

    def evolve(self, root_seq=None, reference_rng=False):
        """seqgen.py:38-67.  Stores node.ancestral_sequence (character arrays) and self.aln (tips)."""
        eng, state2code = self._prepare()
        topo, L = self.topo, self.seq_len
        root_idx = None
        if root_seq is not None and len(root_seq):
            lut = {str(c): i for i, c in enumerate(self.gtr.alphabet)}
            root_idx = np.array([lut[c] for c in np.asarray(list(root_seq) if isinstance(root_seq, str) else root_seq).astype('U1')], dtype=np.uint8)
            if root_idx.shape[0] != L:
                raise ValueError('root sequence length does not match L')
        uniforms = None
        if reference_rng:
            uniforms = np.zeros((topo.n_nodes, L))
            if root_idx is None:
                uniforms[0] = self.rng.random(L)
            for n in topo.nodes:                         # preorder; tips have no children
                for c in n.clades:
                    uniforms[c._fid] = self.rng.random(L)
            seed = 0
        else:
            seed = int(self.rng.integers(0, 2 ** 63 - 1))
        self.states = eng.seqgen(seed, state2code, root_idx=root_idx, uniforms=uniforms)
        for n in topo.nodes:
            n.ancestral_sequence = self.gtr.alphabet[self.states[n._fid]]
        self.aln = self.get_aln()
        return self.aln

    def get_aln(self, internal=False):
        """seqgen.py:69-95: {name: character array} of the tips (and internal nodes on request)."""
        return {n.name: n.ancestral_sequence for n in self.topo.nodes if n.is_terminal() or internal}

    @property
    def engine(self):
        """The engine holding the generated alignment as its tip codes (ready for ttb_marginal)."""
        return self._engine
