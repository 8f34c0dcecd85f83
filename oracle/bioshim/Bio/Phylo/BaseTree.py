"""Minimal tree containers with the Bio.Phylo.BaseTree semantics the reference
relies on (child order = `clades` order; DFS traversals; stable ladderize)."""
import sys

sys.setrecursionlimit(max(sys.getrecursionlimit(), 100000))


class TreeMixin(object):
    def find_clades(self, target=None, terminal=None, order="preorder", **kwargs):
        root = self.root
        if order == "preorder":
            it = _preorder(root)
        elif order == "postorder":
            it = _postorder(root)
        elif order == "level":
            it = _level(root)
        else:
            raise ValueError(order)
        for c in it:
            if terminal is None or c.is_terminal() == terminal:
                yield c

    def get_terminals(self, order="preorder"):
        return list(self.find_clades(terminal=True, order=order))

    def get_nonterminals(self, order="preorder"):
        return list(self.find_clades(terminal=False, order=order))

    def count_terminals(self):
        return sum(1 for _ in self.find_clades(terminal=True))

    def total_branch_length(self):
        return sum(c.branch_length for c in self.find_clades() if c.branch_length)

    def ladderize(self, reverse=False):
        # iterative: count tips per clade in postorder, then stable sort
        counts = {}
        for c in _postorder(self.root):
            counts[id(c)] = 1 if not c.clades else sum(counts[id(ch)] for ch in c.clades)
        for c in _preorder(self.root):
            c.clades.sort(key=lambda x: counts[id(x)], reverse=reverse)

    def is_bifurcating(self):
        return all(len(c.clades) in (0, 2) for c in self.find_clades())

    def get_path(self, target):
        path = []
        n = target
        while getattr(n, "up", None) is not None:
            path.append(n)
            n = n.up
        return path[::-1]


def _preorder(root):
    stack = [root]
    while stack:
        n = stack.pop()
        yield n
        stack.extend(reversed(n.clades))


def _postorder(root):
    stack = [(root, 0)]
    while stack:
        n, i = stack.pop()
        if i < len(n.clades):
            stack.append((n, i + 1))
            stack.append((n.clades[i], 0))
        else:
            yield n


def _level(root):
    from collections import deque
    q = deque([root])
    while q:
        n = q.popleft()
        yield n
        q.extend(n.clades)


class Clade(TreeMixin):
    def __init__(self, branch_length=None, name=None, clades=None, confidence=None, color=None, width=None):
        self.branch_length = branch_length
        self.name = name
        self.clades = clades or []
        self.confidence = confidence

    @property
    def root(self):
        return self

    def is_terminal(self):
        return not self.clades

    def __iter__(self):
        return iter(self.clades)

    def __len__(self):
        return len(self.clades)

    def __getitem__(self, i):
        return self.clades[i]

    def __bool__(self):
        return True

    def __repr__(self):
        return "Clade(name=%r, branch_length=%r)" % (self.name, self.branch_length)


class Tree(TreeMixin):
    def __init__(self, root=None, rooted=True, id=None, name=None):
        self.root = root or Clade()
        self.rooted = rooted
        self.id = id
        self.name = name

    @classmethod
    def from_clade(cls, clade, **kw):
        return cls(root=clade, **kw)

    @property
    def clade(self):
        return self.root

    def is_terminal(self):
        return self.root.is_terminal()

    def root_with_outgroup(self, *a, **k):
        raise NotImplementedError("Bio shim: rerooting is outside the hot path")
