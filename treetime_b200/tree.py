"""Minimal rooted-tree container for the host side.

The reference holds its tree in Bio.Phylo objects (not installed here, and only
a container -- no arithmetic).  This module gives the same small surface the
marginal path touches: `clades` (ordered children), `branch_length`, `name`,
`up`, preorder/postorder traversal in `clades` order and a stable `ladderize`
(treeanc.py:456-457 relies on it for the child order).  All traversals are
iterative so 100k-tip caterpillar trees do not hit the recursion limit.
"""
from collections import deque


class Node(object):
    def __init__(self, name=None, branch_length=None, clades=None):
        self.name = name
        self.branch_length = branch_length
        self.clades = clades if clades is not None else []
        self.up = None
        self.mask = None

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
        return 'Node(name=%r, branch_length=%r)' % (self.name, self.branch_length)

    # the traversal helpers are shared with Tree
    def find_clades(self, terminal=None, order='preorder'):
        return _find(self, terminal, order)

    def get_terminals(self, order='preorder'):
        return list(_find(self, True, order))

    def get_nonterminals(self, order='preorder'):
        return list(_find(self, False, order))

    def count_terminals(self):
        return sum(1 for _ in _find(self, True, 'preorder'))


def _find(root, terminal, order):
    if order == 'preorder':
        it = _preorder(root)
    elif order == 'postorder':
        it = _postorder(root)
    elif order == 'level':
        it = _level(root)
    else:
        raise ValueError("order must be 'preorder', 'postorder' or 'level'")
    for c in it:
        if terminal is None or c.is_terminal() == terminal:
            yield c


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
    q = deque([root])
    while q:
        n = q.popleft()
        yield n
        q.extend(n.clades)


class Tree(object):
    def __init__(self, root=None, rooted=True):
        self.root = root if root is not None else Node()
        self.rooted = rooted

    def find_clades(self, terminal=None, order='preorder'):
        return _find(self.root, terminal, order)

    def get_terminals(self, order='preorder'):
        return list(_find(self.root, True, order))

    def get_nonterminals(self, order='preorder'):
        return list(_find(self.root, False, order))

    def count_terminals(self):
        return self.root.count_terminals()

    def total_branch_length(self):
        return sum(c.branch_length for c in self.find_clades() if c.branch_length)

    def ladderize(self, reverse=False):
        """Stable sort of every node's children by number of tips (ascending)."""
        counts = {}
        for c in _postorder(self.root):
            counts[id(c)] = 1 if not c.clades else sum(counts[id(ch)] for ch in c.clades)
        for c in _preorder(self.root):
            c.clades.sort(key=lambda x: counts[id(x)], reverse=reverse)

    def is_bifurcating(self):
        return all(len(c.clades) in (0, 2) for c in self.find_clades())

    def to_newick(self, fmt='%r'):
        return to_newick(self.root, fmt) + ';'


def read_newick(src):
    """Parse a newick string, a file name or a file object into a Tree."""
    import os
    if hasattr(src, 'read'):
        s = src.read()
    elif isinstance(src, str) and ('(' not in src) and os.path.isfile(src):
        with open(src) as fh:
            s = fh.read()
    else:
        s = src
    s = s.strip()
    pos, n = 0, len(s)
    root = Node()
    stack = []
    cur = root
    while pos < n:
        ch = s[pos]
        if ch == '(':
            child = Node()
            cur.clades.append(child)
            stack.append(cur)
            cur = child
            pos += 1
        elif ch == ',':
            child = Node()
            stack[-1].clades.append(child)
            cur = child
            pos += 1
        elif ch == ')':
            cur = stack.pop()
            pos += 1
        elif ch == ';':
            break
        elif ch == ':':
            pos += 1
            start = pos
            while pos < n and s[pos] not in ',();[':
                pos += 1
            cur.branch_length = float(s[start:pos])
        elif ch == '[':
            pos = s.index(']', pos) + 1
        elif ch.isspace():
            pos += 1
        else:
            if ch in '\'"':
                end = s.index(ch, pos + 1)
                lab = s[pos + 1:end]
                pos = end + 1
            else:
                start = pos
                while pos < n and s[pos] not in ',():;[':
                    pos += 1
                lab = s[start:pos].strip()
            if lab:
                if cur.clades:
                    try:
                        cur.confidence = float(lab)
                    except ValueError:
                        cur.name = lab
                else:
                    cur.name = lab
    if stack:
        raise ValueError('unbalanced parentheses in newick string')
    return Tree(root=root)


def to_newick(root, fmt='%r'):
    out = []
    stack = [(root, 0)]
    while stack:
        c, i = stack.pop()
        if i == 0 and c.clades:
            out.append('(')
        if i < len(c.clades):
            if i > 0:
                out.append(',')
            stack.append((c, i + 1))
            stack.append((c.clades[i], 0))
        else:
            if c.clades:
                out.append(')')
            out.append(c.name or '')
            if c.branch_length is not None:
                out.append(':' + (fmt % float(c.branch_length)))
    return ''.join(out)
