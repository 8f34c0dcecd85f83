from io import StringIO
from .BaseTree import Clade, Tree


def parse_newick_string(s):
    s = s.strip()
    pos = 0
    n = len(s)

    def parse_label():
        nonlocal pos
        if pos < n and s[pos] in "'\"":
            q = s[pos]
            end = s.index(q, pos + 1)
            lab = s[pos + 1:end]
            pos = end + 1
            return lab
        start = pos
        while pos < n and s[pos] not in ",():;[":
            pos += 1
        return s[start:pos].strip()

    def skip_comment():
        nonlocal pos
        while pos < n and s[pos] == "[":
            pos = s.index("]", pos) + 1

    root = Clade()
    stack = [root]
    cur = root
    # iterative parser (deep trees)
    while pos < n:
        ch = s[pos]
        if ch == "(":
            child = Clade()
            cur.clades.append(child)
            stack.append(cur)
            cur = child
            pos += 1
        elif ch == ",":
            parent = stack[-1]
            child = Clade()
            parent.clades.append(child)
            cur = child
            pos += 1
        elif ch == ")":
            cur = stack.pop()
            pos += 1
        elif ch == ";":
            break
        elif ch == ":":
            pos += 1
            start = pos
            while pos < n and s[pos] not in ",();[":
                pos += 1
            cur.branch_length = float(s[start:pos])
        elif ch == "[":
            skip_comment()
        elif ch.isspace():
            pos += 1
        else:
            lab = parse_label()
            if lab:
                if cur.clades:
                    try:
                        cur.confidence = float(lab)
                    except ValueError:
                        cur.name = lab
                else:
                    cur.name = lab
    return Tree(root=root, rooted=True)


def read(file, fmt="newick", **kw):
    if fmt != "newick":
        raise ValueError("Bio shim: only newick is supported")
    if isinstance(file, str):
        with open(file) as fh:
            txt = fh.read()
    else:
        txt = file.read()
    return parse_newick_string(txt)


def _fmt(clade):
    out = []
    # iterative serialisation
    stack = [(clade, 0)]
    while stack:
        c, i = stack.pop()
        if i == 0 and c.clades:
            out.append("(")
        if i < len(c.clades):
            if i > 0:
                out.append(",")
            stack.append((c, i + 1))
            stack.append((c.clades[i], 0))
        else:
            if c.clades:
                out.append(")")
            out.append(c.name or "")
            if c.branch_length is not None:
                out.append(":%r" % float(c.branch_length))
    return "".join(out)


def write(tree, file, fmt="newick", **kw):
    txt = _fmt(tree.root) + ";\n"
    if isinstance(file, str):
        with open(file, "w") as fh:
            fh.write(txt)
    else:
        file.write(txt)
    return 1


def to_string(tree):
    return _fmt(tree.root) + ";"
