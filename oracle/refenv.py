"""TEST-ONLY: make the UNMODIFIED reference importable in the build container.

Puts oracle/bioshim (a stand-in for the missing Biopython) and /root/reference
first on sys.path.  Used by oracle/make_golden.py and
oracle/validate_against_reference.py, and by tests that are skipped when
/root/reference does not exist (i.e. on the GPU box)."""
import os
import sys

REFERENCE = os.environ.get('TREETIME_REFERENCE', '/root/reference')
HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)


def available():
    return os.path.isdir(os.path.join(REFERENCE, 'treetime'))


def activate():
    if not available():
        raise RuntimeError('reference not present at %s' % REFERENCE)
    for p in (REFERENCE, os.path.join(HERE, 'bioshim'), REPO):
        if p not in sys.path:
            sys.path.insert(0, p)
    import treetime  # noqa: F401
    return treetime


def reference_treeanc(newick, aln_chars, gtr, **kw):
    """Build the reference's TreeAnc from a newick string and a dict
    name -> numpy char array (or str)."""
    activate()
    from io import StringIO
    from Bio import Phylo
    from Bio.Align import MultipleSeqAlignment
    from Bio.SeqRecord import SeqRecord
    from Bio.Seq import Seq
    from treetime import TreeAnc
    tree = Phylo.read(StringIO(newick), 'newick')
    aln = MultipleSeqAlignment([SeqRecord(Seq(''.join(aln_chars[k])), id=k, name=k, description='') for k in aln_chars])
    kw.setdefault('verbose', 0)
    return TreeAnc(tree=tree, aln=aln, gtr=gtr, **kw)
