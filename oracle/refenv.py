"""TEST-ONLY: make the UNMODIFIED reference importable.

Puts oracle/bioshim (a stand-in for the missing Biopython) and the reference first on sys.path.  The
reference is /root/reference in the build container or -- on the GPU box, where that path does not exist
-- the pip-installed, git-ignored copy oracle/_ref made by oracle/stage_ref.py.  Used by
oracle/make_golden.py, oracle/validate_against_reference.py, the live-reference tests (skipped when
neither exists) and bench.py's `--impl reference` arm."""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
STAGED = os.path.join(HERE, '_ref')


def _find():
    for p in (os.environ.get('TREETIME_REFERENCE'), '/root/reference', STAGED):
        if p and os.path.isfile(os.path.join(p, 'treetime', 'treeanc.py')):
            return p
    return os.environ.get('TREETIME_REFERENCE', '/root/reference')


REFERENCE = _find()


def available():
    return os.path.isfile(os.path.join(REFERENCE, 'treetime', 'treeanc.py'))


def is_staged_copy():
    return os.path.abspath(REFERENCE) == os.path.abspath(STAGED)


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
