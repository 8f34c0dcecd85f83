from . import SeqIO
from .Align import MultipleSeqAlignment


def read(handle, fmt="fasta"):
    return MultipleSeqAlignment(list(SeqIO.parse(handle, fmt)))


def write(aln, handle, fmt="fasta"):
    return SeqIO.write(list(aln), handle, fmt)
