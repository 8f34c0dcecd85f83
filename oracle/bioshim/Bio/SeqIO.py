from .Seq import Seq
from .SeqRecord import SeqRecord


def parse(handle, fmt="fasta"):
    if fmt != "fasta":
        raise ValueError("Bio shim: only fasta is supported")
    close = False
    if isinstance(handle, str):
        handle = open(handle)
        close = True
    name, chunks = None, []
    try:
        for line in handle:
            line = line.rstrip("\n\r")
            if line.startswith(">"):
                if name is not None:
                    yield SeqRecord(Seq("".join(chunks)), id=name.split()[0] if name else "", name=name.split()[0] if name else "", description=name)
                name, chunks = line[1:], []
            elif line:
                chunks.append(line.strip())
        if name is not None:
            yield SeqRecord(Seq("".join(chunks)), id=name.split()[0] if name else "", name=name.split()[0] if name else "", description=name)
    finally:
        if close:
            handle.close()


def read(handle, fmt="fasta"):
    recs = list(parse(handle, fmt))
    if len(recs) != 1:
        raise ValueError("expected exactly one record")
    return recs[0]


def write(records, handle, fmt="fasta"):
    close = False
    if isinstance(handle, str):
        handle = open(handle, "w")
        close = True
    n = 0
    for r in records:
        handle.write(">%s\n%s\n" % (r.id, str(r.seq)))
        n += 1
    if close:
        handle.close()
    return n
