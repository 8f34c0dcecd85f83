from .Seq import Seq


class SeqRecord(object):
    def __init__(self, seq=None, id="<unknown id>", name="<unknown name>", description="<unknown description>"):
        self.seq = seq if isinstance(seq, Seq) or seq is None else Seq(seq)
        self.id = id
        self.name = name
        self.description = description

    def __len__(self):
        return len(self.seq)

    def __str__(self):
        return str(self.seq)
