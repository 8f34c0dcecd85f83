class MultipleSeqAlignment(object):
    def __init__(self, records=(), alphabet=None):
        self._records = list(records)

    def __iter__(self):
        return iter(self._records)

    def __len__(self):
        return len(self._records)

    def __getitem__(self, i):
        return self._records[i]

    def append(self, rec):
        self._records.append(rec)

    def get_alignment_length(self):
        return len(self._records[0].seq) if self._records else 0
