class Seq(str):
    """str subclass; the reference only does str(seq) / isinstance checks."""
    def __new__(cls, data=""):
        return super().__new__(cls, str(data))
