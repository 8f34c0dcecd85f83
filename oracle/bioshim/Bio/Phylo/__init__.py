from . import BaseTree
from ._newick import read, write, parse_newick_string  # noqa: F401
