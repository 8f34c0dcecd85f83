"""Numerical constants of the marginal path (values as in the reference's
treetime/config.py:3-7,34; they are part of the arithmetic contract)."""
BIG_NUMBER = 1e10
TINY_NUMBER = 1e-12
SUPERTINY_NUMBER = 1e-24
MIN_BRANCH_LENGTH = 1e-3   # in units of one_mutation = 1/full_length
MAX_BRANCH_LENGTH = 4.0
SUCCESS = 'success'
ERROR = 'error'
VERBOSE = 3
