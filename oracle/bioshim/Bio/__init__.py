"""TEST-ONLY stand-in for Biopython (not installed in this image, no network).

It exists so that the UNMODIFIED reference under /root/reference can be imported
in the build container to (a) validate oracle/flat_numpy.py and (b) generate the
golden fixtures under tests/golden/.  Nothing in treetime_b200/ imports it.

Only the container types the reference touches on the marginal-reconstruction
path are provided (tree/clade, alignment records, newick reader); no numerics
live in Biopython on that path.
"""
__version__ = "1.85+shim"
