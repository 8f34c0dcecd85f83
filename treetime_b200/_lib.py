"""ctypes binding of libttb.so (the C-ABI declared in include/ttb.h).

There is no CPU fallback: if the shared library is missing or no CUDA device is
usable, constructing an engine raises.  `build()` compiles the library in-tree
with nvcc for sm_100a (cross-compiles without a GPU).
"""
import ctypes
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, 'libttb.so')
SOURCES = [os.path.join(HERE, 'csrc', 'ttb_api.cu')]
HEADERS = [os.path.join(HERE, 'csrc', 'ttb_kernels.cuh'), os.path.join(os.path.dirname(HERE), 'include', 'ttb.h')]

NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
              '-Xcompiler', '-fPIC', '-shared']


class TTBError(RuntimeError):
    """Error reported by libttb.so (carries the library's message and code)."""

    def __init__(self, code, msg):
        super(TTBError, self).__init__('libttb error %d: %s' % (code, msg))
        self.code = code


def needs_build():
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    return any(os.path.getmtime(s) > t for s in SOURCES + HEADERS if os.path.exists(s))


def build(force=False, verbose=False):
    """Compile csrc/*.cu -> treetime_b200/libttb.so for sm_100a."""
    if not force and not needs_build():
        return LIB_PATH
    nvcc = os.environ.get('NVCC', 'nvcc')
    cmd = [nvcc] + NVCC_FLAGS + ['-o', LIB_PATH] + SOURCES
    if verbose:
        print(' '.join(cmd), file=sys.stderr)
    subprocess.run(cmd, check=True)
    return LIB_PATH


_lib = None

_c_int_p = ctypes.POINTER(ctypes.c_int32)
_c_dbl_p = ctypes.POINTER(ctypes.c_double)
_c_u8_p = ctypes.POINTER(ctypes.c_uint8)
_H = ctypes.c_void_p

# name -> argtypes; must list every symbol include/ttb.h declares (tests check this)
SIGNATURES = {
    'ttb_last_error': ([], ctypes.c_char_p),
    'ttb_version': ([], ctypes.c_int),
    'ttb_supports_n_states': ([ctypes.c_int], ctypes.c_int),
    'ttb_create': ([ctypes.POINTER(_H), ctypes.c_int, ctypes.c_int], ctypes.c_int),
    'ttb_destroy': ([_H], ctypes.c_int),
    'ttb_set_stream': ([_H, ctypes.c_void_p], ctypes.c_int),
    'ttb_set_tree': ([_H, ctypes.c_int32, _c_int_p, _c_int_p, _c_int_p, _c_int_p], ctypes.c_int),
    'ttb_set_patterns': ([_H, ctypes.c_int64, _c_u8_p, ctypes.c_int32, _c_dbl_p, _c_dbl_p], ctypes.c_int),
    'ttb_set_gtr': ([_H, _c_dbl_p, _c_dbl_p, _c_dbl_p, _c_dbl_p, ctypes.c_double, ctypes.c_int32], ctypes.c_int),
    'ttb_set_gtr_site_specific': ([_H, _c_dbl_p, _c_dbl_p, _c_dbl_p, _c_dbl_p, _c_dbl_p, _c_dbl_p, ctypes.c_int32,
                                   ctypes.c_double, ctypes.c_int32, ctypes.c_int32], ctypes.c_int),
    'ttb_set_branch_lengths': ([_H, _c_dbl_p], ctypes.c_int),
    'ttb_marginal': ([_H, ctypes.c_int32], ctypes.c_int),
    'ttb_results': ([_H, _c_dbl_p, ctypes.POINTER(ctypes.c_int64)], ctypes.c_int),
    'ttb_results_device_ptr': ([_H, ctypes.POINTER(ctypes.c_void_p)], ctypes.c_int),
    'ttb_sync': ([_H], ctypes.c_int),
    'ttb_fetch_site_lh': ([_H, _c_dbl_p], ctypes.c_int),
    'ttb_fetch_node': ([_H, ctypes.c_int32, ctypes.c_int32, _c_dbl_p], ctypes.c_int),
    'ttb_fetch_seq_idx': ([_H, ctypes.c_int32, _c_int_p, _c_u8_p], ctypes.c_int),
    'ttb_branch_objective': ([_H, ctypes.c_int32, _c_int_p, _c_int_p, _c_dbl_p, _c_dbl_p], ctypes.c_int),
    'ttb_branch_hamming': ([_H, ctypes.c_int32, _c_int_p, _c_int_p, _c_dbl_p, _c_dbl_p], ctypes.c_int),
    'ttb_mutation_counts': ([_H, _c_dbl_p, _c_dbl_p], ctypes.c_int),
    'ttb_device_bytes': ([_H, ctypes.POINTER(ctypes.c_int64)], ctypes.c_int),
    'ttb_launch_count': ([_H, ctypes.POINTER(ctypes.c_int64)], ctypes.c_int),
}


def load():
    """dlopen libttb.so and attach the prototypes.  Raises if it was not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError('%s not found: run `python -c "import __graft_entry__ as g; g.build()"` '
                          '(or treetime_b200._lib.build()) first. There is no CPU fallback.' % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (argtypes, restype) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = restype
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        raise TTBError(rc, load().ttb_last_error().decode('utf-8', 'replace'))
