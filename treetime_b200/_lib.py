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
LIB_PATH = os.environ.get('TTB_LIB') or os.path.join(HERE, 'libttb.so')      # TTB_LIB: A/B measurements against another build
CSRC = os.path.join(HERE, 'csrc')
BUILD_DIR = os.path.join(HERE, 'build')
Q_VALUES = (2, 3, 4, 5, 6, 7, 8, 20, 21, 22)      # alphabet sizes with compiled kernels
SOURCES = [os.path.join(CSRC, 'ttb_api.cu'), os.path.join(CSRC, 'ttb_q.cu'), os.path.join(CSRC, 'ttb_brent.cu')]
HEADERS = [os.path.join(CSRC, f) for f in ('ttb_kernels.cuh', 'ttb_mma.cuh', 'ttb_qops.h', 'ttb_brent.h')] + \
          [os.path.join(os.path.dirname(HERE), 'include', 'ttb.h')]

NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
              '-Xcompiler', '-fPIC']


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


def build(force=False, verbose=False, extra_flags=()):
    """Compile csrc/*.cu -> treetime_b200/libttb.so for sm_100a (nvcc cross-compiles without
    a GPU).  One object per alphabet size (ttb_q.cu with -DTTB_Q=q), built in parallel."""
    if not force and not needs_build():
        return LIB_PATH
    from concurrent.futures import ThreadPoolExecutor
    nvcc = os.environ.get('NVCC', 'nvcc')
    os.makedirs(BUILD_DIR, exist_ok=True)
    jobs = [([nvcc] + NVCC_FLAGS + list(extra_flags) + ['-c', os.path.join(CSRC, 'ttb_api.cu'), '-o',
                                                        os.path.join(BUILD_DIR, 'ttb_api.o')])]
    # the Brent state machine mirrors scipy's scalar arithmetic: no fused multiply-adds
    jobs.append([nvcc] + NVCC_FLAGS + list(extra_flags) + ['-fmad=false', '-c', os.path.join(CSRC, 'ttb_brent.cu'), '-o',
                                                           os.path.join(BUILD_DIR, 'ttb_brent.o')])
    for q in Q_VALUES:
        jobs.append([nvcc] + NVCC_FLAGS + list(extra_flags) + ['-DTTB_Q=%d' % q, '-c', os.path.join(CSRC, 'ttb_q.cu'),
                                                               '-o', os.path.join(BUILD_DIR, 'ttb_q%d.o' % q)])

    def run(cmd):
        if verbose:
            print(' '.join(cmd), file=sys.stderr)
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError('nvcc failed:\n%s\n%s' % (' '.join(cmd), r.stderr))
        return r.stderr

    with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 4)) as ex:
        logs = list(ex.map(run, jobs))
    objs = [j[-1] for j in jobs]
    run([nvcc, '-shared', '-o', LIB_PATH] + objs)
    return logs


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
    'ttb_set_message_storage': ([_H, ctypes.c_int32], ctypes.c_int),
    'ttb_set_tree': ([_H, ctypes.c_int32, _c_int_p, _c_int_p, _c_int_p, _c_int_p], ctypes.c_int),
    'ttb_set_patterns': ([_H, ctypes.c_int64, _c_u8_p, ctypes.c_int32, _c_dbl_p, _c_dbl_p], ctypes.c_int),
    'ttb_set_patterns_sparse': ([_H, ctypes.c_int64, _c_u8_p, ctypes.c_int64, _c_int_p, _c_int_p, _c_u8_p, ctypes.c_int32,
                                 _c_dbl_p, _c_dbl_p], ctypes.c_int),
    'ttb_alignment_stats': ([_H, ctypes.c_int64, ctypes.c_int64, _c_u8_p, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32,
                             ctypes.c_int32, _c_u8_p, _c_u8_p, _c_u8_p], ctypes.c_int),
    'ttb_set_patterns_from_alignment': ([_H, ctypes.c_int64, ctypes.POINTER(ctypes.c_int64), _c_u8_p, _c_int_p, _c_u8_p,
                                         ctypes.c_int32, ctypes.c_int32, _c_dbl_p, _c_dbl_p], ctypes.c_int),
    'ttb_set_gtr': ([_H, _c_dbl_p, _c_dbl_p, _c_dbl_p, _c_dbl_p, ctypes.c_double, ctypes.c_int32], ctypes.c_int),
    'ttb_set_gtr_site_specific': ([_H, _c_dbl_p, _c_dbl_p, _c_dbl_p, _c_dbl_p, _c_dbl_p, _c_dbl_p, ctypes.c_int32,
                                   ctypes.c_double, ctypes.c_int32, ctypes.c_int32], ctypes.c_int),
    'ttb_set_branch_lengths': ([_H, _c_dbl_p], ctypes.c_int),
    'ttb_set_branch_masks': ([_H, ctypes.c_int32, _c_u8_p, _c_int_p], ctypes.c_int),
    'ttb_marginal': ([_H, ctypes.c_int32], ctypes.c_int),
    'ttb_joint': ([_H, ctypes.c_int32], ctypes.c_int),
    'ttb_joint_retrace': ([_H, _c_u8_p, ctypes.c_int32], ctypes.c_int),
    'ttb_sample_states': ([_H, ctypes.c_int32, _c_int_p, _c_dbl_p, ctypes.POINTER(ctypes.c_int64), ctypes.POINTER(ctypes.c_int64)],
                          ctypes.c_int),
    'ttb_results': ([_H, _c_dbl_p, ctypes.POINTER(ctypes.c_int64)], ctypes.c_int),
    'ttb_results_tips': ([_H, ctypes.POINTER(ctypes.c_int64)], ctypes.c_int),
    'ttb_results_device_ptr': ([_H, ctypes.POINTER(ctypes.c_void_p)], ctypes.c_int),
    'ttb_sync': ([_H], ctypes.c_int),
    'ttb_fetch_site_lh': ([_H, _c_dbl_p], ctypes.c_int),
    'ttb_fetch_node': ([_H, ctypes.c_int32, ctypes.c_int32, _c_dbl_p], ctypes.c_int),
    'ttb_fetch_seq_idx': ([_H, ctypes.c_int32, _c_int_p, _c_u8_p], ctypes.c_int),
    'ttb_fetch_all_seq_idx': ([_H, _c_u8_p], ctypes.c_int),
    'ttb_fetch_mutations': ([_H, _c_u8_p, ctypes.c_int32, _c_int_p, _c_int_p, _c_u8_p, ctypes.POINTER(ctypes.c_int64)], ctypes.c_int),
    'ttb_enqueue_fetch_site_lh': ([_H, _c_dbl_p], ctypes.c_int),
    'ttb_enqueue_fetch_all_seq_idx': ([_H, _c_u8_p], ctypes.c_int),
    'ttb_profile_marginal': ([_H, ctypes.c_int32, _c_dbl_p, _c_int_p], ctypes.c_int),
    'ttb_branch_objective': ([_H, ctypes.c_int32, _c_int_p, _c_int_p, _c_dbl_p, _c_dbl_p], ctypes.c_int),
    'ttb_brent_begin': ([_H, ctypes.c_int32, _c_int_p, _c_int_p, _c_dbl_p, _c_dbl_p, _c_dbl_p, ctypes.c_double, ctypes.c_int32], ctypes.c_int),
    'ttb_brent_eval': ([_H], ctypes.c_int),
    'ttb_brent_f_device_ptr': ([_H, ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.c_int32)], ctypes.c_int),
    'ttb_brent_update': ([_H, ctypes.c_int32, ctypes.POINTER(ctypes.c_int32)], ctypes.c_int),
    'ttb_brent_result': ([_H, _c_dbl_p, _c_dbl_p, _c_int_p, _c_int_p], ctypes.c_int),
    'ttb_branch_hamming': ([_H, ctypes.c_int32, _c_int_p, _c_int_p, _c_dbl_p, _c_dbl_p], ctypes.c_int),
    'ttb_mutation_counts': ([_H, _c_dbl_p, _c_dbl_p], ctypes.c_int),
    'ttb_mutation_counts_per_site': ([_H, _c_dbl_p, _c_dbl_p], ctypes.c_int),
    'ttb_seqgen': ([_H, ctypes.c_uint64, _c_u8_p, _c_dbl_p, _c_u8_p, _c_u8_p], ctypes.c_int),
    'ttb_branch_state_pairs': ([_H, ctypes.c_int32, _c_int_p, ctypes.c_int32, ctypes.c_int32, _c_dbl_p, _c_int_p], ctypes.c_int),
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
        if os.environ.get('TTB_LIB') and not hasattr(lib, name):
            continue                      # an older build under measurement: newer entry points are simply absent
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = restype
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        raise TTBError(rc, load().ttb_last_error().decode('utf-8', 'replace'))
