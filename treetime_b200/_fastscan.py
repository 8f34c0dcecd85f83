"""C-speed scans over the tree's node objects for the per-pass host work of the mirror (csrc/ttb_fastscan.c, built
in-tree with gcc against the CPython headers; loaded with ctypes.PyDLL).  Optional: without the shared object the
callers use their numpy / map() forms -- this is host logic, not a device fallback."""
import ctypes
import os
import subprocess
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, 'csrc', 'ttb_fastscan.c')
LIB_PATH = os.path.join(HERE, '_ttb_fastscan.so')
_lib = None
_tried = False


def build(force=False):
    if not force and os.path.exists(LIB_PATH) and os.path.getmtime(LIB_PATH) >= os.path.getmtime(SRC):
        return LIB_PATH
    cmd = [os.environ.get('CC', 'gcc'), '-O2', '-shared', '-fPIC', '-I' + sysconfig.get_paths()['include'], SRC, '-o', LIB_PATH]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError('gcc failed:\n%s\n%s' % (' '.join(cmd), r.stderr))
    return LIB_PATH


def load():
    """The library, or None when it has not been built (or cannot be loaded with this interpreter)."""
    global _lib, _tried
    if _tried:
        return _lib
    _tried = True
    try:
        lib = ctypes.PyDLL(LIB_PATH)
        lib.ttb_scan_float_attr.argtypes = [ctypes.py_object, ctypes.py_object, ctypes.c_void_p, ctypes.c_ssize_t]
        lib.ttb_scan_float_attr.restype = ctypes.c_int
        lib.ttb_any_not_none.argtypes = [ctypes.py_object, ctypes.py_object]
        lib.ttb_any_not_none.restype = ctypes.c_int
        lib.ttb_scan_nodes.argtypes = [ctypes.py_object, ctypes.py_object, ctypes.py_object, ctypes.c_void_p, ctypes.c_ssize_t]
        lib.ttb_scan_nodes.restype = ctypes.c_int
        _lib = lib
    except (OSError, AttributeError):
        _lib = None
    return _lib


def scan_float_attr(dicts, key, out, start=0):
    """out[start:] = [float(d[key]) for d in dicts[start:]] into a float64 array; False if the fast path cannot do it."""
    lib = load()
    if lib is None or out.dtype.str != '<f8' or not out.flags.c_contiguous or out.shape[0] != len(dicts):
        return False
    return lib.ttb_scan_float_attr(dicts, key, out.ctypes.data, start) == 0


def any_not_none(dicts, key):
    """True / False, or None if the fast path cannot tell."""
    lib = load()
    if lib is None:
        return None
    r = lib.ttb_any_not_none(dicts, key)
    return None if r < 0 else bool(r)


def scan_nodes(dicts, float_key, mask_key, out, start=0):
    """Both per-pass scans in one cache-friendly walk: fills out[start:] like scan_float_attr and returns whether any
    dict has a non-None `mask_key` (True / False); None if the fast path cannot do it (nothing is guaranteed about `out`)."""
    lib = load()
    if lib is None or out.dtype.str != '<f8' or not out.flags.c_contiguous or out.shape[0] != len(dicts):
        return None
    r = lib.ttb_scan_nodes(dicts, float_key, mask_key, out.ctypes.data, start)
    return None if r < 0 else bool(r)

