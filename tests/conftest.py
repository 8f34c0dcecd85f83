import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'oracle'), os.path.join(ROOT, 'tests')):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')
    config.addinivalue_line('markers', 'reference: needs the unmodified reference under /root/reference (build container only)')


def _cuda_usable():
    """True if libttb.so loads and sees a CUDA device (cudaGetDeviceCount through ttb_create's own check)."""
    try:
        import ctypes
        from treetime_b200 import _lib
        lib = _lib.load()
        h = ctypes.c_void_p()
        rc = lib.ttb_create(ctypes.byref(h), 0, 5)
        if rc == 0:
            lib.ttb_destroy(h)
        return rc == 0
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    """A plain `pytest tests` on a host without a usable CUDA device skips the gpu-marked tests instead of erroring;
    an explicit `-m gpu` run keeps them (there a missing device must fail loudly, not skip)."""
    import pytest
    if 'gpu' in (config.getoption('-m') or ''):
        return
    gpu_items = [it for it in items if it.get_closest_marker('gpu')]
    if gpu_items and not _cuda_usable():
        skip = pytest.mark.skip(reason='no usable CUDA device')
        for it in gpu_items:
            it.add_marker(skip)
