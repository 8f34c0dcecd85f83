/* ttb_fastscan.c -- host-side helper of the Python mirror, NOT part of the C-ABI (include/ttb.h): per-pass scans over
 * the tree's node objects at C speed.  A pass needs every node's branch length (treeanc.py:752-760 reads
 * node.branch_length / node.mutation_length per node) and has to know whether any node carries a mask (arg.py:128-133);
 * on a 40 000-node tree the numpy.fromiter / map() forms of these scans cost 3.4 + 2 ms next to a 14 ms device pass.
 * Loaded with ctypes.PyDLL (the GIL is held); every function works on a list of the nodes' instance dicts and reports
 * "cannot" (-1) instead of guessing, so that the caller falls back to the generic Python path. */
#define PY_SSIZE_T_CLEAN
#include <Python.h>

/* out[i] = float(dicts[i][key]) for start <= i < len(dicts).  Returns 0, or -1 if an entry is missing / not a number. */
int ttb_scan_float_attr(PyObject* dicts, PyObject* key, double* out, Py_ssize_t start) {
  if (!PyList_CheckExact(dicts)) return -1;
  const Py_ssize_t n = PyList_GET_SIZE(dicts);
  for (Py_ssize_t i = start; i < n; ++i) {
    PyObject* d = PyList_GET_ITEM(dicts, i);
    if (!PyDict_CheckExact(d)) return -1;
    PyObject* v = PyDict_GetItemWithError(d, key); /* borrowed */
    if (!v) {
      PyErr_Clear();
      return -1;
    }
    if (PyFloat_Check(v))
      out[i] = PyFloat_AS_DOUBLE(v);
    else if (PyLong_Check(v)) {
      out[i] = PyLong_AsDouble(v);
      if (out[i] == -1.0 && PyErr_Occurred()) {
        PyErr_Clear();
        return -1;
      }
    } else
      return -1;
  }
  return 0;
}

/* 1 if dicts[i].get(key) is not None for some i, 0 if for none, -1 if the argument is not a list of dicts. */
int ttb_any_not_none(PyObject* dicts, PyObject* key) {
  if (!PyList_CheckExact(dicts)) return -1;
  const Py_ssize_t n = PyList_GET_SIZE(dicts);
  for (Py_ssize_t i = 0; i < n; ++i) {
    PyObject* d = PyList_GET_ITEM(dicts, i);
    if (!PyDict_CheckExact(d)) return -1;
    PyObject* v = PyDict_GetItemWithError(d, key);
    if (!v) {
      if (PyErr_Occurred()) {
        PyErr_Clear();
        return -1;
      }
      continue;
    }
    if (v != Py_None) return 1;
  }
  return 0;
}

/* Both scans in one walk over the dicts, software-pipelined against cache misses: on a 200 000-node tree the nodes' dicts,
 * their value arrays and the float objects are scattered over far more memory than the caches hold, and the two separate
 * walks above cost 25 + 20 ms of pointer chasing (40 000 nodes: 0.6 + 0.5 ms).  Here the dict objects are prefetched 48
 * entries ahead, their key / value tables 24 ahead, and the numbers are read in blocks of 32 after their objects have been
 * requested.  out[i] = float(dicts[i][key_f]) for i >= start; returns 1 / 0 = some / no dicts[i].get(key_m) is not None,
 * -1 = cannot (same rules as the single scans). */
#define TTB_SCAN_BLOCK 64
int ttb_scan_nodes(PyObject* dicts, PyObject* key_f, PyObject* key_m, double* out, Py_ssize_t start) {
  if (!PyList_CheckExact(dicts)) return -1;
  const Py_ssize_t n = PyList_GET_SIZE(dicts);
  int masked = 0;
  PyObject* vals[TTB_SCAN_BLOCK];
  for (Py_ssize_t b = 0; b < n; b += TTB_SCAN_BLOCK) {
    const Py_ssize_t e = b + TTB_SCAN_BLOCK < n ? b + TTB_SCAN_BLOCK : n;
    for (Py_ssize_t i = b; i < e; ++i) {
      if (i + 48 < n) __builtin_prefetch(PyList_GET_ITEM(dicts, i + 48));
      if (i + 24 < n) {
        PyObject* d8 = PyList_GET_ITEM(dicts, i + 24);
        if (PyDict_CheckExact(d8)) {
          /* a combined-table dict keeps indices + entries behind ma_keys (a few cache lines for ~15 attributes), a
           * split-table one its values behind ma_values */
          const char* kk = (const char*)((PyDictObject*)d8)->ma_keys;
          __builtin_prefetch(kk);
          __builtin_prefetch(kk + 64);
          __builtin_prefetch(kk + 128);
          __builtin_prefetch(kk + 192);
          __builtin_prefetch(kk + 256);
          __builtin_prefetch(((PyDictObject*)d8)->ma_values);
        }
      }
      PyObject* d = PyList_GET_ITEM(dicts, i);
      if (!PyDict_CheckExact(d)) return -1;
      PyObject* m = PyDict_GetItemWithError(d, key_m);
      if (!m) {
        if (PyErr_Occurred()) { PyErr_Clear(); return -1; }
      } else if (m != Py_None)
        masked = 1;
      PyObject* v = NULL;
      if (i >= start) {
        v = PyDict_GetItemWithError(d, key_f);
        if (!v) { PyErr_Clear(); return -1; }
        __builtin_prefetch(v);
      }
      vals[i - b] = v;
    }
    for (Py_ssize_t i = b; i < e; ++i) {
      PyObject* v = vals[i - b];
      if (!v) continue;
      if (PyFloat_Check(v))
        out[i] = PyFloat_AS_DOUBLE(v);
      else if (PyLong_Check(v)) {
        out[i] = PyLong_AsDouble(v);
        if (out[i] == -1.0 && PyErr_Occurred()) { PyErr_Clear(); return -1; }
      } else
        return -1;
    }
  }
  return masked;
}

