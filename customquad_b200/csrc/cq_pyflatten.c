/*
 * cq_pyflatten.c -- host-side helper of customquad_b200/quadrature.py (CPython extension `_cq_flatten`).
 *
 * The reference hands its assemblers LISTS of per-cell arrays (demo/poisson.py:109-113: qr_pts[k], qr_w[k], qr_n[k],
 * read one by one inside the numba loop, assemble_matrix.py:84-101).  The device path wants one flat array per
 * quantity.  Doing that with a Python-level loop + np.concatenate over millions of arrays takes seconds; here it is
 * one pass over the list through the buffer protocol:
 *
 *   flatten(seq, out, lens) -> total number of doubles
 *     seq : list / tuple of C-contiguous float64 buffers (numpy arrays of any shape)
 *     out : writable float64 buffer receiving the concatenation, or None (count only)
 *     lens: writable int64 buffer [len(seq)] receiving the number of doubles of every item, or None
 *   Raises TypeError for an item that is not a contiguous float64 buffer (the caller then uses the numpy path).
 */
#define PY_SSIZE_T_CLEAN
#include <Python.h>
#include <string.h>
#define NPY_NO_DEPRECATED_API NPY_1_7_API_VERSION
#include <numpy/arrayobject.h>

static PyObject* flatten(PyObject* self, PyObject* args) {
  PyObject *seq, *out_obj, *lens_obj;
  if (!PyArg_ParseTuple(args, "OOO", &seq, &out_obj, &lens_obj)) return NULL;
  PyObject* fast = PySequence_Fast(seq, "flatten: first argument must be a sequence");
  if (!fast) return NULL;
  const Py_ssize_t n = PySequence_Fast_GET_SIZE(fast);
  Py_buffer out = {0}, lens = {0};
  int have_out = 0, have_lens = 0;
  if (out_obj != Py_None) {
    if (PyObject_GetBuffer(out_obj, &out, PyBUF_WRITABLE | PyBUF_C_CONTIGUOUS) < 0) { Py_DECREF(fast); return NULL; }
    have_out = 1;
  }
  if (lens_obj != Py_None) {
    if (PyObject_GetBuffer(lens_obj, &lens, PyBUF_WRITABLE | PyBUF_C_CONTIGUOUS) < 0) goto fail;
    have_lens = 1;
    if (lens.len < (Py_ssize_t)(n * 8)) { PyErr_SetString(PyExc_ValueError, "flatten: lens buffer too small"); goto fail; }
  }
  {
    Py_ssize_t at = 0; /* bytes */
    PyObject** items = PySequence_Fast_ITEMS(fast);
    for (Py_ssize_t k = 0; k < n; ++k) {
      if (PyArray_Check(items[k])) {
        /* numpy arrays (the usual case): read the array struct directly -- the buffer protocol costs ~300 ns per item */
        PyArrayObject* a = (PyArrayObject*)items[k];
        if (PyArray_TYPE(a) != NPY_DOUBLE || !PyArray_IS_C_CONTIGUOUS(a) || !PyArray_ISNOTSWAPPED(a)) {
          PyErr_Format(PyExc_TypeError, "flatten: item %zd is not a C-contiguous float64 array", k);
          goto fail;
        }
        const Py_ssize_t nb = (Py_ssize_t)PyArray_NBYTES(a);
        if (have_out) {
          if (at + nb > out.len) { PyErr_SetString(PyExc_ValueError, "flatten: output buffer too small"); goto fail; }
          memcpy((char*)out.buf + at, PyArray_DATA(a), (size_t)nb);
        }
        if (have_lens) ((long long*)lens.buf)[k] = (long long)(nb / 8);
        at += nb;
        continue;
      }
      Py_buffer b;
      if (PyObject_GetBuffer(items[k], &b, PyBUF_C_CONTIGUOUS | PyBUF_FORMAT) < 0) {
        PyErr_Clear();
        PyErr_Format(PyExc_TypeError, "flatten: item %zd is not a C-contiguous buffer", k);
        goto fail;
      }
      const int is_f64 = b.itemsize == 8 && b.format && (strcmp(b.format, "d") == 0 || strcmp(b.format, "<d") == 0 || strcmp(b.format, "=d") == 0);
      if (!is_f64) {
        PyBuffer_Release(&b);
        PyErr_Format(PyExc_TypeError, "flatten: item %zd is not float64", k);
        goto fail;
      }
      if (have_out) {
        if (at + b.len > out.len) {
          PyBuffer_Release(&b);
          PyErr_SetString(PyExc_ValueError, "flatten: output buffer too small");
          goto fail;
        }
        memcpy((char*)out.buf + at, b.buf, (size_t)b.len);
      }
      if (have_lens) ((long long*)lens.buf)[k] = (long long)(b.len / 8);
      at += b.len;
      PyBuffer_Release(&b);
    }
    if (have_out) PyBuffer_Release(&out);
    if (have_lens) PyBuffer_Release(&lens);
    Py_DECREF(fast);
    return PyLong_FromSsize_t(at / 8);
  }
fail:
  if (have_out) PyBuffer_Release(&out);
  if (have_lens) PyBuffer_Release(&lens);
  Py_DECREF(fast);
  return NULL;
}

static PyMethodDef methods[] = {{"flatten", flatten, METH_VARARGS, "concatenate a sequence of float64 buffers"}, {NULL, NULL, 0, NULL}};
static struct PyModuleDef moddef = {PyModuleDef_HEAD_INIT, "_cq_flatten", NULL, -1, methods};
PyMODINIT_FUNC PyInit__cq_flatten(void) {
  import_array();
  return PyModule_Create(&moddef);
}
