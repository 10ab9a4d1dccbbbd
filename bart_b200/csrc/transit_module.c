/* transit_module.c -- CPython extension exposing libbart_b200 with the Python surface of the
 * reference's SWIG module (modules/transit/transit/src/transit.i:12-31,52-86):
 *
 *   transit_init(argc:int, argv:list[str|bytes])     TypeError on a non-list / non-string item
 *   get_no_samples() -> int
 *   get_waveno_arr(n:int) -> ndarray[n] float64      (ARGOUT_ARRAY1)
 *   set_radius(float); set_cloudtop(float); set_scattering(int, float)
 *   run_transit(profiles: 1-D float64 array-like, nwave:int) -> ndarray[nwave] float64
 *   free_memory()
 *
 * so that code/BARTfunc.py:28-30,229-234,350-363,406 drives it unchanged.  Additive:
 *   run_transit_batch(profiles[M, n_in]) -> (spectra[M, nwave], status[M])
 *   set_filters(start, count, weight, star|None, rprs); band_flux_batch(profiles) -> (flux, status)
 * Errors of the library surface as RuntimeError (the reference exit()s the interpreter).
 */
#define PY_SSIZE_T_CLEAN
#include <Python.h>
#define NPY_NO_DEPRECATED_API NPY_1_7_API_VERSION
#include <numpy/arrayobject.h>
#include "bart_b200.h"

static int check_error(void) {
  if (bart_error_pending()) {
    PyErr_SetString(PyExc_RuntimeError, bart_last_error());
    bart_clear_error();
    return -1;
  }
  return 0;
}

static PyObject *py_transit_init(PyObject *self, PyObject *args) {
  int argc;
  PyObject *list;
  if (!PyArg_ParseTuple(args, "iO", &argc, &list)) return NULL;
  if (!PyList_Check(list)) { PyErr_SetString(PyExc_TypeError, "not a list"); return NULL; }
  Py_ssize_t n = PyList_Size(list);
  char **argv = (char **)malloc((n + 1) * sizeof(char *));
  PyObject **keep = (PyObject **)calloc(n + 1, sizeof(PyObject *));
  for (Py_ssize_t i = 0; i < n; i++) {
    PyObject *o = PyList_GetItem(list, i);
    if (PyBytes_Check(o)) argv[i] = PyBytes_AsString(o);
    else if (PyUnicode_Check(o)) {
      keep[i] = PyUnicode_AsASCIIString(o);
      if (!keep[i]) { free(argv); free(keep); return NULL; }
      argv[i] = PyBytes_AsString(keep[i]);
    } else {
      PyErr_SetString(PyExc_TypeError, "List must contain strings");
      for (Py_ssize_t k = 0; k < i; k++) Py_XDECREF(keep[k]);
      free(argv); free(keep);
      return NULL;
    }
  }
  argv[n] = NULL;
  if (argc > n) argc = (int)n;
  Py_BEGIN_ALLOW_THREADS
  transit_init(argc, argv);
  Py_END_ALLOW_THREADS
  for (Py_ssize_t k = 0; k < n; k++) Py_XDECREF(keep[k]);
  free(argv); free(keep);
  if (check_error()) return NULL;
  Py_RETURN_NONE;
}

static PyObject *py_get_no_samples(PyObject *self, PyObject *args) {
  return PyLong_FromLong(get_no_samples());
}

static PyObject *py_get_waveno_arr(PyObject *self, PyObject *args) {
  int n;
  if (!PyArg_ParseTuple(args, "i", &n)) return NULL;
  if (n < 0) { PyErr_SetString(PyExc_ValueError, "Array dimension must be non-negative"); return NULL; }
  npy_intp dims[1] = {n};
  PyObject *arr = PyArray_ZEROS(1, dims, NPY_DOUBLE, 0);
  if (!arr) return NULL;
  get_waveno_arr((double *)PyArray_DATA((PyArrayObject *)arr), n);
  return arr;
}

static PyObject *py_set_radius(PyObject *self, PyObject *args) {
  double r;
  if (!PyArg_ParseTuple(args, "d", &r)) return NULL;
  set_radius(r);
  Py_RETURN_NONE;
}

static PyObject *py_set_cloudtop(PyObject *self, PyObject *args) {
  double r;
  if (!PyArg_ParseTuple(args, "d", &r)) return NULL;
  set_cloudtop(r);
  Py_RETURN_NONE;
}

static PyObject *py_set_scattering(PyObject *self, PyObject *args) {
  int flag; double v;
  if (!PyArg_ParseTuple(args, "id", &flag, &v)) return NULL;
  set_scattering(flag, v);
  Py_RETURN_NONE;
}

static PyObject *py_run_transit(PyObject *self, PyObject *args) {
  PyObject *in; int nwave;
  if (!PyArg_ParseTuple(args, "Oi", &in, &nwave)) return NULL;
  PyArrayObject *a = (PyArrayObject *)PyArray_FROM_OTF(in, NPY_DOUBLE, NPY_ARRAY_IN_ARRAY);
  if (!a) return NULL;
  if (PyArray_NDIM(a) != 1) {
    Py_DECREF(a);
    PyErr_SetString(PyExc_TypeError, "Array must have 1 dimensions.");
    return NULL;
  }
  npy_intp dims[1] = {nwave};
  PyObject *out = PyArray_ZEROS(1, dims, NPY_DOUBLE, 0);
  if (!out) { Py_DECREF(a); return NULL; }
  double *pin = (double *)PyArray_DATA(a), *pout = (double *)PyArray_DATA((PyArrayObject *)out);
  int n_in = (int)PyArray_DIM(a, 0);
  Py_BEGIN_ALLOW_THREADS
  run_transit(pin, n_in, pout, nwave);
  Py_END_ALLOW_THREADS
  Py_DECREF(a);
  if (check_error()) { Py_DECREF(out); return NULL; }
  return out;
}

static PyObject *py_free_memory(PyObject *self, PyObject *args) {
  free_memory();
  if (check_error()) return NULL;
  Py_RETURN_NONE;
}

/* ---- additive ---- */
static PyObject *py_run_transit_batch(PyObject *self, PyObject *args) {
  PyObject *in;
  if (!PyArg_ParseTuple(args, "O", &in)) return NULL;
  PyArrayObject *a = (PyArrayObject *)PyArray_FROM_OTF(in, NPY_DOUBLE, NPY_ARRAY_IN_ARRAY);
  if (!a) return NULL;
  if (PyArray_NDIM(a) != 2) { Py_DECREF(a); PyErr_SetString(PyExc_TypeError, "profiles must be 2-D [models, values]"); return NULL; }
  int M = (int)PyArray_DIM(a, 0), n_in = (int)PyArray_DIM(a, 1), nw = get_no_samples();
  npy_intp d2[2] = {M, nw}, d1[1] = {M};
  PyObject *spec = PyArray_ZEROS(2, d2, NPY_DOUBLE, 0);
  PyObject *st = PyArray_ZEROS(1, d1, NPY_INT32, 0);
  if (!spec || !st) { Py_DECREF(a); Py_XDECREF(spec); Py_XDECREF(st); return NULL; }
  double *pin = (double *)PyArray_DATA(a), *ps = (double *)PyArray_DATA((PyArrayObject *)spec);
  int *pst = (int *)PyArray_DATA((PyArrayObject *)st);
  Py_BEGIN_ALLOW_THREADS
  bart_run_batch(pin, M, n_in, ps, nw, pst);
  Py_END_ALLOW_THREADS
  Py_DECREF(a);
  if (check_error()) { Py_DECREF(spec); Py_DECREF(st); return NULL; }
  return Py_BuildValue("NN", spec, st);
}

static PyObject *py_set_filters(PyObject *self, PyObject *args) {
  PyObject *ostart, *ocount, *oweight, *ostar; double rprs;
  if (!PyArg_ParseTuple(args, "OOOOd", &ostart, &ocount, &oweight, &ostar, &rprs)) return NULL;
  PyArrayObject *s = (PyArrayObject *)PyArray_FROM_OTF(ostart, NPY_INT32, NPY_ARRAY_IN_ARRAY);
  PyArrayObject *c = (PyArrayObject *)PyArray_FROM_OTF(ocount, NPY_INT32, NPY_ARRAY_IN_ARRAY);
  PyArrayObject *w = (PyArrayObject *)PyArray_FROM_OTF(oweight, NPY_DOUBLE, NPY_ARRAY_IN_ARRAY);
  PyArrayObject *st = NULL;
  if (ostar != Py_None) st = (PyArrayObject *)PyArray_FROM_OTF(ostar, NPY_DOUBLE, NPY_ARRAY_IN_ARRAY);
  if (!s || !c || !w || (ostar != Py_None && !st)) { Py_XDECREF(s); Py_XDECREF(c); Py_XDECREF(w); Py_XDECREF(st); return NULL; }
  /* the library trusts the lengths: count[] as long as start[], weight[] / star[] as long as
     the counts add up to */
  {
    const char *bad = NULL;
    long long total = 0;
    if (PyArray_SIZE(c) != PyArray_SIZE(s)) bad = "start and count differ in length";
    else {
      const int *pc = (const int *)PyArray_DATA(c);
      for (npy_intp i = 0; i < PyArray_SIZE(c); i++) { if (pc[i] < 0) bad = "negative count"; total += pc[i]; }
      if (!bad && PyArray_SIZE(w) != total) bad = "weight must hold sum(count) values";
      if (!bad && st && PyArray_SIZE(st) != total) bad = "star must hold sum(count) values";
    }
    if (bad) {
      Py_DECREF(s); Py_DECREF(c); Py_DECREF(w); Py_XDECREF(st);
      PyErr_SetString(PyExc_ValueError, bad);
      return NULL;
    }
  }
  bart_set_filters((int)PyArray_SIZE(s), (int *)PyArray_DATA(s), (int *)PyArray_DATA(c),
                   (double *)PyArray_DATA(w), st ? (double *)PyArray_DATA(st) : NULL, rprs);
  Py_DECREF(s); Py_DECREF(c); Py_DECREF(w); Py_XDECREF(st);
  if (check_error()) return NULL;
  Py_RETURN_NONE;
}

static PyObject *py_band_flux_batch(PyObject *self, PyObject *args) {
  PyObject *in; int nfilt;
  if (!PyArg_ParseTuple(args, "Oi", &in, &nfilt)) return NULL;
  PyArrayObject *a = (PyArrayObject *)PyArray_FROM_OTF(in, NPY_DOUBLE, NPY_ARRAY_IN_ARRAY);
  if (!a) return NULL;
  if (PyArray_NDIM(a) != 2) { Py_DECREF(a); PyErr_SetString(PyExc_TypeError, "profiles must be 2-D"); return NULL; }
  int M = (int)PyArray_DIM(a, 0), n_in = (int)PyArray_DIM(a, 1);
  /* the library writes M x (configured filters) doubles: the caller's count must be that one */
  if (nfilt != bart_nfilters()) {
    Py_DECREF(a);
    PyErr_Format(PyExc_ValueError, "band_flux_batch: nfilters %d given, %d configured by set_filters",
                 nfilt, bart_nfilters());
    return NULL;
  }
  npy_intp d2[2] = {M, nfilt}, d1[1] = {M};
  PyObject *bf = PyArray_ZEROS(2, d2, NPY_DOUBLE, 0);
  PyObject *st = PyArray_ZEROS(1, d1, NPY_INT32, 0);
  if (!bf || !st) { Py_DECREF(a); Py_XDECREF(bf); Py_XDECREF(st); return NULL; }
  double *pin = (double *)PyArray_DATA(a), *pb = (double *)PyArray_DATA((PyArrayObject *)bf);
  int *pst = (int *)PyArray_DATA((PyArrayObject *)st);
  Py_BEGIN_ALLOW_THREADS
  bart_bandflux_batch(pin, M, n_in, pb, pst);
  Py_END_ALLOW_THREADS
  Py_DECREF(a);
  if (check_error()) { Py_DECREF(bf); Py_DECREF(st); return NULL; }
  return Py_BuildValue("NN", bf, st);
}

static PyMethodDef methods[] = {
  {"transit_init", py_transit_init, METH_VARARGS, "transit_init(argc, argv)"},
  {"get_no_samples", py_get_no_samples, METH_NOARGS, "get_no_samples() -> int"},
  {"get_waveno_arr", py_get_waveno_arr, METH_VARARGS, "get_waveno_arr(n) -> ndarray"},
  {"set_radius", py_set_radius, METH_VARARGS, "set_radius(refradius)"},
  {"set_cloudtop", py_set_cloudtop, METH_VARARGS, "set_cloudtop(log10 bar)"},
  {"set_scattering", py_set_scattering, METH_VARARGS, "set_scattering(flag, logext)"},
  {"run_transit", py_run_transit, METH_VARARGS, "run_transit(profiles, nwave) -> ndarray"},
  {"free_memory", py_free_memory, METH_NOARGS, "free_memory()"},
  {"run_transit_batch", py_run_transit_batch, METH_VARARGS, "run_transit_batch(profiles[M,n]) -> (spectra, status)"},
  {"set_filters", py_set_filters, METH_VARARGS, "set_filters(start, count, weight, star|None, rprs)"},
  {"band_flux_batch", py_band_flux_batch, METH_VARARGS, "band_flux_batch(profiles[M,n], nfilters) -> (bandflux, status)"},
  {NULL, NULL, 0, NULL}
};

static struct PyModuleDef moddef = {PyModuleDef_HEAD_INIT, "transit_module",
  "B200-native drop-in for BART's transit_module (SWIG surface of transit.i)", -1, methods};

PyMODINIT_FUNC PyInit_transit_module(void) {
  import_array();
  bart_set_error_mode(1);
  return PyModule_Create(&moddef);
}
