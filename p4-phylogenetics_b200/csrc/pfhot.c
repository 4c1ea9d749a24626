/* pfhot.c -- CPython bindings of the calls p4 issues thousands of times per evaluation.
 *
 * p4's Tree.setCStuff sends 4 calls per node before every likelihood calculation
 * (p4/tree.py:9338-9355) and Chain.proposeSp a few dozen more; through ctypes each costs
 * about half a microsecond of argument marshalling, which on a 125k-pattern shard is as long
 * as the kernels take.  These are the same wrappers as in pf.py -- same names, argument
 * orders and error behaviour (Pf/pfmodule.c:1907-2195, 2431-2491) -- as METH_FASTCALL
 * functions over the C ABI of include/p4b200.h.  pf.py installs them over its ctypes
 * versions when this module is present; nothing else changes.
 */
#define PY_SSIZE_T_CLEAN
#include <Python.h>
#include <stdio.h>

#include "../../include/p4b200.h"

static PyObject *g_fatal = NULL;   /* pf.P4bFatal */

static PyObject *fatal(void)
{
    const char *msg = p4b_lastError();
    if (!msg) msg = "p4b200: unknown error";
    PySys_WriteStdout("%s\n", msg);
    PyErr_SetString(g_fatal ? g_fatal : PyExc_RuntimeError, msg);
    return NULL;
}

static int as_ptr(PyObject *o, void **out)
{
    void *p = PyLong_AsVoidPtr(o);
    if (!p && PyErr_Occurred()) return -1;
    *out = p;
    return 0;
}
static int as_int(PyObject *o, int *out)
{
    long v;
    if (PyLong_Check(o)) v = PyLong_AsLong(o);
    else {
        PyObject *i = PyNumber_Index(o);   /* numpy integers */
        if (!i) return -1;
        v = PyLong_AsLong(i);
        Py_DECREF(i);
    }
    if (v == -1 && PyErr_Occurred()) return -1;
    *out = (int)v;
    return 0;
}
static int as_dbl(PyObject *o, double *out)
{
    const double v = PyFloat_AsDouble(o);
    if (v == -1.0 && PyErr_Occurred()) return -1;
    *out = v;
    return 0;
}
#define NARGS(n)                                                                   \
    if (nargs != (n)) {                                                            \
        PyErr_Format(PyExc_TypeError, "%s() takes %d arguments", __func__ + 3, (n)); \
        return NULL;                                                               \
    }
#define DONE(rc)               \
    if ((rc) != 0) return fatal(); \
    Py_RETURN_NONE

static PyObject *hf_p4_setNodeRelation(PyObject *self, PyObject *const *a, Py_ssize_t nargs)
{
    void *n; int rel, num;
    NARGS(3);
    if (as_ptr(a[0], &n) || as_int(a[1], &rel) || as_int(a[2], &num)) return NULL;
    DONE(p4b_setNodeRelation(n, rel, num));
}
static PyObject *hf_p4_setTreeRoot(PyObject *self, PyObject *const *a, Py_ssize_t nargs)
{
    void *t, *n;
    NARGS(2);
    if (as_ptr(a[0], &t) || as_ptr(a[1], &n)) return NULL;
    DONE(p4b_setTreeRoot(t, n));
}
static PyObject *hf_p4_setBrLen(PyObject *self, PyObject *const *a, Py_ssize_t nargs)
{
    void *n; double v;
    NARGS(2);
    if (as_ptr(a[0], &n) || as_dbl(a[1], &v)) return NULL;
    DONE(p4b_setBrLen(n, v));
}
#define SETNUM(NAME)                                                                        \
    static PyObject *hf_##NAME(PyObject *self, PyObject *const *a, Py_ssize_t nargs)        \
    {                                                                                       \
        void *n; int p, v;                                                                  \
        NARGS(3);                                                                           \
        if (as_ptr(a[0], &n) || as_int(a[1], &p) || as_int(a[2], &v)) return NULL;          \
        DONE(p4b_##NAME(n, p, v));                                                          \
    }
static PyObject *hf_p4_setCompNum(PyObject *self, PyObject *const *a, Py_ssize_t nargs)
{
    void *n; int p, v;
    NARGS(3);
    if (as_ptr(a[0], &n) || as_int(a[1], &p) || as_int(a[2], &v)) return NULL;
    DONE(p4b_setCompNum(n, p, v));
}
static PyObject *hf_p4_setRMatrixNum(PyObject *self, PyObject *const *a, Py_ssize_t nargs)
{
    void *n; int p, v;
    NARGS(3);
    if (as_ptr(a[0], &n) || as_int(a[1], &p) || as_int(a[2], &v)) return NULL;
    DONE(p4b_setRMatrixNum(n, p, v));
}
static PyObject *hf_p4_setGdasrvNum(PyObject *self, PyObject *const *a, Py_ssize_t nargs)
{
    void *n; int p, v;
    NARGS(3);
    if (as_ptr(a[0], &n) || as_int(a[1], &p) || as_int(a[2], &v)) return NULL;
    DONE(p4b_setGdasrvNum(n, p, v));
}
static PyObject *hf_p4_setRMatrixBigR(PyObject *self, PyObject *const *a, Py_ssize_t nargs)
{
    void *m; int p, r, i, j; double v;
    NARGS(6);
    if (as_ptr(a[0], &m) || as_int(a[1], &p) || as_int(a[2], &r) || as_int(a[3], &i) || as_int(a[4], &j) || as_dbl(a[5], &v)) return NULL;
    DONE(p4b_setRMatrixBigR(m, p, r, i, j, v));
}
static PyObject *hf_p4_setKappa(PyObject *self, PyObject *const *a, Py_ssize_t nargs)
{
    void *m; int p, r; double v;
    NARGS(4);
    if (as_ptr(a[0], &m) || as_int(a[1], &p) || as_int(a[2], &r) || as_dbl(a[3], &v)) return NULL;
    DONE(p4b_setKappa(m, p, r, v));
}
static PyObject *hf_p4_setPInvarVal(PyObject *self, PyObject *const *a, Py_ssize_t nargs)
{
    void *m; int p; double v;
    NARGS(3);
    if (as_ptr(a[0], &m) || as_int(a[1], &p) || as_dbl(a[2], &v)) return NULL;
    DONE(p4b_setPInvarVal(m, p, v));
}
static PyObject *hf_p4_setRelRateVal(PyObject *self, PyObject *const *a, Py_ssize_t nargs)
{
    void *m; int p; double v;
    NARGS(3);
    if (as_ptr(a[0], &m) || as_int(a[1], &p) || as_dbl(a[2], &v)) return NULL;
    DONE(p4b_setRelRateVal(m, p, v));
}
static PyObject *hf_p4_setPrams(PyObject *self, PyObject *const *a, Py_ssize_t nargs)
{
    void *t; int p;
    NARGS(2);
    if (as_ptr(a[0], &t) || as_int(a[1], &p)) return NULL;
    DONE(p4b_setPrams(t, p));
}
static PyObject *hf_p4_calculateBigPDecks(PyObject *self, PyObject *const *a, Py_ssize_t nargs)
{
    void *n;
    NARGS(1);
    if (as_ptr(a[0], &n)) return NULL;
    DONE(p4b_calculateBigPDecks(n));
}
static PyObject *hf_p4_setConditionalLikelihoodsOfInternalNodePart(PyObject *self, PyObject *const *a, Py_ssize_t nargs)
{
    void *n; int p;
    NARGS(2);
    if (as_ptr(a[0], &n) || as_int(a[1], &p)) return NULL;
    DONE(p4b_setConditionalLikelihoodsOfInternalNodePart(n, p));
}
static PyObject *hf_p4_partLogLike(PyObject *self, PyObject *const *a, Py_ssize_t nargs)
{
    void *t, *part; int p, sl; double v;
    NARGS(4);
    if (as_ptr(a[0], &t) || as_ptr(a[1], &part) || as_int(a[2], &p) || as_int(a[3], &sl)) return NULL;
    v = p4b_partLogLike(t, part, p, sl);
    if (v != v) return fatal();
    return PyFloat_FromDouble(v);
}
static PyObject *hf_p4_treeLogLike(PyObject *self, PyObject *const *a, Py_ssize_t nargs)
{
    void *t; int sl; double v;
    NARGS(2);
    if (as_ptr(a[0], &t) || as_int(a[1], &sl)) return NULL;
    v = p4b_treeLogLike(t, sl);
    if (v != v) {
        const char *msg = p4b_lastError();
        if (msg && msg[0]) return fatal();
    }
    return PyFloat_FromDouble(v);
}
static PyObject *hf_p4_copyCondLikes(PyObject *self, PyObject *const *a, Py_ssize_t nargs)
{
    void *x, *y; int all;
    NARGS(3);
    if (as_ptr(a[0], &x) || as_ptr(a[1], &y) || as_int(a[2], &all)) return NULL;
    DONE(p4b_copyCondLikes(x, y, all));
}
static PyObject *hf_p4_copyBigPDecks(PyObject *self, PyObject *const *a, Py_ssize_t nargs)
{
    void *x, *y; int all;
    NARGS(3);
    if (as_ptr(a[0], &x) || as_ptr(a[1], &y) || as_int(a[2], &all)) return NULL;
    DONE(p4b_copyBigPDecks(x, y, all));
}
static PyObject *hf_p4_copyModelPrams(PyObject *self, PyObject *const *a, Py_ssize_t nargs)
{
    void *x, *y;
    NARGS(2);
    if (as_ptr(a[0], &x) || as_ptr(a[1], &y)) return NULL;
    DONE(p4b_copyModelPrams(x, y));
}
static PyObject *hf_set_fatal(PyObject *self, PyObject *cls)
{
    Py_XDECREF(g_fatal);
    Py_INCREF(cls);
    g_fatal = cls;
    Py_RETURN_NONE;
}

#define F(NAME) {#NAME, (PyCFunction)(void (*)(void))hf_##NAME, METH_FASTCALL, "see p4_phylogenetics_b200.pf." #NAME}
static PyMethodDef methods[] = {
    F(p4_setNodeRelation), F(p4_setTreeRoot), F(p4_setBrLen), F(p4_setCompNum), F(p4_setRMatrixNum), F(p4_setGdasrvNum),
    F(p4_setRMatrixBigR), F(p4_setKappa), F(p4_setPInvarVal), F(p4_setRelRateVal), F(p4_setPrams),
    F(p4_calculateBigPDecks), F(p4_setConditionalLikelihoodsOfInternalNodePart), F(p4_partLogLike), F(p4_treeLogLike),
    F(p4_copyCondLikes), F(p4_copyBigPDecks), F(p4_copyModelPrams),
    {"set_fatal", hf_set_fatal, METH_O, "install the exception class raised on engine errors"},
    {NULL, NULL, 0, NULL}};
static struct PyModuleDef moddef = {PyModuleDef_HEAD_INIT, "_pfhot", "fast bindings of the hot pf calls", -1, methods};
PyMODINIT_FUNC PyInit__pfhot(void) { return PyModule_Create(&moddef); }
