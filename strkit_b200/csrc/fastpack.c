/* fastpack.c -- CPython helper of strkit_b200.batcher.pack_loci: copies the per-read strings of a block of loci
 * (what call_locus hands to get_repeat_count per read, call_locus.py:1144-1155) into the flat arrays of a ReadBatch
 * in two passes over the Python objects, without building intermediate Python lists or one big joined string.
 *
 *   pack(loci) -> (arena, seq_off, lens, est_cn, read_begin, motif_off, motif_len)   all bytearray
 *
 * `loci` is a sequence of objects with the attributes of batcher.LocusReads (motif, est_cn, tr_seqs,
 * flank_left_seqs, flank_right_seqs).  Layout = the one documented in batcher.ReadBatch: per read fl + tr + fr
 * contiguous at seq_off[r], motifs after all reads.  Host-side plumbing only: nothing here is on the GPU path, and
 * pack_loci falls back to its pure-Python body when this module is not built.
 */
#define PY_SSIZE_T_CLEAN
#include <Python.h>
#include <stdint.h>
#include <string.h>

typedef struct {
    PyObject *motif, *est, *tr, *fl, *fr; /* est/tr/fl/fr: PySequence_Fast results (new references) */
    Py_ssize_t n;
} Locus;

static void free_loci(Locus *v, Py_ssize_t n) {
    for (Py_ssize_t i = 0; i < n; ++i) {
        Py_XDECREF(v[i].motif);
        Py_XDECREF(v[i].est);
        Py_XDECREF(v[i].tr);
        Py_XDECREF(v[i].fl);
        Py_XDECREF(v[i].fr);
    }
    PyMem_Free(v);
}

static PyObject *fast_attr(PyObject *o, const char *name, const char *what) {
    PyObject *a = PyObject_GetAttrString(o, name);
    if (!a) return NULL;
    PyObject *f = PySequence_Fast(a, what);
    Py_DECREF(a);
    return f;
}

/* ASCII str -> (pointer, length); anything else is an error (the device alphabet is ASCII) */
static int ascii_view(PyObject *s, const char **p, Py_ssize_t *len) {
    if (!PyUnicode_Check(s)) {
        PyErr_SetString(PyExc_TypeError, "pack_loci: sequences must be str");
        return -1;
    }
    if (!PyUnicode_IS_ASCII(s)) {
        PyErr_SetString(PyExc_ValueError, "pack_loci: sequences must be ASCII");
        return -1;
    }
    *p = (const char *)PyUnicode_1BYTE_DATA(s);
    *len = PyUnicode_GET_LENGTH(s);
    return 0;
}

static PyObject *pack(PyObject *self, PyObject *arg) {
    (void)self;
    PyObject *seq = PySequence_Fast(arg, "pack_loci: loci must be a sequence");
    if (!seq) return NULL;
    const Py_ssize_t n_loci = PySequence_Fast_GET_SIZE(seq);
    Locus *loci = (Locus *)PyMem_Calloc((size_t)(n_loci ? n_loci : 1), sizeof(Locus));
    PyObject *res = NULL, *arena = NULL, *seq_off = NULL, *lens = NULL, *est = NULL, *rb = NULL, *moff = NULL, *mlen = NULL;
    if (!loci) {
        Py_DECREF(seq);
        return PyErr_NoMemory();
    }
    /* pass 1: sizes */
    Py_ssize_t n_reads = 0;
    int64_t seq_bytes = 0, motif_bytes = 0;
    for (Py_ssize_t i = 0; i < n_loci; ++i) {
        PyObject *lr = PySequence_Fast_GET_ITEM(seq, i);
        Locus *L = &loci[i];
        if (!(L->motif = PyObject_GetAttrString(lr, "motif"))) goto fail;
        if (!(L->est = fast_attr(lr, "est_cn", "est_cn must be a sequence"))) goto fail;
        if (!(L->tr = fast_attr(lr, "tr_seqs", "tr_seqs must be a sequence"))) goto fail;
        if (!(L->fl = fast_attr(lr, "flank_left_seqs", "flank_left_seqs must be a sequence"))) goto fail;
        if (!(L->fr = fast_attr(lr, "flank_right_seqs", "flank_right_seqs must be a sequence"))) goto fail;
        L->n = PySequence_Fast_GET_SIZE(L->tr);
        if (PySequence_Fast_GET_SIZE(L->est) != L->n || PySequence_Fast_GET_SIZE(L->fl) != L->n ||
            PySequence_Fast_GET_SIZE(L->fr) != L->n) {
            PyErr_SetString(PyExc_ValueError, "LocusReads: per-read sequences must have equal lengths");
            goto fail;
        }
        const char *p;
        Py_ssize_t len;
        if (ascii_view(L->motif, &p, &len)) goto fail;
        motif_bytes += len;
        for (Py_ssize_t r = 0; r < L->n; ++r) {
            if (ascii_view(PySequence_Fast_GET_ITEM(L->fl, r), &p, &len)) goto fail;
            seq_bytes += len;
            if (ascii_view(PySequence_Fast_GET_ITEM(L->tr, r), &p, &len)) goto fail;
            seq_bytes += len;
            if (ascii_view(PySequence_Fast_GET_ITEM(L->fr, r), &p, &len)) goto fail;
            seq_bytes += len;
        }
        n_reads += L->n;
    }
    arena = PyByteArray_FromStringAndSize(NULL, (Py_ssize_t)(seq_bytes + motif_bytes));
    seq_off = PyByteArray_FromStringAndSize(NULL, n_reads * 8);
    lens = PyByteArray_FromStringAndSize(NULL, n_reads * 12);
    est = PyByteArray_FromStringAndSize(NULL, n_reads * 4);
    rb = PyByteArray_FromStringAndSize(NULL, (n_loci + 1) * 8);
    moff = PyByteArray_FromStringAndSize(NULL, n_loci * 8);
    mlen = PyByteArray_FromStringAndSize(NULL, n_loci * 4);
    if (!arena || !seq_off || !lens || !est || !rb || !moff || !mlen) goto fail;
    {
        /* pass 2: copy */
        char *a = PyByteArray_AS_STRING(arena);
        uint64_t *so = (uint64_t *)PyByteArray_AS_STRING(seq_off);
        int32_t *ln = (int32_t *)PyByteArray_AS_STRING(lens);
        int32_t *ec = (int32_t *)PyByteArray_AS_STRING(est);
        int64_t *rbp = (int64_t *)PyByteArray_AS_STRING(rb);
        uint64_t *mo = (uint64_t *)PyByteArray_AS_STRING(moff);
        int32_t *ml = (int32_t *)PyByteArray_AS_STRING(mlen);
        int64_t at = 0, mat = seq_bytes;
        Py_ssize_t r_glob = 0;
        rbp[0] = 0;
        for (Py_ssize_t i = 0; i < n_loci; ++i) {
            Locus *L = &loci[i];
            const char *p;
            Py_ssize_t len;
            ascii_view(L->motif, &p, &len);
            memcpy(a + mat, p, (size_t)len);
            mo[i] = (uint64_t)mat;
            ml[i] = (int32_t)len;
            mat += len;
            for (Py_ssize_t r = 0; r < L->n; ++r, ++r_glob) {
                long e = PyLong_AsLong(PySequence_Fast_GET_ITEM(L->est, r));
                if (e == -1 && PyErr_Occurred()) goto fail;
                if (e < INT32_MIN || e > INT32_MAX) {
                    PyErr_SetString(PyExc_OverflowError, "pack_loci: est_cn out of int32 range");
                    goto fail;
                }
                ec[r_glob] = (int32_t)e;
                so[r_glob] = (uint64_t)at;
                PyObject *parts[3] = {PySequence_Fast_GET_ITEM(L->fl, r), PySequence_Fast_GET_ITEM(L->tr, r),
                                      PySequence_Fast_GET_ITEM(L->fr, r)};
                for (int k = 0; k < 3; ++k) {
                    ascii_view(parts[k], &p, &len);
                    if (len > INT32_MAX) {
                        PyErr_SetString(PyExc_OverflowError, "pack_loci: sequence too long");
                        goto fail;
                    }
                    memcpy(a + at, p, (size_t)len);
                    ln[3 * r_glob + k] = (int32_t)len;
                    at += len;
                }
            }
            rbp[i + 1] = (int64_t)r_glob;
        }
    }
    res = PyTuple_Pack(7, arena, seq_off, lens, est, rb, moff, mlen);
fail:
    Py_XDECREF(arena);
    Py_XDECREF(seq_off);
    Py_XDECREF(lens);
    Py_XDECREF(est);
    Py_XDECREF(rb);
    Py_XDECREF(moff);
    Py_XDECREF(mlen);
    free_loci(loci, n_loci);
    Py_DECREF(seq);
    return res;
}

static PyMethodDef methods[] = {{"pack", pack, METH_O, "pack(loci) -> 7 bytearrays (see fastpack.c)"}, {NULL, NULL, 0, NULL}};
static struct PyModuleDef moddef = {PyModuleDef_HEAD_INIT, "_fastpack", "C helper of strkit_b200.batcher.pack_loci", -1, methods,
                                    NULL, NULL, NULL, NULL};
PyMODINIT_FUNC PyInit__fastpack(void) { return PyModule_Create(&moddef); }
