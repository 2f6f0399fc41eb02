/* fastpack.c -- CPython helper of strkit_b200.batcher.pack_loci: copies the per-read strings of a block of loci
 * (what call_locus hands to get_repeat_count per read, call_locus.py:1144-1155) into the flat arrays of a ReadBatch.
 *
 *   pack(loci, nibble=0, threads=0) -> (arena, seq_off, lens, est_cn, read_begin, motif_off, motif_len)   all bytearray
 *
 * `loci` is a sequence of objects with the attributes of batcher.LocusReads (motif, est_cn, tr_seqs,
 * flank_left_seqs, flank_right_seqs); sequences are ASCII str or bytes.  Pass 1 (under the GIL) walks the Python
 * objects once and records raw pointers, lengths and offsets; pass 2 (GIL released, `threads` pthreads over ranges of
 * reads, 0 = one per core up to 16) moves the bytes -- a plain copy, or the nibble encoding of
 * batcher.ARENA_NIBBLE (two symbols per byte, low nibble first, code = index into "ACGTRYSWKMBDHVNX"; every read and
 * motif starts on a byte boundary, so threads never share a byte).  A byte outside the alphabet has no nibble code:
 * ValueError, and pack_loci falls back to the ASCII layout.  Layout = the one documented in batcher.ReadBatch: per
 * read fl + tr + fr contiguous at seq_off[r] (in symbols), motifs after all reads.  Host-side plumbing only.
 */
#define PY_SSIZE_T_CLEAN
#include <Python.h>
#include <pthread.h>
#include <stdint.h>
#include <string.h>
#include <unistd.h>

/* ASCII str or bytes -> (pointer, length) */
static int seq_view(PyObject *s, const char **p, Py_ssize_t *len) {
    if (PyUnicode_Check(s)) {
        if (!PyUnicode_IS_ASCII(s)) {
            PyErr_SetString(PyExc_ValueError, "pack_loci: sequences must be ASCII");
            return -1;
        }
        *p = (const char *)PyUnicode_1BYTE_DATA(s);
        *len = PyUnicode_GET_LENGTH(s);
        return 0;
    }
    if (PyBytes_Check(s)) {
        *p = PyBytes_AS_STRING(s);
        *len = PyBytes_GET_SIZE(s);
        return 0;
    }
    PyErr_SetString(PyExc_TypeError, "pack_loci: sequences must be str or bytes");
    return -1;
}

static PyObject *fast_attr(PyObject *o, const char *name, const char *what) {
    PyObject *a = PyObject_GetAttrString(o, name);
    if (!a) return NULL;
    PyObject *f = PySequence_Fast(a, what);
    Py_DECREF(a);
    return f;
}

typedef struct {
    const char *p;
    int32_t len;
    uint64_t off; /* symbol offset of the piece in the arena */
} Piece;

typedef struct {
    const Piece *pieces;
    Py_ssize_t begin, end; /* pieces [begin, end); a read's three pieces never straddle two jobs */
    char *arena;
    int nibble;
    int bad; /* out: a byte without a nibble code was seen */
} CopyJob;

static unsigned char g_code[256];
static void init_codes(void) {
    static const char alpha[] = "ACGTRYSWKMBDHVNX";
    memset(g_code, 16, sizeof(g_code));
    for (int i = 0; i < 16; ++i) {
        g_code[(unsigned char)alpha[i]] = (unsigned char)i;
        g_code[(unsigned char)(alpha[i] - 'A' + 'a')] = (unsigned char)i;
    }
}

static void *copy_worker(void *arg) {
    CopyJob *jb = (CopyJob *)arg;
    unsigned bad = 0;
    for (Py_ssize_t k = jb->begin; k < jb->end; ++k) {
        const Piece *pc = &jb->pieces[k];
        if (!jb->nibble) {
            memcpy(jb->arena + pc->off, pc->p, (size_t)pc->len);
            continue;
        }
        const unsigned char *src = (const unsigned char *)pc->p;
        unsigned char *dst = (unsigned char *)jb->arena;
        uint64_t at = pc->off;
        int32_t i = 0;
        if ((at & 1) && pc->len > 0) { /* finish the byte the previous piece of this read started */
            const unsigned c = g_code[src[0]];
            bad |= c;
            dst[at >> 1] |= (unsigned char)(c << 4);
            ++at, ++i;
        }
        for (; i + 1 < pc->len; i += 2, at += 2) {
            const unsigned c0 = g_code[src[i]], c1 = g_code[src[i + 1]];
            bad |= c0 | c1;
            dst[at >> 1] = (unsigned char)(c0 | (c1 << 4));
        }
        if (i < pc->len) {
            const unsigned c = g_code[src[i]];
            bad |= c;
            dst[at >> 1] = (unsigned char)c;
        }
    }
    jb->bad = (bad & 16u) != 0;
    return NULL;
}

static PyObject *pack(PyObject *self, PyObject *args, PyObject *kwargs) {
    (void)self;
    static char *kwlist[] = {"loci", "nibble", "threads", NULL};
    PyObject *arg = NULL;
    int nibble = 0, threads = 0;
    if (!PyArg_ParseTupleAndKeywords(args, kwargs, "O|pi", kwlist, &arg, &nibble, &threads)) return NULL;
    PyObject *seq = PySequence_Fast(arg, "pack_loci: loci must be a sequence");
    if (!seq) return NULL;
    const Py_ssize_t n_loci = PySequence_Fast_GET_SIZE(seq);
    PyObject **keep = (PyObject **)PyMem_Calloc((size_t)(5 * n_loci + 1), sizeof(PyObject *)); /* new references */
    PyObject *res = NULL, *arena = NULL, *seq_off = NULL, *lens = NULL, *est = NULL, *rb = NULL, *moff = NULL, *mlen = NULL;
    Piece *pieces = NULL;
    if (!keep) {
        Py_DECREF(seq);
        return PyErr_NoMemory();
    }
    /* count reads first (cheap), then one walk that fills everything but the arena */
    Py_ssize_t n_reads = 0;
    for (Py_ssize_t i = 0; i < n_loci; ++i) {
        PyObject *lr = PySequence_Fast_GET_ITEM(seq, i);
        PyObject **k = keep + 5 * i;
        if (!(k[0] = PyObject_GetAttrString(lr, "motif"))) goto fail;
        if (!(k[1] = fast_attr(lr, "est_cn", "est_cn must be a sequence"))) goto fail;
        if (!(k[2] = fast_attr(lr, "tr_seqs", "tr_seqs must be a sequence"))) goto fail;
        if (!(k[3] = fast_attr(lr, "flank_left_seqs", "flank_left_seqs must be a sequence"))) goto fail;
        if (!(k[4] = fast_attr(lr, "flank_right_seqs", "flank_right_seqs must be a sequence"))) goto fail;
        const Py_ssize_t n = PySequence_Fast_GET_SIZE(k[2]);
        if (PySequence_Fast_GET_SIZE(k[1]) != n || PySequence_Fast_GET_SIZE(k[3]) != n || PySequence_Fast_GET_SIZE(k[4]) != n) {
            PyErr_SetString(PyExc_ValueError, "LocusReads: per-read sequences must have equal lengths");
            goto fail;
        }
        n_reads += n;
    }
    seq_off = PyByteArray_FromStringAndSize(NULL, n_reads * 8);
    lens = PyByteArray_FromStringAndSize(NULL, n_reads * 12);
    est = PyByteArray_FromStringAndSize(NULL, n_reads * 4);
    rb = PyByteArray_FromStringAndSize(NULL, (n_loci + 1) * 8);
    moff = PyByteArray_FromStringAndSize(NULL, n_loci * 8);
    mlen = PyByteArray_FromStringAndSize(NULL, n_loci * 4);
    pieces = (Piece *)PyMem_Malloc(sizeof(Piece) * (size_t)(3 * n_reads + n_loci + 1));
    if (!seq_off || !lens || !est || !rb || !moff || !mlen || !pieces) {
        if (!PyErr_Occurred()) PyErr_NoMemory();
        goto fail;
    }
    uint64_t at = 0; /* symbols */
    {
        uint64_t *so = (uint64_t *)PyByteArray_AS_STRING(seq_off);
        int32_t *ln = (int32_t *)PyByteArray_AS_STRING(lens);
        int32_t *ec = (int32_t *)PyByteArray_AS_STRING(est);
        int64_t *rbp = (int64_t *)PyByteArray_AS_STRING(rb);
        uint64_t *mo = (uint64_t *)PyByteArray_AS_STRING(moff);
        int32_t *ml = (int32_t *)PyByteArray_AS_STRING(mlen);
        Py_ssize_t r_glob = 0;
        rbp[0] = 0;
        for (Py_ssize_t i = 0; i < n_loci; ++i) {
            PyObject **k = keep + 5 * i;
            const Py_ssize_t n = PySequence_Fast_GET_SIZE(k[2]);
            PyObject **e_items = PySequence_Fast_ITEMS(k[1]), **tr_items = PySequence_Fast_ITEMS(k[2]);
            PyObject **fl_items = PySequence_Fast_ITEMS(k[3]), **fr_items = PySequence_Fast_ITEMS(k[4]);
            for (Py_ssize_t r = 0; r < n; ++r, ++r_glob) {
                /* every str is its own heap object: the walk is bound by cache misses on the object headers, so the
                 * headers of a few reads ahead are prefetched while this one is recorded */
                if (r + 6 < n) {
                    __builtin_prefetch(fl_items[r + 6]);
                    __builtin_prefetch(tr_items[r + 6]);
                    __builtin_prefetch(fr_items[r + 6]);
                    __builtin_prefetch(e_items[r + 6]);
                }
                const long e = PyLong_AsLong(e_items[r]);
                if (e == -1 && PyErr_Occurred()) goto fail;
                if (e < INT32_MIN || e > INT32_MAX) {
                    PyErr_SetString(PyExc_OverflowError, "pack_loci: est_cn out of int32 range");
                    goto fail;
                }
                ec[r_glob] = (int32_t)e;
                so[r_glob] = at;
                PyObject *parts[3] = {fl_items[r], tr_items[r], fr_items[r]};
                for (int q = 0; q < 3; ++q) {
                    const char *p;
                    Py_ssize_t len;
                    if (seq_view(parts[q], &p, &len)) goto fail;
                    if (len > INT32_MAX) {
                        PyErr_SetString(PyExc_OverflowError, "pack_loci: sequence too long");
                        goto fail;
                    }
                    Piece *pc = &pieces[3 * r_glob + q];
                    pc->p = p, pc->len = (int32_t)len, pc->off = at;
                    ln[3 * r_glob + q] = (int32_t)len;
                    at += (uint64_t)len;
                }
                if (nibble) at += at & 1; /* next read starts on a byte boundary */
            }
            rbp[i + 1] = (int64_t)r_glob;
        }
        for (Py_ssize_t i = 0; i < n_loci; ++i) { /* motifs after all reads */
            const char *p;
            Py_ssize_t len;
            if (seq_view(keep[5 * i], &p, &len)) goto fail;
            if (len > INT32_MAX) {
                PyErr_SetString(PyExc_OverflowError, "pack_loci: motif too long");
                goto fail;
            }
            Piece *pc = &pieces[3 * n_reads + i];
            pc->p = p, pc->len = (int32_t)len, pc->off = at;
            mo[i] = at;
            ml[i] = (int32_t)len;
            at += (uint64_t)len;
            if (nibble) at += at & 1;
        }
    }
    arena = PyByteArray_FromStringAndSize(NULL, (Py_ssize_t)(nibble ? (at + 1) / 2 : at));
    if (!arena) goto fail;
    {
        /* pass 2: move the bytes, GIL released (the str / bytes objects are kept alive by `keep` and `seq`) */
        const Py_ssize_t n_pieces = 3 * n_reads + n_loci;
        long ncpu = sysconf(_SC_NPROCESSORS_ONLN);
        int nt = threads > 0 ? threads : (int)(ncpu < 1 ? 1 : (ncpu > 16 ? 16 : ncpu));
        if ((Py_ssize_t)nt > n_reads / 4096 + 1) nt = (int)(n_reads / 4096 + 1);
        CopyJob jobs[64];
        pthread_t th[64];
        if (nt > 64) nt = 64;
        char *a = PyByteArray_AS_STRING(arena);
        for (int t = 0; t < nt; ++t) {
            const Py_ssize_t r0 = n_reads * t / nt, r1 = n_reads * (t + 1) / nt;
            jobs[t].pieces = pieces;
            jobs[t].begin = 3 * r0;
            jobs[t].end = t == nt - 1 ? n_pieces : 3 * r1; /* the last job also takes the motifs */
            jobs[t].arena = a;
            jobs[t].nibble = nibble;
            jobs[t].bad = 0;
        }
        int bad = 0;
        Py_BEGIN_ALLOW_THREADS
        if (nt == 1) {
            copy_worker(&jobs[0]);
        } else {
            int started = 0;
            for (int t = 0; t < nt; ++t) {
                if (pthread_create(&th[t], NULL, copy_worker, &jobs[t]) != 0) break;
                ++started;
            }
            for (int t = started; t < nt; ++t) copy_worker(&jobs[t]); /* could not spawn: do it here */
            for (int t = 0; t < started; ++t) pthread_join(th[t], NULL);
        }
        Py_END_ALLOW_THREADS
        for (int t = 0; t < nt; ++t) bad |= jobs[t].bad;
        if (bad) {
            PyErr_SetString(PyExc_ValueError, "pack_loci: a byte outside ACGTRYSWKMBDHVNX has no nibble code");
            goto fail;
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
    PyMem_Free(pieces);
    for (Py_ssize_t i = 0; i < 5 * n_loci; ++i) Py_XDECREF(keep[i]);
    PyMem_Free(keep);
    Py_DECREF(seq);
    return res;
}

static PyMethodDef methods[] = {{"pack", (PyCFunction)(void (*)(void))pack, METH_VARARGS | METH_KEYWORDS,
                                 "pack(loci, nibble=False, threads=0) -> 7 bytearrays (see fastpack.c)"},
                                {NULL, NULL, 0, NULL}};
static struct PyModuleDef moddef = {PyModuleDef_HEAD_INIT, "_fastpack", "C helper of strkit_b200.batcher.pack_loci", -1, methods,
                                    NULL, NULL, NULL, NULL};
PyMODINIT_FUNC PyInit__fastpack(void) {
    init_codes();
    return PyModule_Create(&moddef);
}
