/* fastpack.c -- CPython helper of strkit_b200.batcher.pack_loci: copies the per-read strings of a block of loci
 * (what call_locus hands to get_repeat_count per read, call_locus.py:1144-1155) into the flat arrays of a ReadBatch.
 *
 *   pack(loci, nibble=0, threads=0) -> (arena, seq_off, lens, est_cn, read_begin, motif_off, motif_len)   all bytearray
 *
 * `loci` is a sequence of objects with the attributes of batcher.LocusReads (motif, est_cn, tr_seqs,
 * flank_left_seqs, flank_right_seqs); sequences are ASCII str or bytes.
 *   pass 0 (under the GIL): per LOCUS, fetch the five attributes and take references to the sequences; per read, the
 *           start estimate (small ints: interpreter singletons, cache-resident).
 *   pass 1 (GIL released, `threads` pthreads over ranges of loci balanced by reads; 0 = one per core up to 16):
 *           A. every thread measures the three strings of each of its reads (lens[], per-thread symbol totals);
 *           B. with the totals prefixed into arena offsets, every thread writes seq_off[] and moves the bytes --
 *              a plain copy, or the nibble encoding of batcher.ARENA_NIBBLE (two symbols per byte, low nibble first,
 *              code = index into "ACGTRYSWKMBDHVNX"; every read and motif starts on a byte boundary, so threads never
 *              share a byte).
 *           The walk over the str objects is the expensive part (each is its own heap object: three cache misses per
 *           read), so it is the part that is threaded.  The threads only READ immutable objects (type, length, the
 *           inline character data of a compact ASCII str / of a bytes) that pass 0 keeps alive through the lists that
 *           hold them; they take no references and call no API that can raise.  As with any buffer handed to a
 *           GIL-free section, the caller must not mutate the lists while the call runs.
 * A byte outside the alphabet has no nibble code: ValueError, and pack_loci falls back to the ASCII layout.  Layout =
 * the one documented in batcher.ReadBatch: per read fl + tr + fr contiguous at seq_off[r] (in symbols), motifs after
 * all reads.  Host-side plumbing only.
 */
#define PY_SSIZE_T_CLEAN
#include <Python.h>
#include <pthread.h>
#include <stdint.h>
#include <string.h>
#include <unistd.h>

/* ASCII str or bytes -> (pointer, length); no Python API that can raise or allocate: usable without the GIL.
 * returns 0 ok, 1 not ASCII, 2 not str / bytes, 3 too long */
static inline int seq_view_nogil(PyObject *s, const char **p, Py_ssize_t *len) {
    if (PyUnicode_Check(s)) {
        if (!PyUnicode_IS_ASCII(s)) return 1;
        *p = (const char *)PyUnicode_1BYTE_DATA(s);
        *len = PyUnicode_GET_LENGTH(s);
    } else if (PyBytes_Check(s)) {
        *p = PyBytes_AS_STRING(s);
        *len = PyBytes_GET_SIZE(s);
    } else {
        return 2;
    }
    return *len > INT32_MAX ? 3 : 0;
}

static void set_view_error(int code, const char *what) {
    if (code == 1)
        PyErr_Format(PyExc_ValueError, "pack_loci: %s must be ASCII", what);
    else if (code == 2)
        PyErr_Format(PyExc_TypeError, "pack_loci: %s must be str or bytes", what);
    else
        PyErr_Format(PyExc_OverflowError, "pack_loci: %s too long", what);
}

/* attribute names, interned once at module init (PyObject_GetAttrString builds a str per call) */
static PyObject *g_names[5];

static PyObject *fast_attr(PyObject *o, PyObject *name, const char *what) {
    PyObject *a = PyObject_GetAttr(o, name);
    if (!a) return NULL;
    PyObject *f = PySequence_Fast(a, what);
    Py_DECREF(a);
    return f;
}

static unsigned char g_code[256];
static void init_codes(void) {
    static const char alpha[] = "ACGTRYSWKMBDHVNX";
    memset(g_code, 16, sizeof(g_code));
    for (int i = 0; i < 16; ++i) {
        g_code[(unsigned char)alpha[i]] = (unsigned char)i;
        g_code[(unsigned char)(alpha[i] - 'A' + 'a')] = (unsigned char)i;
    }
}

/* one piece into the arena at symbol offset `at`; returns the OR of the nibble codes seen (bit 4 = no code) */
static inline unsigned put_piece(char *arena, uint64_t at, const char *p, int32_t len, int nibble) {
    if (!nibble) {
        memcpy(arena + at, p, (size_t)len);
        return 0;
    }
    const unsigned char *src = (const unsigned char *)p;
    unsigned char *dst = (unsigned char *)arena;
    unsigned bad = 0;
    int32_t i = 0;
    if ((at & 1) && len > 0) { /* finish the byte the previous piece of this read started */
        const unsigned c = g_code[src[0]];
        bad |= c;
        dst[at >> 1] |= (unsigned char)(c << 4);
        ++at, ++i;
    }
    for (; i + 1 < len; i += 2, at += 2) {
        const unsigned c0 = g_code[src[i]], c1 = g_code[src[i + 1]];
        bad |= c0 | c1;
        dst[at >> 1] = (unsigned char)(c0 | (c1 << 4));
    }
    if (i < len) {
        const unsigned c = g_code[src[i]];
        bad |= c;
        dst[at >> 1] = (unsigned char)c;
    }
    return bad;
}

typedef struct {
    PyObject **keep;      /* per locus: motif, est_cn, tr_seqs, flank_left_seqs, flank_right_seqs (sequences: fast) */
    const int64_t *rbp;   /* read_begin */
    Py_ssize_t l0, l1;    /* my loci */
    int nibble, phase;    /* phase 0: measure, 1: place + copy */
    int32_t *ln;          /* lens[3 * reads] */
    uint64_t *so;         /* seq_off[reads] */
    char *arena;
    uint64_t base, total; /* symbols: my first offset (in), my total (out of phase 0) */
    int err;              /* seq_view_nogil code of the first bad sequence, 0 = none */
    int bad;              /* a byte without a nibble code was seen */
} WalkJob;

static void *walk_worker(void *arg) {
    WalkJob *jb = (WalkJob *)arg;
    uint64_t at = jb->phase ? jb->base : 0;
    unsigned bad = 0;
    for (Py_ssize_t i = jb->l0; i < jb->l1; ++i) {
        PyObject **k = jb->keep + 5 * i;
        const Py_ssize_t n = PySequence_Fast_GET_SIZE(k[2]);
        PyObject **tr_items = PySequence_Fast_ITEMS(k[2]);
        PyObject **fl_items = PySequence_Fast_ITEMS(k[3]), **fr_items = PySequence_Fast_ITEMS(k[4]);
        Py_ssize_t r_glob = (Py_ssize_t)jb->rbp[i];
        for (Py_ssize_t r = 0; r < n; ++r, ++r_glob) {
            /* every str is its own heap object: the walk is bound by cache misses on the object headers, so the
             * headers of a few reads ahead are prefetched while this one is handled */
            if (r + 6 < n) {
                __builtin_prefetch(fl_items[r + 6]);
                __builtin_prefetch(tr_items[r + 6]);
                __builtin_prefetch(fr_items[r + 6]);
            }
            PyObject *parts[3] = {fl_items[r], tr_items[r], fr_items[r]};
            if (jb->phase) jb->so[r_glob] = at;
            for (int q = 0; q < 3; ++q) {
                const char *p = NULL;
                Py_ssize_t len = 0;
                const int e = seq_view_nogil(parts[q], &p, &len);
                if (e) {
                    if (!jb->err) jb->err = e;
                    len = 0;
                }
                if (jb->phase)
                    bad |= put_piece(jb->arena, at, p, (int32_t)len, jb->nibble);
                else
                    jb->ln[3 * r_glob + q] = (int32_t)len;
                at += (uint64_t)len;
            }
            if (jb->nibble) at += at & 1; /* next read starts on a byte boundary */
        }
    }
    if (!jb->phase) jb->total = at;
    jb->bad = (bad & 16u) != 0;
    return NULL;
}

static void run_jobs(WalkJob *jobs, int nt) {
    pthread_t th[64];
    if (nt == 1) {
        walk_worker(&jobs[0]);
        return;
    }
    int started = 0;
    for (int t = 1; t < nt; ++t) { /* job 0 runs here */
        if (pthread_create(&th[t], NULL, walk_worker, &jobs[t]) != 0) break;
        ++started;
    }
    walk_worker(&jobs[0]);
    for (int t = 1 + started; t < nt; ++t) walk_worker(&jobs[t]); /* could not spawn: do it here */
    for (int t = 1; t <= started; ++t) pthread_join(th[t], NULL);
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
    if (!keep) {
        Py_DECREF(seq);
        return PyErr_NoMemory();
    }
    /* pass 0a: attributes and read counts */
    Py_ssize_t n_reads = 0;
    for (Py_ssize_t i = 0; i < n_loci; ++i) {
        PyObject *lr = PySequence_Fast_GET_ITEM(seq, i);
        PyObject **k = keep + 5 * i;
        if (!(k[0] = PyObject_GetAttr(lr, g_names[0]))) goto fail;
        if (!(k[1] = fast_attr(lr, g_names[1], "est_cn must be a sequence"))) goto fail;
        if (!(k[2] = fast_attr(lr, g_names[2], "tr_seqs must be a sequence"))) goto fail;
        if (!(k[3] = fast_attr(lr, g_names[3], "flank_left_seqs must be a sequence"))) goto fail;
        if (!(k[4] = fast_attr(lr, g_names[4], "flank_right_seqs must be a sequence"))) goto fail;
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
    if (!seq_off || !lens || !est || !rb || !moff || !mlen) {
        if (!PyErr_Occurred()) PyErr_NoMemory();
        goto fail;
    }
    {
        uint64_t *so = (uint64_t *)PyByteArray_AS_STRING(seq_off);
        int32_t *ln = (int32_t *)PyByteArray_AS_STRING(lens);
        int32_t *ec = (int32_t *)PyByteArray_AS_STRING(est);
        int64_t *rbp = (int64_t *)PyByteArray_AS_STRING(rb);
        uint64_t *mo = (uint64_t *)PyByteArray_AS_STRING(moff);
        int32_t *ml = (int32_t *)PyByteArray_AS_STRING(mlen);
        /* pass 0b: read_begin and the start estimates */
        Py_ssize_t r_glob = 0;
        rbp[0] = 0;
        for (Py_ssize_t i = 0; i < n_loci; ++i) {
            PyObject **k = keep + 5 * i;
            const Py_ssize_t n = PySequence_Fast_GET_SIZE(k[1]);
            PyObject **e_items = PySequence_Fast_ITEMS(k[1]);
            for (Py_ssize_t r = 0; r < n; ++r, ++r_glob) {
                const long e = PyLong_AsLong(e_items[r]);
                if (e == -1 && PyErr_Occurred()) goto fail;
                if (e < INT32_MIN || e > INT32_MAX) {
                    PyErr_SetString(PyExc_OverflowError, "pack_loci: est_cn out of int32 range");
                    goto fail;
                }
                ec[r_glob] = (int32_t)e;
            }
            rbp[i + 1] = (int64_t)r_glob;
        }
        /* motifs: measured here, placed after all reads */
        for (Py_ssize_t i = 0; i < n_loci; ++i) {
            const char *p;
            Py_ssize_t len;
            const int e = seq_view_nogil(keep[5 * i], &p, &len);
            if (e) {
                set_view_error(e, "motif");
                goto fail;
            }
            ml[i] = (int32_t)len;
        }
        /* pass 1: ranges of loci with about equal numbers of reads */
        long ncpu = sysconf(_SC_NPROCESSORS_ONLN);
        int nt = threads > 0 ? threads : (int)(ncpu < 1 ? 1 : (ncpu > 16 ? 16 : ncpu));
        if ((Py_ssize_t)nt > n_reads / 4096 + 1) nt = (int)(n_reads / 4096 + 1);
        if (nt > 64) nt = 64;
        WalkJob jobs[64];
        {
            Py_ssize_t i = 0;
            for (int t = 0; t < nt; ++t) {
                const int64_t want = (int64_t)n_reads * (t + 1) / nt;
                jobs[t].l0 = i;
                while (i < n_loci && (t == nt - 1 || rbp[i + 1] <= want)) ++i;
                jobs[t].l1 = i;
                jobs[t].keep = keep, jobs[t].rbp = rbp, jobs[t].nibble = nibble, jobs[t].phase = 0;
                jobs[t].ln = ln, jobs[t].so = so, jobs[t].arena = NULL;
                jobs[t].base = jobs[t].total = 0, jobs[t].err = 0, jobs[t].bad = 0;
            }
        }
        Py_BEGIN_ALLOW_THREADS
        run_jobs(jobs, nt);
        Py_END_ALLOW_THREADS
        uint64_t at = 0; /* symbols */
        for (int t = 0; t < nt; ++t) {
            if (jobs[t].err) {
                set_view_error(jobs[t].err, "sequences");
                goto fail;
            }
            jobs[t].base = at;
            at += jobs[t].total; /* (a thread's total ends on a byte boundary in the nibble layout) */
        }
        for (Py_ssize_t i = 0; i < n_loci; ++i) { /* motifs after all reads */
            mo[i] = at;
            at += (uint64_t)ml[i];
            if (nibble) at += at & 1;
        }
        arena = PyByteArray_FromStringAndSize(NULL, (Py_ssize_t)(nibble ? (at + 1) / 2 : at));
        if (!arena) goto fail;
        char *a = PyByteArray_AS_STRING(arena);
        for (int t = 0; t < nt; ++t) jobs[t].phase = 1, jobs[t].arena = a;
        unsigned bad = 0;
        Py_BEGIN_ALLOW_THREADS
        run_jobs(jobs, nt);
        for (Py_ssize_t i = 0; i < n_loci; ++i) {
            const char *p = NULL;
            Py_ssize_t len = 0;
            seq_view_nogil(keep[5 * i], &p, &len);
            bad |= put_piece(a, mo[i], p, (int32_t)len, nibble);
        }
        Py_END_ALLOW_THREADS
        for (int t = 0; t < nt; ++t) bad |= jobs[t].bad ? 16u : 0u;
        if (bad & 16u) {
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
    static const char *names[5] = {"motif", "est_cn", "tr_seqs", "flank_left_seqs", "flank_right_seqs"};
    init_codes();
    for (int i = 0; i < 5; ++i)
        if (!(g_names[i] = PyUnicode_InternFromString(names[i]))) return NULL;
    return PyModule_Create(&moddef);
}
