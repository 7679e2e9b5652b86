/*
 * hostmap.c -- the host half of `Indexer.search_knn` after the GPU search: row ids -> external
 * ids, as a CPython extension (`_b2ip_hostmap`).
 *
 * Replaces the reference's nq*k-iteration Python comprehension
 *     db_ids = [[str(self.index_id_to_db_id[i]) for i in query_top_idxs] for query_top_idxs in indexes]
 *     result.extend([(db_ids[i], scores[i]) for i in range(len(db_ids))])
 * (reference src/index.py:44-45) with C that produces the SAME objects: a list of
 * (list[str] of length k, scores_row) tuples.  `str(x)` of an exact `str` is x itself, which is
 * what the reference's ids are (TSV ids, src/data.py:668-672), so the common case is a pointer
 * copy + incref; any other id type goes through PyObject_Str exactly like the reference.
 * Negative rows index from the end like a Python list does (the reference's `[-1]` quirk for
 * faiss's -1 padding when the index holds fewer than k rows); out-of-range rows raise IndexError.
 *
 * The work is a random gather over a 21M-entry pointer array and 21M object headers: two cache
 * misses per id, ~50 ns per id on one core even with software prefetch -- 0.5 s for 100k x 100
 * results, more than an 8-GPU search takes.  So the gather runs on a small pool of helper threads:
 *   - the calling thread (which holds the GIL for the whole call, so no Python code runs anywhere
 *     else in the process) first creates every result list / tuple;
 *   - then it and the helpers fill disjoint ranges of result rows: item pointer into the list
 *     slot, reference count raised with an ATOMIC add (several threads may hit the same id);
 *     immortal objects are left alone like Py_INCREF does; an id that is not an exact `str` is
 *     left for the calling thread, which calls PyObject_Str on it afterwards (object creation
 *     needs the GIL holder);
 *   - the helpers never allocate, never call into the interpreter and touch nothing but
 *     `ob_refcnt` of the id objects and slots of lists nobody else can see yet.
 * B2IP_MAP_THREADS sets the number of threads (default: min(8, online cores); 1 = serial).
 *
 * The cyclic garbage collector is what made this call slow in practice, not the gather: 2*nq new
 * containers trip the generation thresholds hundreds of times, and every full collection walks the
 * 21M-entry id list (one cache miss per id object).  Measured for 100k x 100 results over 21M ids,
 * 8 threads: 1.5 s with the collector running, 0.13 s without.  So the collector is switched off
 * for the duration of the call (and put back as it was), and the containers built here are taken
 * off its lists before they are returned: a list of `str` and a (list, score row) tuple cannot be
 * part of a reference cycle, which is the same reasoning by which CPython itself untracks tuples and
 * dicts of atomic values -- later collections in the caller's program then do not walk nq*k
 * pointers either.  (Rows that needed PyObject_Str, i.e. ids that are not `str`, stay tracked.)
 * B2IP_MAP_UNTRACK=0 keeps everything tracked.
 */
#define PY_SSIZE_T_CLEAN
#include <Python.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <unistd.h>

#define PF_FAR 32   /* ids ahead for the ob_item[] slot */
#define PF_NEAR 12  /* ids ahead for the object header (its refcount is written) */
#define MAX_THREADS 16
#define PAR_MIN_IDS 32768   /* below this the pool's wake-up costs more than it saves */

static inline Py_ssize_t norm_row(int64_t r, Py_ssize_t n) { return (Py_ssize_t)(r < 0 ? r + n : r); }

typedef struct {
    PyObject** items;        /* ids->ob_item */
    Py_ssize_t n;            /* len(ids) */
    const int64_t* rows;     /* [nq*k] */
    PyObject** row_lists;    /* [nq] result row lists (items NULL on entry) */
    Py_ssize_t nq, k;
    int bad;                 /* a row index was out of range */
    int leftover;            /* some id is not an exact str: slot left NULL for the caller */
} job_t;

/* rows [q0, q1): item pointers + atomic increfs.  No interpreter calls. */
static void fill_range(job_t* jb, Py_ssize_t q0, Py_ssize_t q1) {
    PyObject** items = jb->items;
    const Py_ssize_t n = jb->n, k = jb->k;
    const int64_t* r = jb->rows;
    const Py_ssize_t t_end = q1 * k;
    int bad = 0, leftover = 0;
    for (Py_ssize_t q = q0; q < q1 && !bad; q++) {
        PyObject** dst = ((PyListObject*)jb->row_lists[q])->ob_item;
        const Py_ssize_t base = q * k;
        for (Py_ssize_t j = 0; j < k; j++) {
            const Py_ssize_t t = base + j;
            if (t + PF_FAR < t_end) {
                const Py_ssize_t f = norm_row(r[t + PF_FAR], n);
                if ((size_t)f < (size_t)n) __builtin_prefetch(items + f, 0, 0);
            }
            if (t + PF_NEAR < t_end) {
                const Py_ssize_t f = norm_row(r[t + PF_NEAR], n);
                if ((size_t)f < (size_t)n) __builtin_prefetch(items[f], 1, 0);
            }
            const Py_ssize_t i = norm_row(r[t], n);
            if ((size_t)i >= (size_t)n) { bad = 1; break; }
            PyObject* item = items[i];
            if (PyUnicode_CheckExact(item)) {
                if (!_Py_IsImmortal(item)) __atomic_fetch_add(&item->ob_refcnt, 1, __ATOMIC_RELAXED);
                dst[j] = item;
            } else {
                leftover = 1;            /* dst[j] stays NULL: str(item) by the GIL holder */
            }
        }
    }
    if (bad) __atomic_store_n(&jb->bad, 1, __ATOMIC_RELAXED);
    if (leftover) __atomic_store_n(&jb->leftover, 1, __ATOMIC_RELAXED);
}

/* ---------------------------------------------------------------- helper pool */
static struct {
    pthread_t th[MAX_THREADS];
    int n_helpers;           /* threads besides the caller */
    int started;
    pthread_mutex_t m;
    pthread_cond_t wake, done;
    unsigned long gen;
    int left;
    job_t* job;
    Py_ssize_t per;          /* rows per participant */
} pool = {.m = PTHREAD_MUTEX_INITIALIZER, .wake = PTHREAD_COND_INITIALIZER, .done = PTHREAD_COND_INITIALIZER};

static void* helper_main(void* arg) {
    const int me = (int)(intptr_t)arg;       /* participant index 1..n_helpers (0 = the caller) */
    unsigned long seen = 0;
    pthread_mutex_lock(&pool.m);
    for (;;) {
        while (pool.gen == seen) pthread_cond_wait(&pool.wake, &pool.m);
        seen = pool.gen;
        job_t* jb = pool.job;
        const Py_ssize_t per = pool.per;
        pthread_mutex_unlock(&pool.m);
        Py_ssize_t q0 = me * per, q1 = q0 + per;
        if (q0 > jb->nq) q0 = jb->nq;
        if (q1 > jb->nq) q1 = jb->nq;
        if (q1 > q0) fill_range(jb, q0, q1);
        pthread_mutex_lock(&pool.m);
        if (--pool.left == 0) pthread_cond_signal(&pool.done);
    }
    return NULL;
}

static int pool_threads(void) {
    if (!pool.started) {
        pool.started = 1;
        long want = sysconf(_SC_NPROCESSORS_ONLN);
        if (want > 8) want = 8;
        const char* e = getenv("B2IP_MAP_THREADS");
        if (e) want = atol(e);
        if (want < 1) want = 1;
        if (want > MAX_THREADS) want = MAX_THREADS;
        pool.n_helpers = 0;
        for (long i = 1; i < want; i++) {
            if (pthread_create(&pool.th[pool.n_helpers], NULL, helper_main, (void*)(intptr_t)(pool.n_helpers + 1)) != 0) break;
            pthread_detach(pool.th[pool.n_helpers]);
            pool.n_helpers++;
        }
    }
    return pool.n_helpers + 1;
}

static void run_job(job_t* jb) {
    const int T = (jb->nq * jb->k >= PAR_MIN_IDS) ? pool_threads() : 1;
    if (T <= 1) { fill_range(jb, 0, jb->nq); return; }
    const Py_ssize_t per = (jb->nq + T - 1) / T;
    pthread_mutex_lock(&pool.m);
    pool.job = jb; pool.per = per; pool.left = pool.n_helpers; pool.gen++;
    pthread_cond_broadcast(&pool.wake);
    pthread_mutex_unlock(&pool.m);
    fill_range(jb, 0, per < jb->nq ? per : jb->nq);          /* the caller is participant 0 */
    pthread_mutex_lock(&pool.m);
    while (pool.left != 0) pthread_cond_wait(&pool.done, &pool.m);
    pthread_mutex_unlock(&pool.m);
}

static int untrack_results(void) {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("B2IP_MAP_UNTRACK");
        v = !(e && e[0] == '0');
    }
    return v;
}

/* map_ids(ids: list, rows: int64 C-contiguous buffer [nq*k], nq: int, k: int, scores: sequence | None)
 *   -> [ (list[str], scores[i]) ... ]   (or [list[str] ...] when scores is None) */
static PyObject* map_ids(PyObject* self, PyObject* args) {
    PyObject *ids, *scores;
    Py_buffer rows;
    Py_ssize_t nq, k;
    (void)self;
    if (!PyArg_ParseTuple(args, "O!y*nnO", &PyList_Type, &ids, &rows, &nq, &k, &scores)) return NULL;
    PyObject* out = NULL;
    PyObject* fast_scores = NULL;
    PyObject** row_lists = NULL;
    int gc_was_on = 0;
    if (nq < 0 || k < 0 || rows.len != (Py_ssize_t)(nq * k * (Py_ssize_t)sizeof(int64_t))) {
        PyErr_SetString(PyExc_ValueError, "map_ids: rows buffer is not int64[nq*k]");
        goto fail;
    }
    if (scores != Py_None) {
        fast_scores = PySequence_Fast(scores, "map_ids: scores must be a sequence of nq rows");
        if (!fast_scores) goto fail;
        if (PySequence_Fast_GET_SIZE(fast_scores) != nq) {
            PyErr_SetString(PyExc_ValueError, "map_ids: len(scores) != nq");
            goto fail;
        }
    }
    gc_was_on = PyGC_Disable();
    out = PyList_New(nq);
    if (!out) goto fail;
    row_lists = (PyObject**)malloc((size_t)(nq > 0 ? nq : 1) * sizeof(PyObject*));
    if (!row_lists) { PyErr_NoMemory(); goto fail; }
    /* 1. every container, by the GIL holder (items of the row lists are NULL until step 2) */
    for (Py_ssize_t q = 0; q < nq; q++) {
        PyObject* row = PyList_New(k);
        if (!row) goto fail;
        PyObject* entry = row;
        if (fast_scores) {
            PyObject* s = PySequence_Fast_GET_ITEM(fast_scores, q);
            entry = PyTuple_New(2);
            if (!entry) { Py_DECREF(row); goto fail; }
            Py_INCREF(s);
            PyTuple_SET_ITEM(entry, 0, row);
            PyTuple_SET_ITEM(entry, 1, s);
        }
        PyList_SET_ITEM(out, q, entry);      /* `out` owns the row from here on: `goto fail` frees it */
        row_lists[q] = row;
    }
    /* 2. the gather (this thread + helpers; no interpreter calls inside) */
    job_t jb;
    jb.items = ((PyListObject*)ids)->ob_item;
    jb.n = PyList_GET_SIZE(ids);
    jb.rows = (const int64_t*)rows.buf;
    jb.row_lists = row_lists;
    jb.nq = nq; jb.k = k; jb.bad = 0; jb.leftover = 0;
    run_job(&jb);
    if (jb.bad) {
        PyErr_SetString(PyExc_IndexError, "list index out of range");
        goto fail;                           /* list_dealloc skips the slots that are still NULL */
    }
    /* 3. ids that are not exact `str`: str(id), like the reference */
    if (jb.leftover) {
        const int64_t* r = (const int64_t*)rows.buf;
        for (Py_ssize_t q = 0; q < nq; q++) {
            PyObject** dst = ((PyListObject*)row_lists[q])->ob_item;
            for (Py_ssize_t j = 0; j < k; j++) {
                if (dst[j]) continue;
                const Py_ssize_t n_now = PyList_GET_SIZE(ids);       /* PyObject_Str may run Python code */
                const Py_ssize_t i = norm_row(r[q * k + j], n_now);
                if ((size_t)i >= (size_t)n_now) {
                    PyErr_SetString(PyExc_IndexError, "list index out of range");
                    goto fail;
                }
                PyObject* s = PyObject_Str(((PyListObject*)ids)->ob_item[i]);
                if (!s) goto fail;
                dst[j] = s;
            }
        }
    }
    /* 4. off the collector's lists: list[str] and (list[str], score row) cannot form cycles */
    if (!jb.leftover && untrack_results()) {
        for (Py_ssize_t q = 0; q < nq; q++) {
            PyObject* entry = PyList_GET_ITEM(out, q);
            PyObject_GC_UnTrack(row_lists[q]);
            if (entry != row_lists[q] && !PyObject_GC_IsTracked(PyTuple_GET_ITEM(entry, 1))) PyObject_GC_UnTrack(entry);
        }
    }
    free(row_lists);
    Py_XDECREF(fast_scores);
    PyBuffer_Release(&rows);
    if (gc_was_on) PyGC_Enable();
    return out;
fail:
    if (gc_was_on) PyGC_Enable();
    free(row_lists);
    Py_XDECREF(out);
    Py_XDECREF(fast_scores);
    PyBuffer_Release(&rows);
    return NULL;
}

static PyObject* map_threads(PyObject* self, PyObject* noargs) {
    (void)self; (void)noargs;
    return PyLong_FromLong(pool_threads());
}

static PyMethodDef methods[] = {
    {"map_ids", map_ids, METH_VARARGS,
     "map_ids(ids, rows_int64_buffer, nq, k, scores_or_None) -> [(list[str], scores[i]), ...]"},
    {"map_threads", map_threads, METH_NOARGS, "number of threads map_ids uses for large inputs"},
    {NULL, NULL, 0, NULL}};

static struct PyModuleDef moddef = {PyModuleDef_HEAD_INIT, "_b2ip_hostmap",
                                    "row -> external id mapping of Indexer.search_knn (src/index.py:44-45)",
                                    -1, methods, NULL, NULL, NULL, NULL};

PyMODINIT_FUNC PyInit__b2ip_hostmap(void) { return PyModule_Create(&moddef); }
