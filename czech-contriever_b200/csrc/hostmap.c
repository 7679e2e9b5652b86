/*
 * hostmap.c -- the host half of `Indexer.search_knn` after the GPU search: row ids -> external
 * ids, as a CPython extension (`_b2ip_hostmap`).
 *
 * Replaces the reference's nq*k-iteration Python comprehension
 *     db_ids = [[str(self.index_id_to_db_id[i]) for i in query_top_idxs] for query_top_idxs in indexes]
 *     result.extend([(db_ids[i], scores[i]) for i in range(len(db_ids))])
 * (reference src/index.py:44-45) with one C loop that produces the SAME objects: a list of
 * (list[str] of length k, scores_row) tuples.  `str(x)` of an exact `str` is x itself, which is
 * what the reference's ids are (TSV ids, src/data.py:668-672), so the common case is a pointer
 * copy + incref; any other id type goes through PyObject_Str exactly like the reference.
 * Negative rows index from the end like a Python list does (the reference's `[-1]` quirk for
 * faiss's -1 padding when the index holds fewer than k rows); out-of-range rows raise IndexError.
 *
 * The loop is a random gather over a 21M-entry pointer array and 21M object headers (two cache
 * misses per id): both are software-prefetched a few ids ahead, which is worth ~5x over numpy's
 * fancy-index + tolist (measured in tools/hostmap_bench.py).
 */
#define PY_SSIZE_T_CLEAN
#include <Python.h>
#include <stdint.h>

#define PF_FAR 24   /* ids ahead for the ob_item[] slot */
#define PF_NEAR 8   /* ids ahead for the object header (its refcount is written) */

static inline Py_ssize_t norm_row(int64_t r, Py_ssize_t n) { return (Py_ssize_t)(r < 0 ? r + n : r); }

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
    const int64_t* r = (const int64_t*)rows.buf;
    const Py_ssize_t n = PyList_GET_SIZE(ids);
    const Py_ssize_t total = nq * k;
    out = PyList_New(nq);
    if (!out) goto fail;
    for (Py_ssize_t q = 0; q < nq; q++) {
        PyObject* row = PyList_New(k);
        if (!row) goto fail;
        /* the row list is owned by `out` (or its tuple) from here on, so `goto fail` frees it */
        PyObject* entry = row;
        if (fast_scores) {
            PyObject* s = PySequence_Fast_GET_ITEM(fast_scores, q);
            entry = PyTuple_New(2);
            if (!entry) { Py_DECREF(row); goto fail; }
            Py_INCREF(s);
            PyTuple_SET_ITEM(entry, 0, row);
            PyTuple_SET_ITEM(entry, 1, s);
        }
        PyList_SET_ITEM(out, q, entry);
        const Py_ssize_t base = q * k;
        for (Py_ssize_t j = 0; j < k; j++) {
            const Py_ssize_t t = base + j;
            /* `ids` cannot change under us: the GIL is held and PyObject_Str of a non-str id is
             * the only call-out; ob_item is re-read after it */
            PyObject** items = ((PyListObject*)ids)->ob_item;
            if (t + PF_FAR < total) {
                const Py_ssize_t f = norm_row(r[t + PF_FAR], n);
                if ((size_t)f < (size_t)n) __builtin_prefetch(items + f, 0, 0);
            }
            if (t + PF_NEAR < total) {
                const Py_ssize_t f = norm_row(r[t + PF_NEAR], n);
                if ((size_t)f < (size_t)n) __builtin_prefetch(items[f], 1, 0);
            }
            const Py_ssize_t i = norm_row(r[t], n);
            if ((size_t)i >= (size_t)PyList_GET_SIZE(ids)) {
                PyErr_SetString(PyExc_IndexError, "list index out of range");
                goto fail;
            }
            PyObject* item = items[i];
            PyObject* s;
            if (PyUnicode_CheckExact(item)) {
                Py_INCREF(item);
                s = item;
            } else {
                s = PyObject_Str(item);
                if (!s) goto fail;
            }
            PyList_SET_ITEM(row, j, s);
        }
    }
    Py_XDECREF(fast_scores);
    PyBuffer_Release(&rows);
    return out;
fail:
    Py_XDECREF(out);
    Py_XDECREF(fast_scores);
    PyBuffer_Release(&rows);
    return NULL;
}

static PyMethodDef methods[] = {
    {"map_ids", map_ids, METH_VARARGS,
     "map_ids(ids, rows_int64_buffer, nq, k, scores_or_None) -> [(list[str], scores[i]), ...]"},
    {NULL, NULL, 0, NULL}};

static struct PyModuleDef moddef = {PyModuleDef_HEAD_INIT, "_b2ip_hostmap",
                                    "row -> external id mapping of Indexer.search_knn (src/index.py:44-45)",
                                    -1, methods, NULL, NULL, NULL, NULL};

PyMODINIT_FUNC PyInit__b2ip_hostmap(void) { return PyModule_Create(&moddef); }
