/*
 * flatip_oracle.c -- CPU restatement of the exact inner-product top-k search that the
 * reference reaches through `faiss.IndexFlatIP.search` (reference call site:
 * src/index.py:42, index built at src/index.py:21, rows appended at src/index.py:30).
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing on the product path may link, import or call this
 * file; it is the checker for tests/, __graft_entry__.smoke() and the cpu_baseline /
 * `--impl reference` legs of bench.py.
 *
 * PARITY UNPINNED: the arithmetic lives in the third-party dependency faiss-cpu==1.8.0
 * (reference environment.yml:138), which is neither vendored under /root/reference nor
 * installable offline, and the reference ships no tests / golden vectors for this path.
 * This file restates faiss 1.8.0's published algorithm:
 *   IndexFlat::search (METRIC_INNER_PRODUCT) -> knn_inner_product ->
 *     nx <  20 : exhaustive_inner_product_seq   (per-query SIMD dots, OpenMP over queries)
 *     nx >= 20 : exhaustive_inner_product_blas  (sgemm on 4096-query x 1024-row blocks)
 *   result handler chosen by k: k==1 Top1, k<100 binary min-heap, k>=100 reservoir of
 *   capacity 2k with fuzzy partition; insertion test is STRICT (threshold < score);
 *   rows come back sorted by score descending, unfilled slots are (-1, -FLT_MAX).
 * Where faiss's choice is implementation-defined (which of several exactly-equal
 * scores survives at the k-th boundary, the pivot of the fuzzy partition) this
 * restatement picks one valid realisation; tests compare tie-aware.
 *
 * It is pinned instead against an fp64 brute force on the same inputs (tests/test_oracle.py).
 */
#include <float.h>
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* faiss/utils/distances.cpp globals (v1.8.0 defaults) */
enum {
    BLAS_THRESHOLD = 20,     /* distance_compute_blas_threshold      */
    BLAS_QUERY_BS = 4096,    /* distance_compute_blas_query_bs       */
    BLAS_DATABASE_BS = 1024, /* distance_compute_blas_database_bs    */
    MIN_K_RESERVOIR = 100    /* distance_compute_min_k_reservoir     */
};

/* Fortran sgemm, either LP64 (int) or ILP64 (int64) integer arguments. */
typedef void (*sgemm_lp64_t)(const char*, const char*, const int*, const int*, const int*,
                             const float*, const float*, const int*, const float*, const int*,
                             const float*, float*, const int*);
typedef void (*sgemm_ilp64_t)(const char*, const char*, const int64_t*, const int64_t*,
                              const int64_t*, const float*, const float*, const int64_t*,
                              const float*, const int64_t*, const float*, float*,
                              const int64_t*);

/* ---------------------------------------------------------------- ordering ------- */
/* CMin<float,int64>: keeps the LARGEST values; cmp(a,b) = a < b; neutral = -FLT_MAX.   */
static inline int lt2(float a, float b, int64_t ia, int64_t ib) {
    return (a < b) || (a == b && ia < ib);
}

/* ---------------------------------------------------------------- min-heap ------- */
/* 1-based binary heap, root = smallest (value, id).  Mirrors faiss utils/Heap.h usage
 * by HeapBlockResultHandler: heapify with neutral, replace-top when root < score.      */
static void heap_init(size_t k, float* v, int64_t* id) {
    for (size_t i = 0; i < k; i++) { v[i] = -FLT_MAX; id[i] = -1; }
}

static void heap_sift_from_top(size_t k, float* v0, int64_t* id0, float val, int64_t idx) {
    float* v = v0 - 1; int64_t* id = id0 - 1;
    size_t i = 1;
    for (;;) {
        size_t l = 2 * i, r = l + 1;
        if (l > k) break;
        size_t c = (r > k || lt2(v[l], v[r], id[l], id[r])) ? l : r; /* smaller child */
        if (lt2(val, v[c], idx, id[c])) break;
        v[i] = v[c]; id[i] = id[c]; i = c;
    }
    v[i] = val; id[i] = idx;
}

static void heap_push_back(size_t n_after, float* v0, int64_t* id0, float val, int64_t idx) {
    float* v = v0 - 1; int64_t* id = id0 - 1;
    size_t i = n_after;
    while (i > 1) {
        size_t p = i >> 1;
        if (!lt2(val, v[p], idx, id[p])) break;
        v[i] = v[p]; id[i] = id[p]; i = p;
    }
    v[i] = val; id[i] = idx;
}

/* Pop the root repeatedly and write from the back: output is score-descending, entries
 * whose id is -1 (never filled) end up at the tail -- faiss heap_reorder behaviour.    */
static void heap_sort_desc(size_t k, float* v, int64_t* id) {
    size_t filled = 0;
    for (size_t i = 0; i < k; i++) {
        float val = v[0]; int64_t idx = id[0];
        size_t n = k - i; /* current heap size */
        /* pop: move last to root, sift down in a heap of n-1 */
        float lv = v[n - 1]; int64_t li = id[n - 1];
        if (n > 1) heap_sift_from_top(n - 1, v, id, lv, li);
        v[k - filled - 1] = val; id[k - filled - 1] = idx;
        if (idx != -1) filled++;
    }
    /* compact: the `filled` valid entries sit in the last `filled` slots, descending */
    memmove(v, v + (k - filled), filled * sizeof(float));
    memmove(id, id + (k - filled), filled * sizeof(int64_t));
    for (size_t i = filled; i < k; i++) { v[i] = -FLT_MAX; id[i] = -1; }
}

/* ---------------------------------------------------------------- reservoir ------ */
typedef struct {
    float* vals; int64_t* ids;
    size_t n;        /* requested k */
    size_t capacity; /* 2k */
    size_t i;        /* fill */
    float threshold; /* strict lower bound for admission */
} reservoir_t;

static int cmp_float_desc(const void* a, const void* b) {
    float x = *(const float*)a, y = *(const float*)b;
    return (x < y) - (x > y);
}

/* Keep q in [q_min, q_max] best entries; returns the partition value.  faiss's
 * partition_fuzzy picks its pivot by median-of-3 sampling; any q in the window is a
 * valid outcome, we take the midpoint.  Entries strictly above the pivot all survive,
 * entries equal to it survive in array order until q is reached.                       */
static float reservoir_partition(reservoir_t* r, size_t q_min, size_t q_max) {
    size_t q = (q_min + q_max) / 2;
    size_t n = r->i;
    float* tmp = (float*)malloc(n * sizeof(float));
    memcpy(tmp, r->vals, n * sizeof(float));
    qsort(tmp, n, sizeof(float), cmp_float_desc);
    float thresh = tmp[q - 1];
    free(tmp);
    size_t n_gt = 0;
    for (size_t j = 0; j < n; j++) n_gt += (thresh < r->vals[j]);
    size_t n_eq = q - n_gt, w = 0;
    for (size_t j = 0; j < n; j++) {
        if (thresh < r->vals[j]) { r->vals[w] = r->vals[j]; r->ids[w] = r->ids[j]; w++; }
        else if (n_eq > 0 && r->vals[j] == thresh) {
            r->vals[w] = r->vals[j]; r->ids[w] = r->ids[j]; w++; n_eq--;
        }
    }
    r->i = w;
    return thresh;
}

static inline void reservoir_add(reservoir_t* r, float val, int64_t id) {
    if (r->threshold < val) {
        if (r->i == r->capacity)
            r->threshold = reservoir_partition(r, r->n, (r->capacity + r->n) / 2);
        r->vals[r->i] = val; r->ids[r->i] = id; r->i++;
    }
}

static void reservoir_to_result(reservoir_t* r, float* hv, int64_t* hi) {
    size_t n = r->n, m = r->i < n ? r->i : n;
    for (size_t j = 0; j < m; j++) heap_push_back(j + 1, hv, hi, r->vals[j], r->ids[j]);
    if (r->i < n) {
        for (size_t j = m; j < n; j++) { hv[j] = -FLT_MAX; hi[j] = -1; }
        /* order the filled prefix, leave the neutral tail */
        heap_sort_desc(m, hv, hi);
    } else {
        for (size_t j = n; j < r->i; j++)
            if (hv[0] < r->vals[j]) heap_sift_from_top(n, hv, hi, r->vals[j], r->ids[j]);
        heap_sort_desc(n, hv, hi);
    }
}

/* ---------------------------------------------------------------- handler -------- */
typedef struct {
    int kind; /* 0 top1, 1 heap, 2 reservoir */
    size_t k;
    float* D; int64_t* I;   /* [nx,k] outputs, also the heap storage */
    reservoir_t* res;       /* per query of the current block (kind 2) */
    float* res_vals; int64_t* res_ids;
} handler_t;

static inline void handler_add(handler_t* h, size_t q, size_t q_in_block, float s, int64_t j) {
    float* hv = h->D + q * h->k; int64_t* hi = h->I + q * h->k;
    if (h->kind == 0) { if (hv[0] < s) { hv[0] = s; hi[0] = j; } }
    else if (h->kind == 1) { if (hv[0] < s) heap_sift_from_top(h->k, hv, hi, s, j); }
    else reservoir_add(&h->res[q_in_block], s, j);
}

static void handler_begin(handler_t* h, size_t i0, size_t i1) {
    for (size_t q = i0; q < i1; q++) {
        if (h->kind == 2) {
            reservoir_t* r = &h->res[q - i0];
            r->n = h->k; r->capacity = 2 * h->k; r->i = 0; r->threshold = -FLT_MAX;
            r->vals = h->res_vals + (q - i0) * r->capacity;
            r->ids = h->res_ids + (q - i0) * r->capacity;
        } else heap_init(h->k, h->D + q * h->k, h->I + q * h->k);
    }
}

static void handler_end(handler_t* h, size_t i0, size_t i1) {
#pragma omp parallel for schedule(static)
    for (int64_t q = (int64_t)i0; q < (int64_t)i1; q++) {
        float* hv = h->D + q * h->k; int64_t* hi = h->I + q * h->k;
        if (h->kind == 2) reservoir_to_result(&h->res[q - i0], hv, hi);
        else if (h->kind == 1) heap_sort_desc(h->k, hv, hi);
    }
}

/* ---------------------------------------------------------------- arithmetic ----- */
static inline float dot_f32(const float* a, const float* b, size_t d) {
    /* fvec_inner_product: plain fp32 accumulation (the compiler vectorises it) */
    float s = 0.f;
    for (size_t i = 0; i < d; i++) s += a[i] * b[i];
    return s;
}

/* ip[nxi, nyi] = x[nxi,d] . y[nyi,d]^T without BLAS (portable path, same blocking). */
static void sgemm_internal(const float* x, const float* y, size_t d, size_t nxi, size_t nyi,
                           float* ip) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < (int64_t)nxi; i++)
        for (size_t j = 0; j < nyi; j++) ip[i * nyi + j] = dot_f32(x + i * d, y + j * d, d);
}

/*
 * Exact IP top-k of nx queries against ny database rows (both row-major fp32, dim d).
 *   D[nx,k] scores (descending), I[nx,k] 0-based row ids; tail (-FLT_MAX, -1) if ny < k.
 *   sgemm_ptr: optional Fortran sgemm_; blas_ints: 0 = no BLAS (internal loops),
 *   32 = LP64 symbol, 64 = ILP64 symbol.
 * Returns 0, or -1 on bad arguments / allocation failure.
 */
int oracle_flatip_search(const float* x, const float* y, int64_t d, int64_t nx, int64_t ny,
                         int64_t k, float* D, int64_t* I, void* sgemm_ptr, int blas_ints) {
    if (d <= 0 || nx < 0 || ny < 0 || k <= 0 || !D || !I) return -1;
    if (nx == 0) return 0;
    handler_t h;
    memset(&h, 0, sizeof(h));
    h.k = (size_t)k; h.D = D; h.I = I;
    h.kind = (k == 1) ? 0 : (k < MIN_K_RESERVOIR ? 1 : 2);
    size_t bs_x = (size_t)(nx < BLAS_QUERY_BS ? nx : BLAS_QUERY_BS);
    if (h.kind == 2) {
        h.res = (reservoir_t*)malloc(bs_x * sizeof(reservoir_t));
        h.res_vals = (float*)malloc(bs_x * 2 * h.k * sizeof(float));
        h.res_ids = (int64_t*)malloc(bs_x * 2 * h.k * sizeof(int64_t));
        if (!h.res || !h.res_vals || !h.res_ids) return -1;
    }

    if (nx < BLAS_THRESHOLD) {
        /* exhaustive_inner_product_seq: threads over queries, rows scanned in order */
        handler_begin(&h, 0, (size_t)nx);
#pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < nx; i++)
            for (int64_t j = 0; j < ny; j++)
                handler_add(&h, (size_t)i, (size_t)i, dot_f32(x + i * d, y + j * d, (size_t)d), j);
        handler_end(&h, 0, (size_t)nx);
    } else {
        /* exhaustive_inner_product_blas */
        float* ip = (float*)malloc((size_t)BLAS_QUERY_BS * BLAS_DATABASE_BS * sizeof(float));
        if (!ip) return -1;
        for (int64_t i0 = 0; i0 < nx; i0 += BLAS_QUERY_BS) {
            int64_t i1 = i0 + BLAS_QUERY_BS < nx ? i0 + BLAS_QUERY_BS : nx;
            handler_begin(&h, (size_t)i0, (size_t)i1);
            for (int64_t j0 = 0; j0 < ny; j0 += BLAS_DATABASE_BS) {
                int64_t j1 = j0 + BLAS_DATABASE_BS < ny ? j0 + BLAS_DATABASE_BS : ny;
                int64_t nyi = j1 - j0, nxi = i1 - i0;
                const float one = 1.f, zero = 0.f;
                if (sgemm_ptr && blas_ints == 64) {
                    ((sgemm_ilp64_t)sgemm_ptr)("T", "N", &nyi, &nxi, &d, &one, y + j0 * d, &d,
                                               x + i0 * d, &d, &zero, ip, &nyi);
                } else if (sgemm_ptr && blas_ints == 32) {
                    int a = (int)nyi, b = (int)nxi, c = (int)d;
                    ((sgemm_lp64_t)sgemm_ptr)("T", "N", &a, &b, &c, &one, y + j0 * d, &c,
                                              x + i0 * d, &c, &zero, ip, &a);
                } else {
                    sgemm_internal(x + i0 * d, y + j0 * d, (size_t)d, (size_t)nxi, (size_t)nyi, ip);
                }
                /* add_results(j0, j1, ip_block): threads over the block's queries */
#pragma omp parallel for schedule(static)
                for (int64_t i = i0; i < i1; i++) {
                    const float* line = ip + (i - i0) * nyi;
                    for (int64_t j = 0; j < nyi; j++)
                        handler_add(&h, (size_t)i, (size_t)(i - i0), line[j], j0 + j);
                }
            }
            handler_end(&h, (size_t)i0, (size_t)i1);
        }
        free(ip);
    }
    free(h.res); free(h.res_vals); free(h.res_ids);
    return 0;
}

int oracle_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
