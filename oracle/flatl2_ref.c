/*
 * oracle/flatl2_ref.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement of the arithmetic behind `faiss.IndexFlatL2.search()` as the
 * AGPlace reference calls it (reference call sites: test.py:27-32,
 * datasets/datasets_ws_kitti360.py:976-993, datasets/datasets_ws_nuscenes.py:1241-1258,
 * datasets_ws.py:689-706).  The arithmetic itself lives in the third-party wheel
 * `faiss-cpu`, which the reference installs UN-PINNED (README.md:45) and does not
 * vendor; faiss is absent from this image and cannot be installed (no network).
 * This file therefore restates faiss's published IndexFlatL2 algorithm
 * (faiss/utils/distances.cpp: exhaustive_L2sqr_seq / exhaustive_L2sqr_blas,
 * faiss/impl/ResultHandler.h: Top1/Heap/Reservoir handlers, faiss/utils/Heap.h):
 *
 *   nq <  20 : per query, exact sum_i (x_i - y_i)^2 in fp32, 8-lane SIMD order
 *   nq >= 20 : row norms, blocks of 4096 queries x 1024 database rows,
 *              ip = sgemm block, dis = (xn + yn) - 2*ip, clamp at 0
 *   k == 1   : running strict minimum      (lowest index wins exact ties)
 *   k <  100 : max-heap, strict `thr > dis` admission, final ascending sort
 *   k >= 100 : reservoir of capacity 2k, strict admission, shrink, final sort
 *   padding  : (FLT_MAX, -1) when fewer than k database rows exist
 *
 * PARITY PIN: the reference ships no golden vectors / tests for this path and faiss
 * cannot be run here.  The restatement is pinned against the one known-answer vector
 * real faiss publishes (the output of its tutorial/python/1-Flat.py, transcribed into
 * tests/golden/faiss_tutorial_1flat.json: all ids and printed distances reproduced),
 * an independent fp64 brute force (oracle/flatl2_oracle.py) and hand-written
 * known-answer vectors (tests/golden); beyond that single published vector every
 * parity statement is "against the restatement".
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library.  The sgemm of the nq >= 20 path is supplied by the
 * caller (numpy/OpenBLAS in flatl2_oracle.py), exactly as faiss delegates to BLAS;
 * a plain C fallback sgemm is provided for self-contained use.
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORACLE_API __attribute__((visibility("default")))

/* ---- fvec_L2sqr / fvec_norm_L2sqr: 8 accumulator lanes like the AVX2 build ---- */
static float l2sqr_8lane(const float* x, const float* y, int64_t d) {
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    int64_t i = 0;
    for (; i + 8 <= d; i += 8)
        for (int l = 0; l < 8; ++l) {
            float t = x[i + l] - y[i + l];
            acc[l] += t * t;
        }
    float s = ((acc[0] + acc[4]) + (acc[2] + acc[6])) + ((acc[1] + acc[5]) + (acc[3] + acc[7]));
    for (; i < d; ++i) {
        float t = x[i] - y[i];
        s += t * t;
    }
    return s;
}

static float normsqr_8lane(const float* x, int64_t d) {
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    int64_t i = 0;
    for (; i + 8 <= d; i += 8)
        for (int l = 0; l < 8; ++l) acc[l] += x[i + l] * x[i + l];
    float s = ((acc[0] + acc[4]) + (acc[2] + acc[6])) + ((acc[1] + acc[5]) + (acc[3] + acc[7]));
    for (; i < d; ++i) s += x[i] * x[i];
    return s;
}

ORACLE_API void oracle_norms_l2sqr(float* norms, const float* x, int64_t d, int64_t n) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) norms[i] = normsqr_8lane(x + i * d, d);
}

/* ---- max-heap on (dis, id), 1-based like faiss/utils/Heap.h, ties ordered by id ---- */
static inline int gt2(float a, int64_t ia, float b, int64_t ib) {
    return a > b || (a == b && ia > ib);
}

static void heap_replace_top(int64_t k, float* hd, int64_t* hi, float v, int64_t id) {
    hd--; hi--; /* 1-based */
    int64_t i = 1;
    for (;;) {
        int64_t i1 = i << 1, i2 = i1 + 1;
        if (i1 > k) break;
        if (i2 == k + 1 || gt2(hd[i1], hi[i1], hd[i2], hi[i2])) {
            if (gt2(v, id, hd[i1], hi[i1])) break;
            hd[i] = hd[i1]; hi[i] = hi[i1]; i = i1;
        } else {
            if (gt2(v, id, hd[i2], hi[i2])) break;
            hd[i] = hd[i2]; hi[i] = hi[i2]; i = i2;
        }
    }
    hd[i] = v; hi[i] = id;
}

typedef struct { float d; int64_t i; } pair_t;
static int pair_cmp(const void* a, const void* b) {
    const pair_t* p = (const pair_t*)a; const pair_t* q = (const pair_t*)b;
    if (p->d < q->d) return -1;
    if (p->d > q->d) return 1;
    /* id -1 (padding) sorts last among equal FLT_MAX */
    uint64_t pi = (uint64_t)p->i, qi = (uint64_t)q->i;
    return pi < qi ? -1 : (pi > qi ? 1 : 0);
}

/* sort n (dis,id) pairs ascending by (dis,id) and emit the first k, pad (FLT_MAX,-1) */
static void emit_sorted(pair_t* tmp, int64_t n, int64_t k, float* D, int64_t* I) {
    qsort(tmp, (size_t)n, sizeof(pair_t), pair_cmp);
    int64_t m = n < k ? n : k;
    for (int64_t j = 0; j < m; ++j) { D[j] = tmp[j].d; I[j] = tmp[j].i; }
    for (int64_t j = m; j < k; ++j) { D[j] = FLT_MAX; I[j] = -1; }
}

/* ---- result-handler state: one per query, lives across database blocks ---- */
typedef struct {
    int64_t nq, k, cap;   /* cap = k (heap / top1) or 2k (reservoir) */
    int mode;             /* 0 = top1, 1 = heap, 2 = reservoir */
    float* dis;           /* [nq * cap] */
    int64_t* ids;         /* [nq * cap] */
    int64_t* fill;        /* reservoir: entries used   [nq] */
    float* thr;           /* reservoir: admission threshold [nq] */
} handler_t;

ORACLE_API handler_t* oracle_handler_new(int64_t nq, int64_t k) {
    handler_t* h = (handler_t*)calloc(1, sizeof(handler_t));
    h->nq = nq; h->k = k;
    h->mode = (k == 1) ? 0 : (k < 100 ? 1 : 2);
    h->cap = h->mode == 2 ? 2 * k : k;
    h->dis = (float*)malloc(sizeof(float) * (size_t)(nq * h->cap));
    h->ids = (int64_t*)malloc(sizeof(int64_t) * (size_t)(nq * h->cap));
    h->fill = (int64_t*)calloc((size_t)nq, sizeof(int64_t));
    h->thr = (float*)malloc(sizeof(float) * (size_t)nq);
    for (int64_t i = 0; i < nq * h->cap; ++i) { h->dis[i] = FLT_MAX; h->ids[i] = -1; }
    for (int64_t i = 0; i < nq; ++i) h->thr[i] = FLT_MAX;
    return h;
}

ORACLE_API void oracle_handler_free(handler_t* h) {
    if (!h) return;
    free(h->dis); free(h->ids); free(h->fill); free(h->thr); free(h);
}

/* reservoir shrink: keep the k smallest by (dis,id); threshold becomes the k-th distance */
static void reservoir_shrink(handler_t* h, int64_t q) {
    int64_t n = h->fill[q], k = h->k;
    float* d = h->dis + q * h->cap; int64_t* id = h->ids + q * h->cap;
    pair_t* tmp = (pair_t*)malloc(sizeof(pair_t) * (size_t)n);
    for (int64_t j = 0; j < n; ++j) { tmp[j].d = d[j]; tmp[j].i = id[j]; }
    qsort(tmp, (size_t)n, sizeof(pair_t), pair_cmp);
    int64_t m = n < k ? n : k;
    for (int64_t j = 0; j < m; ++j) { d[j] = tmp[j].d; id[j] = tmp[j].i; }
    h->fill[q] = m;
    if (m == k) h->thr[q] = d[k - 1];
    free(tmp);
}

static inline void handler_add(handler_t* h, int64_t q, float dis, int64_t j) {
    float* d = h->dis + q * h->cap; int64_t* id = h->ids + q * h->cap;
    if (h->mode == 0) {
        if (dis < d[0]) { d[0] = dis; id[0] = j; }
    } else if (h->mode == 1) {
        if (d[0] > dis) heap_replace_top(h->k, d, id, dis, j);
    } else {
        if (h->thr[q] > dis) {
            if (h->fill[q] == h->cap) reservoir_shrink(h, q);
            /* the shrink may have lowered the threshold below dis */
            if (h->thr[q] > dis) { d[h->fill[q]] = dis; id[h->fill[q]] = j; h->fill[q]++; }
        }
    }
}

/* add_results for one sgemm block: ip[(i1-i0) x (j1-j0)] row-major holds inner products;
 * converts in place to distances like faiss does, then feeds the handlers. */
ORACLE_API void oracle_add_ip_block(handler_t* h, int64_t i0, int64_t i1, int64_t j0, int64_t j1,
                                    float* ip, const float* x_norms, const float* y_norms) {
    int64_t nb = j1 - j0;
#pragma omp parallel for schedule(static)
    for (int64_t i = i0; i < i1; ++i) {
        float* line = ip + (i - i0) * nb;
        for (int64_t j = j0; j < j1; ++j) {
            float dis = x_norms[i] + y_norms[j] - 2 * line[j - j0];
            if (dis < 0) dis = 0;
            line[j - j0] = dis;
        }
        for (int64_t j = j0; j < j1; ++j) handler_add(h, i, line[j - j0], j);
    }
}

/* end_multiple + heap_reorder: ascending (dis,id), padded */
ORACLE_API void oracle_handler_finish(handler_t* h, float* D, int64_t* I) {
#pragma omp parallel for schedule(static)
    for (int64_t q = 0; q < h->nq; ++q) {
        int64_t n = h->mode == 2 ? h->fill[q] : h->k;
        pair_t* tmp = (pair_t*)malloc(sizeof(pair_t) * (size_t)(n > 0 ? n : 1));
        int64_t m = 0;
        for (int64_t j = 0; j < n; ++j) {
            int64_t id = h->ids[q * h->cap + j];
            if (id < 0) continue; /* neutral heap slots */
            tmp[m].d = h->dis[q * h->cap + j]; tmp[m].i = id; ++m;
        }
        emit_sorted(tmp, m, h->k, D + q * h->k, I + q * h->k);
        free(tmp);
    }
}

/* exhaustive_L2sqr_seq: the nq < 20 branch (also callable for any nq as an exact check) */
ORACLE_API void oracle_knn_l2sqr_seq(const float* x, const float* y, int64_t d, int64_t nx, int64_t ny,
                                     int64_t k, float* D, int64_t* I) {
    handler_t* h = oracle_handler_new(nx, k);
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < nx; ++i)
        for (int64_t j = 0; j < ny; ++j) handler_add(h, i, l2sqr_8lane(x + i * d, y + j * d, d), j);
    oracle_handler_finish(h, D, I);
    oracle_handler_free(h);
}

/* plain C sgemm fallback: ip[i][j] = <x_i, y_j>, used when no BLAS is supplied */
ORACLE_API void oracle_ip_block(const float* x, const float* y, int64_t d, int64_t i0, int64_t i1,
                                int64_t j0, int64_t j1, float* ip) {
    int64_t nb = j1 - j0;
#pragma omp parallel for schedule(static)
    for (int64_t i = i0; i < i1; ++i)
        for (int64_t j = j0; j < j1; ++j) {
            const float* a = x + i * d; const float* b = y + j * d;
            float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            int64_t t = 0;
            for (; t + 8 <= d; t += 8)
                for (int l = 0; l < 8; ++l) acc[l] += a[t + l] * b[t + l];
            float s = ((acc[0] + acc[4]) + (acc[2] + acc[6])) + ((acc[1] + acc[5]) + (acc[3] + acc[7]));
            for (; t < d; ++t) s += a[t] * b[t];
            ip[(i - i0) * nb + (j - j0)] = s;
        }
}

/* exhaustive_L2sqr_blas with the C fallback sgemm (self-contained variant) */
ORACLE_API void oracle_knn_l2sqr_blas_c(const float* x, const float* y, int64_t d, int64_t nx, int64_t ny,
                                        int64_t k, float* D, int64_t* I) {
    const int64_t bs_x = 4096, bs_y = 1024;
    float* xn = (float*)malloc(sizeof(float) * (size_t)(nx > 0 ? nx : 1));
    float* yn = (float*)malloc(sizeof(float) * (size_t)(ny > 0 ? ny : 1));
    float* ip = (float*)malloc(sizeof(float) * (size_t)(bs_x * bs_y));
    oracle_norms_l2sqr(xn, x, d, nx);
    oracle_norms_l2sqr(yn, y, d, ny);
    handler_t* h = oracle_handler_new(nx, k);
    for (int64_t i0 = 0; i0 < nx; i0 += bs_x) {
        int64_t i1 = i0 + bs_x < nx ? i0 + bs_x : nx;
        for (int64_t j0 = 0; j0 < ny; j0 += bs_y) {
            int64_t j1 = j0 + bs_y < ny ? j0 + bs_y : ny;
            oracle_ip_block(x, y, d, i0, i1, j0, j1, ip);
            oracle_add_ip_block(h, i0, i1, j0, j1, ip, xn, yn);
        }
    }
    oracle_handler_finish(h, D, I);
    oracle_handler_free(h);
    free(xn); free(yn); free(ip);
}

ORACLE_API int oracle_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
