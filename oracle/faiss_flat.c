/*
 * oracle/faiss_flat.c — TEST INFRASTRUCTURE ONLY. Not part of the product; nothing under lotus_b200/
 * may import, link or execute this file. Only tests/, __graft_entry__.smoke() and bench.py's CPU
 * baseline legs use it, and only as the checker / the timed CPU baseline.
 *
 * PARITY UNPINNED: the arithmetic of the reference's hot path lives in the third-party wheel faiss-cpu
 * (pinned `faiss-cpu>=1.8.0.post1,<2.0.0`, /root/reference/pyproject.toml:24; locked 1.13.0,
 * /root/reference/uv.lock:573-574). It is not vendored under /root/reference, it is not installed in this
 * image and cannot be installed (no network), and the reference's tests hold no numeric golden vectors
 * for this path. This file therefore RESTATES the published algorithm of faiss 1.13
 * from its public sources (file names below are upstream faiss paths) and is anchored on the reference's
 * call sites:
 *     faiss.index_factory(d,"Flat",metric) / Index.add / Index.search   lotus/vector_store/faiss_vs.py:23-24,63-67,75
 *     faiss.Kmeans(d,k,niter,verbose).train ; kmeans.index.search(x,1)  lotus/utils.py:61-65
 *     `_scores > threshold` over the all-pairs join                     lotus/sem_ops/sem_dedup.py:45-46
 *
 * What is restated
 *   - IndexFlat::search -> knn_inner_product / knn_L2sqr (faiss/utils/distances.cpp) with
 *     HeapBlockResultHandler (faiss/impl/ResultHandler.h) over faiss/utils/Heap.h heaps whose sift
 *     comparisons use CMin/CMax::cmp2 (faiss/utils/ordered_key_value.h): admission is STRICT on the value,
 *     eviction removes the lexicographic (value,id) extreme, heap_reorder emits best first. Database rows
 *     are offered in ascending id (the BLAS path walks blocks j0 ascending). k == 1 uses
 *     Top1BlockResultHandler (first best wins).
 *   - the L2 BLAS path's ||x||^2 + ||y||^2 - 2<x,y> with negative results clamped to 0.
 *   - Clustering::train (faiss/Clustering.cpp): subsample to k*max_points_per_centroid with
 *     rand_perm(seed), initial centroids = first k of rand_perm(seed+1), Lloyd iterations with
 *     compute_centroids (sums in point order, scaled by 1/count) and split_clusters (RandomGenerator(1234),
 *     EPS = 1/1024), and faiss.Kmeans' python wrapper defaults (python/extra_wrappers.py).
 *   - RandomGenerator = std::mt19937 (faiss/utils/random.cpp): rand_int(max) = mt() % max,
 *     rand_float() = mt() / float(mt.max()).
 *
 * Scorers. sgemm's summation order is BLAS-implementation specific, so "the" fp32 score of faiss is only
 * defined up to fp32 round-off. Three scorers are provided:
 *   ORC_SCORER_CANONICAL (0): fp64 accumulation in a fixed order, rounded once to fp32. This is the score
 *       the product reports bit-for-bit (include/lotus_b200.h). Order: element i belongs to accumulator
 *       (i>>2)&31; each accumulator takes its elements in increasing i with fma; the 32 accumulators are
 *       combined with a 16,8,4,2,1 butterfly.
 *   ORC_SCORER_F32_SEQ (1): plain fp32 left-to-right accumulation (faiss's small-batch SIMD path has this
 *       precision class); L2 uses the expanded BLAS form with clamp.
 *   ORC_SCORER_F32_FAST (2): same as 1 but the compiler may vectorise/reassociate — the timed CPU baseline.
 *
 * Build: gcc -O3 -fopenmp -fPIC -shared (oracle/__init__.py build()). -ffp-contract=off so that fma() calls are
 * the only fused operations.
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_IP 0
#define ORC_L2 1
#define ORC_SCORER_CANONICAL 0
#define ORC_SCORER_F32_SEQ 1
#define ORC_SCORER_F32_FAST 2

/* use `n` OpenMP threads from now on (torchrun exports OMP_NUM_THREADS=1 to its workers; the CPU arm wants every core) */
void orc_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* ---------------------------------------------------------------------------------------------------- */
/* scorers                                                                                                */
/* ---------------------------------------------------------------------------------------------------- */

static double butterfly32(double* acc) {
    /* lane l ends with the same total on every lane; we only need lane 0's value. The xor-butterfly
     * computes, at each level, t[l] = t[l] + t[l ^ off] for all l simultaneously. */
    double t[32], u[32];
    memcpy(t, acc, sizeof(t));
    for (int off = 16; off >= 1; off >>= 1) {
        for (int l = 0; l < 32; ++l) u[l] = t[l] + t[l ^ off];
        memcpy(t, u, sizeof(t));
    }
    return t[0];
}

/* canonical inner product, see header comment */
float orc_dot_canonical(const float* a, const float* b, int d) {
    double acc[32];
    for (int l = 0; l < 32; ++l) acc[l] = 0.0;
    for (int i = 0; i < d; ++i) {
        int l = (i >> 2) & 31;
        acc[l] = fma((double)a[i], (double)b[i], acc[l]);
    }
    return (float)butterfly32(acc);
}

/* canonical squared L2 distance */
float orc_l2_canonical(const float* a, const float* b, int d) {
    double acc[32];
    for (int l = 0; l < 32; ++l) acc[l] = 0.0;
    for (int i = 0; i < d; ++i) {
        int l = (i >> 2) & 31;
        double diff = (double)a[i] - (double)b[i];
        acc[l] = fma(diff, diff, acc[l]);
    }
    return (float)butterfly32(acc);
}

/* canonical squared norm as a double (used by the product's certification, exposed for tests) */
double orc_norm2_canonical_f64(const float* a, int d) {
    double acc[32];
    for (int l = 0; l < 32; ++l) acc[l] = 0.0;
    for (int i = 0; i < d; ++i) {
        int l = (i >> 2) & 31;
        acc[l] = fma((double)a[i], (double)a[i], acc[l]);
    }
    return butterfly32(acc);
}

static float dot_f32_seq(const float* a, const float* b, int d) {
    volatile float acc = 0.f; /* volatile: forbid vectorised reassociation */
    for (int i = 0; i < d; ++i) acc = acc + a[i] * b[i];
    return acc;
}

static float dot_f32_fast(const float* a, const float* b, int d) {
    float acc = 0.f;
#pragma omp simd reduction(+ : acc)
    for (int i = 0; i < d; ++i) acc += a[i] * b[i];
    return acc;
}

static inline float score_pair(const float* q, const float* x, int d, int metric, int scorer, float qn, float xn) {
    if (scorer == ORC_SCORER_CANONICAL) return metric == ORC_IP ? orc_dot_canonical(q, x, d) : orc_l2_canonical(q, x, d);
    float ip = scorer == ORC_SCORER_F32_SEQ ? dot_f32_seq(q, x, d) : dot_f32_fast(q, x, d);
    if (metric == ORC_IP) return ip;
    /* faiss/utils/distances.cpp exhaustive_L2sqr_blas: dis = x_norms[i] + y_norms[j] - 2 * ip; if (dis < 0) dis = 0 */
    float dis = qn + xn - 2 * ip;
    if (dis < 0) dis = 0;
    return dis;
}

/* dense score matrix out[nq,n] */
void orc_scores(const float* x, int64_t n, int d, const float* q, int64_t nq, int metric, int scorer, float* out) {
    float* xn = NULL;
    if (metric == ORC_L2 && scorer != ORC_SCORER_CANONICAL) {
        xn = (float*)malloc(sizeof(float) * (size_t)(n > 0 ? n : 1));
        for (int64_t j = 0; j < n; ++j)
            xn[j] = scorer == ORC_SCORER_F32_SEQ ? dot_f32_seq(x + j * d, x + j * d, d) : dot_f32_fast(x + j * d, x + j * d, d);
    }
#pragma omp parallel for schedule(dynamic, 1)
    for (int64_t i = 0; i < nq; ++i) {
        const float* qi = q + i * d;
        float qn = 0.f;
        if (xn) qn = scorer == ORC_SCORER_F32_SEQ ? dot_f32_seq(qi, qi, d) : dot_f32_fast(qi, qi, d);
        for (int64_t j = 0; j < n; ++j) out[i * n + j] = score_pair(qi, x + j * d, d, metric, scorer, qn, xn ? xn[j] : 0.f);
    }
    free(xn);
}

/* ---------------------------------------------------------------------------------------------------- */
/* faiss/utils/Heap.h restated. IS_MAX = 1 -> CMax (max-heap, keeps the k smallest: L2);                  */
/*                               IS_MAX = 0 -> CMin (min-heap, keeps the k largest: IP).                  */
/* ---------------------------------------------------------------------------------------------------- */

static inline int c_cmp(int is_max, float a, float b) { return is_max ? (a > b) : (a < b); }
static inline int c_cmp2(int is_max, float a1, float b1, int64_t a2, int64_t b2) {
    /* ordered_key_value.h: (a1 > b1) || ((a1 == b1) && (a2 > b2)) for CMax, same with < for CMin */
    return is_max ? ((a1 > b1) || ((a1 == b1) && (a2 > b2))) : ((a1 < b1) || ((a1 == b1) && (a2 < b2)));
}
static inline float c_neutral(int is_max) { return is_max ? FLT_MAX : -FLT_MAX; }

static void heap_heapify(int is_max, int k, float* val, int64_t* ids) {
    for (int i = 0; i < k; ++i) {
        val[i] = c_neutral(is_max);
        ids[i] = -1;
    }
}

/* sift the (val,id) that currently sits at 1-based position k (heap_pop) or a new (val,id) (replace_top)
 * down from the root of a heap of size k */
static void heap_sift_from_root(int is_max, int k, float* bh_val, int64_t* bh_ids, float val, int64_t id) {
    bh_val--; /* 1-based */
    bh_ids--;
    size_t i = 1, i1, i2;
    while (1) {
        i1 = i << 1;
        i2 = i1 + 1;
        if (i1 > (size_t)k) break;
        if (i2 == (size_t)k + 1 || c_cmp2(is_max, bh_val[i1], bh_val[i2], bh_ids[i1], bh_ids[i2])) {
            if (c_cmp2(is_max, val, bh_val[i1], id, bh_ids[i1])) break;
            bh_val[i] = bh_val[i1];
            bh_ids[i] = bh_ids[i1];
            i = i1;
        } else {
            if (c_cmp2(is_max, val, bh_val[i2], id, bh_ids[i2])) break;
            bh_val[i] = bh_val[i2];
            bh_ids[i] = bh_ids[i2];
            i = i2;
        }
    }
    bh_val[i] = val;
    bh_ids[i] = id;
}

static void heap_replace_top(int is_max, int k, float* bh_val, int64_t* bh_ids, float val, int64_t id) {
    heap_sift_from_root(is_max, k, bh_val, bh_ids, val, id);
}

static void heap_pop(int is_max, int k, float* bh_val, int64_t* bh_ids) {
    /* Heap.h heap_pop: the last element is re-inserted from the root into a heap that still has k slots */
    heap_sift_from_root(is_max, k, bh_val, bh_ids, bh_val[k - 1], bh_ids[k - 1]);
}

static void heap_reorder(int is_max, int k, float* bh_val, int64_t* bh_ids) {
    int i, ii;
    for (i = 0, ii = 0; i < k; i++) {
        /* top element should be put at the end of the list */
        float val = bh_val[0];
        int64_t id = bh_ids[0];
        /* boundary case: we will over-ride this value if not a true element */
        heap_pop(is_max, k - i, bh_val, bh_ids);
        bh_val[k - ii - 1] = val;
        bh_ids[k - ii - 1] = id;
        if (id != -1) ii++;
    }
    /* Count the number of elements which are effectively returned */
    int nel = ii;
    memmove(bh_val, bh_val + k - ii, (size_t)ii * sizeof(*bh_val));
    memmove(bh_ids, bh_ids + k - ii, (size_t)ii * sizeof(*bh_ids));
    for (; ii < k; ii++) {
        bh_val[ii] = c_neutral(is_max);
        bh_ids[ii] = -1;
    }
    (void)nel;
}

/* IndexFlat::search restated: D[nq,k], I[nq,k] */
int orc_knn(const float* x, int64_t n, int d, const float* q, int64_t nq, int k, int metric, int scorer, float* D,
            int64_t* I) {
    if (k <= 0 || d <= 0) return -1;
    const int is_max = metric == ORC_L2; /* L2 keeps the k smallest in a max-heap */
    float* xn = NULL;
    if (metric == ORC_L2 && scorer != ORC_SCORER_CANONICAL) {
        xn = (float*)malloc(sizeof(float) * (size_t)(n > 0 ? n : 1));
#pragma omp parallel for
        for (int64_t j = 0; j < n; ++j)
            xn[j] = scorer == ORC_SCORER_F32_SEQ ? dot_f32_seq(x + j * d, x + j * d, d) : dot_f32_fast(x + j * d, x + j * d, d);
    }
#pragma omp parallel for schedule(dynamic, 1)
    for (int64_t i = 0; i < nq; ++i) {
        const float* qi = q + i * d;
        float* hv = D + i * k;
        int64_t* hi = I + i * k;
        float qn = 0.f;
        if (xn) qn = scorer == ORC_SCORER_F32_SEQ ? dot_f32_seq(qi, qi, d) : dot_f32_fast(qi, qi, d);
        if (k == 1) {
            /* ResultHandler.h Top1BlockResultHandler: if (C::cmp(min_distance, distance)) replace */
            float best = c_neutral(is_max);
            int64_t bid = -1;
            for (int64_t j = 0; j < n; ++j) {
                float s = score_pair(qi, x + j * d, d, metric, scorer, qn, xn ? xn[j] : 0.f);
                if (c_cmp(is_max, best, s)) {
                    best = s;
                    bid = j;
                }
            }
            hv[0] = best;
            hi[0] = bid;
            continue;
        }
        heap_heapify(is_max, k, hv, hi);
        float thresh = hv[0];
        for (int64_t j = 0; j < n; ++j) {
            float s = score_pair(qi, x + j * d, d, metric, scorer, qn, xn ? xn[j] : 0.f);
            /* HeapBlockResultHandler::add_results: if (C::cmp(thresh, dis)) { heap_replace_top; thresh = top } */
            if (c_cmp(is_max, thresh, s)) {
                heap_replace_top(is_max, k, hv, hi, s, j);
                thresh = hv[0];
            }
        }
        heap_reorder(is_max, k, hv, hi);
    }
    free(xn);
    return 0;
}

/* Blocked variant of orc_knn for the timed CPU baseline: faiss's BLAS path (faiss/utils/distances.cpp
 * exhaustive_inner_product_blas / exhaustive_L2sqr_blas) multiplies a block of queries with a block of
 * database rows so that each database row is read once per query block and the FMA pipes stay busy; this
 * does the same with an ORC_QB-query block per thread and a 4-query x 1-row fp32 FMA micro-kernel.
 * Rows are still offered to each heap in ascending id. Scores are fp32 with FMA contraction (the precision
 * class of sgemm); ties and selection follow the same heap code as orc_knn. */
#define ORC_QB 16
static inline void dot4_fma(const float* q0, const float* q1, const float* q2, const float* q3, const float* x, int d,
                            float* out) {
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma omp simd reduction(+ : a0, a1, a2, a3)
    for (int i = 0; i < d; ++i) {
        const float xv = x[i];
        a0 = __builtin_fmaf(q0[i], xv, a0);
        a1 = __builtin_fmaf(q1[i], xv, a1);
        a2 = __builtin_fmaf(q2[i], xv, a2);
        a3 = __builtin_fmaf(q3[i], xv, a3);
    }
    out[0] = a0; out[1] = a1; out[2] = a2; out[3] = a3;
}

/* 4 queries x 2 database rows: 8 independent accumulators, 6 loads per 8 FMAs */
static inline void dot4x2_fma(const float* q0, const float* q1, const float* q2, const float* q3, const float* xa, const float* xb,
                              int d, float* outa, float* outb) {
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f, b0 = 0.f, b1 = 0.f, b2 = 0.f, b3 = 0.f;
#pragma omp simd reduction(+ : a0, a1, a2, a3, b0, b1, b2, b3)
    for (int i = 0; i < d; ++i) {
        const float xv = xa[i], yv = xb[i];
        const float u0 = q0[i], u1 = q1[i], u2 = q2[i], u3 = q3[i];
        a0 = __builtin_fmaf(u0, xv, a0);
        a1 = __builtin_fmaf(u1, xv, a1);
        a2 = __builtin_fmaf(u2, xv, a2);
        a3 = __builtin_fmaf(u3, xv, a3);
        b0 = __builtin_fmaf(u0, yv, b0);
        b1 = __builtin_fmaf(u1, yv, b1);
        b2 = __builtin_fmaf(u2, yv, b2);
        b3 = __builtin_fmaf(u3, yv, b3);
    }
    outa[0] = a0; outa[1] = a1; outa[2] = a2; outa[3] = a3;
    outb[0] = b0; outb[1] = b1; outb[2] = b2; outb[3] = b3;
}

int orc_knn_blocked(const float* x, int64_t n, int d, const float* q, int64_t nq, int k, int metric, float* D,
                    int64_t* I) {
    if (k <= 0 || d <= 0) return -1;
    const int is_max = metric == ORC_L2;
    float* xn = NULL;
    if (metric == ORC_L2) {
        xn = (float*)malloc(sizeof(float) * (size_t)(n > 0 ? n : 1));
#pragma omp parallel for
        for (int64_t j = 0; j < n; ++j) xn[j] = dot_f32_fast(x + j * d, x + j * d, d);
    }
    const int64_t nblocks = (nq + ORC_QB - 1) / ORC_QB;
#pragma omp parallel for schedule(dynamic, 1)
    for (int64_t b = 0; b < nblocks; ++b) {
        const int64_t q0 = b * ORC_QB;
        const int qb = (int)((nq - q0) < ORC_QB ? (nq - q0) : ORC_QB);
        float qn[ORC_QB], thresh[ORC_QB], sc2[2][ORC_QB];
        const float* qp[ORC_QB];
        for (int t = 0; t < ORC_QB; ++t) qp[t] = q + (q0 + (t < qb ? t : qb - 1)) * d; /* pad the block with its last query */
        for (int t = 0; t < qb; ++t) {
            qn[t] = xn ? dot_f32_fast(qp[t], qp[t], d) : 0.f;
            heap_heapify(is_max, k, D + (q0 + t) * k, I + (q0 + t) * k);
            thresh[t] = D[(q0 + t) * k];
        }
        for (int64_t j2 = 0; j2 < n; j2 += 2) {
            const int nr = (n - j2) < 2 ? 1 : 2; /* rows are scored two at a time, offered to the heaps in ascending id */
            const float* xj = x + j2 * d;
            if (nr == 2)
                for (int t = 0; t < ORC_QB; t += 4)
                    dot4x2_fma(qp[t], qp[t + 1], qp[t + 2], qp[t + 3], xj, xj + d, d, sc2[0] + t, sc2[1] + t);
            else
                for (int t = 0; t < ORC_QB; t += 4) dot4_fma(qp[t], qp[t + 1], qp[t + 2], qp[t + 3], xj, d, sc2[0] + t);
            for (int r = 0; r < nr; ++r) {
            const int64_t j = j2 + r;
            const float* sc = sc2[r];
            for (int t = 0; t < qb; ++t) {
                float s = sc[t];
                if (metric == ORC_L2) {
                    s = qn[t] + xn[j] - 2 * s;
                    if (s < 0) s = 0;
                }
                float* hv = D + (q0 + t) * k;
                int64_t* hi = I + (q0 + t) * k;
                if (k == 1) {
                    if (c_cmp(is_max, hv[0], s)) {
                        hv[0] = s;
                        hi[0] = j;
                    }
                } else if (c_cmp(is_max, thresh[t], s)) {
                    heap_replace_top(is_max, k, hv, hi, s, j);
                    thresh[t] = hv[0];
                }
            }
            }
        }
        if (k > 1)
            for (int t = 0; t < qb; ++t) heap_reorder(is_max, k, D + (q0 + t) * k, I + (q0 + t) * k);
    }
    free(xn);
    return 0;
}

/* The same search, tiled for cache reuse: a thread owns a SUPER-block of `sb` x ORC_QB queries and walks the corpus in tiles of
 * ORC_TILE_ROWS rows (a tile of 256 x 768 fp32 = 786 KB stays in the core's L2), scoring every query block of the super-block
 * against the tile before moving on, so the corpus is streamed from memory once per super-block instead of once per 16 queries
 * (orc_knn_blocked is memory bound on many-core hosts: 3 GB of corpus per 16 queries). Same micro-kernel, same heaps, rows still
 * reach every heap in ascending id. This is what faiss's blocked sgemm + heap pass does at the cache level; it is the timed CPU
 * arm of bench.py. `sb` <= 0 picks the largest super-block that still gives every thread work. */
#define ORC_TILE_ROWS 256
#define ORC_MAX_SB 16
int orc_knn_tiled(const float* x, int64_t n, int d, const float* q, int64_t nq, int k, int metric, int sb, float* D, int64_t* I) {
    if (k <= 0 || d <= 0) return -1;
    const int is_max = metric == ORC_L2;
    float* xn = NULL;
    if (metric == ORC_L2) {
        xn = (float*)malloc(sizeof(float) * (size_t)(n > 0 ? n : 1));
#pragma omp parallel for
        for (int64_t j = 0; j < n; ++j) xn[j] = dot_f32_fast(x + j * d, x + j * d, d);
    }
    const int64_t nblocks = (nq + ORC_QB - 1) / ORC_QB;
    if (sb <= 0) {
        const int64_t per_thread = nblocks / (orc_num_threads() > 0 ? orc_num_threads() : 1);
        sb = (int)(per_thread < 1 ? 1 : (per_thread > ORC_MAX_SB ? ORC_MAX_SB : per_thread));
    }
    if (sb > ORC_MAX_SB) sb = ORC_MAX_SB;
    const int64_t nsuper = (nblocks + sb - 1) / sb;
#pragma omp parallel for schedule(dynamic, 1)
    for (int64_t sbi = 0; sbi < nsuper; ++sbi) {
        const int64_t b0 = sbi * sb;
        const int nb = (int)((nblocks - b0) < sb ? (nblocks - b0) : sb);
        float qn[ORC_MAX_SB][ORC_QB], thresh[ORC_MAX_SB][ORC_QB], sc2[2][ORC_QB];
        const float* qp[ORC_MAX_SB][ORC_QB];
        int qb[ORC_MAX_SB];
        for (int u = 0; u < nb; ++u) {
            const int64_t q0 = (b0 + u) * ORC_QB;
            qb[u] = (int)((nq - q0) < ORC_QB ? (nq - q0) : ORC_QB);
            for (int t = 0; t < ORC_QB; ++t) qp[u][t] = q + (q0 + (t < qb[u] ? t : qb[u] - 1)) * d;
            for (int t = 0; t < qb[u]; ++t) {
                qn[u][t] = xn ? dot_f32_fast(qp[u][t], qp[u][t], d) : 0.f;
                heap_heapify(is_max, k, D + (q0 + t) * k, I + (q0 + t) * k);
                thresh[u][t] = D[(q0 + t) * k];
            }
        }
        for (int64_t t0 = 0; t0 < n; t0 += ORC_TILE_ROWS) {
            const int64_t t1 = (t0 + ORC_TILE_ROWS) < n ? (t0 + ORC_TILE_ROWS) : n;
            for (int u = 0; u < nb; ++u) {
                const int64_t q0 = (b0 + u) * ORC_QB;
                for (int64_t j2 = t0; j2 < t1; j2 += 2) {
                    const int nr = (t1 - j2) < 2 ? 1 : 2;
                    const float* xj = x + j2 * d;
                    if (nr == 2)
                        for (int t = 0; t < ORC_QB; t += 4)
                            dot4x2_fma(qp[u][t], qp[u][t + 1], qp[u][t + 2], qp[u][t + 3], xj, xj + d, d, sc2[0] + t, sc2[1] + t);
                    else
                        for (int t = 0; t < ORC_QB; t += 4) dot4_fma(qp[u][t], qp[u][t + 1], qp[u][t + 2], qp[u][t + 3], xj, d, sc2[0] + t);
                    for (int r = 0; r < nr; ++r) {
                        const int64_t j = j2 + r;
                        const float* sc = sc2[r];
                        for (int t = 0; t < qb[u]; ++t) {
                            float s = sc[t];
                            if (metric == ORC_L2) {
                                s = qn[u][t] + xn[j] - 2 * s;
                                if (s < 0) s = 0;
                            }
                            float* hv = D + (q0 + t) * k;
                            int64_t* hi = I + (q0 + t) * k;
                            if (k == 1) {
                                if (c_cmp(is_max, hv[0], s)) {
                                    hv[0] = s;
                                    hi[0] = j;
                                }
                            } else if (c_cmp(is_max, thresh[u][t], s)) {
                                heap_replace_top(is_max, k, hv, hi, s, j);
                                thresh[u][t] = hv[0];
                            }
                        }
                    }
                }
            }
        }
        if (k > 1)
            for (int u = 0; u < nb; ++u)
                for (int t = 0; t < qb[u]; ++t) heap_reorder(is_max, k, D + ((b0 + u) * ORC_QB + t) * k, I + ((b0 + u) * ORC_QB + t) * k);
    }
    free(xn);
    return 0;
}

/* all pairs i<j with canonical IP score > thr (strict). Returns the number found; writes up to cap. */
int64_t orc_threshold_pairs(const float* x, int64_t n, int d, float thr, int64_t* out_i, int64_t* out_j, int64_t cap) {
    int64_t cnt = 0;
    for (int64_t i = 0; i < n; ++i)
        for (int64_t j = i + 1; j < n; ++j) {
            float s = orc_dot_canonical(x + i * d, x + j * d, d);
            if (s > thr) {
                if (cnt < cap) {
                    out_i[cnt] = i;
                    out_j[cnt] = j;
                }
                cnt++;
            }
        }
    return cnt;
}

/* ---------------------------------------------------------------------------------------------------- */
/* std::mt19937 (32-bit Mersenne Twister, the generator behind faiss::RandomGenerator)                   */
/* ---------------------------------------------------------------------------------------------------- */
typedef struct {
    uint32_t mt[624];
    int idx;
} orc_mt19937;

void orc_mt_seed(orc_mt19937* g, uint32_t seed) {
    g->mt[0] = seed;
    for (int i = 1; i < 624; ++i) g->mt[i] = 1812433253u * (g->mt[i - 1] ^ (g->mt[i - 1] >> 30)) + (uint32_t)i;
    g->idx = 624;
}

uint32_t orc_mt_next(orc_mt19937* g) {
    if (g->idx >= 624) {
        for (int i = 0; i < 624; ++i) {
            uint32_t y = (g->mt[i] & 0x80000000u) | (g->mt[(i + 1) % 624] & 0x7fffffffu);
            g->mt[i] = g->mt[(i + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
        }
        g->idx = 0;
    }
    uint32_t y = g->mt[g->idx++];
    y ^= y >> 11;
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= y >> 18;
    return y;
}

/* first `count` outputs of mt19937(seed) (known-answer tests) */
void orc_mt_fill(uint32_t seed, uint32_t* out, int count) {
    orc_mt19937 g;
    orc_mt_seed(&g, seed);
    for (int i = 0; i < count; ++i) out[i] = orc_mt_next(&g);
}

/* faiss/utils/random.cpp rand_perm: Fisher-Yates with rand_int(n - i) = mt() % (n - i) */
void orc_rand_perm(int32_t* perm, int64_t n, int64_t seed) {
    for (int64_t i = 0; i < n; i++) perm[i] = (int32_t)i;
    orc_mt19937 g;
    orc_mt_seed(&g, (uint32_t)seed);
    for (int64_t i = 0; i + 1 < n; i++) {
        int i2 = (int)(i + (int64_t)(orc_mt_next(&g) % (uint32_t)(n - i)));
        int32_t t = perm[i];
        perm[i] = perm[i2];
        perm[i2] = t;
    }
}

/* ---------------------------------------------------------------------------------------------------- */
/* faiss/Clustering.cpp restated                                                                         */
/* ---------------------------------------------------------------------------------------------------- */

/* compute_centroids: sums in point order, then scale by 1/count (float) */
void orc_compute_centroids(int d, int k, int64_t n, const float* x, const int64_t* assign, float* hassign,
                           float* centroids) {
    memset(hassign, 0, sizeof(float) * (size_t)k);
    memset(centroids, 0, sizeof(float) * (size_t)k * d);
    for (int64_t i = 0; i < n; ++i) {
        int64_t ci = assign[i];
        float* c = centroids + ci * d;
        const float* xi = x + i * d;
        hassign[ci] += 1.0f;
        for (int j = 0; j < d; ++j) c[j] += xi[j];
    }
    for (int ci = 0; ci < k; ++ci) {
        if (hassign[ci] == 0) continue;
        float norm = 1 / hassign[ci];
        float* c = centroids + (size_t)ci * d;
        for (int j = 0; j < d; ++j) c[j] *= norm;
    }
}

/* split_clusters: EPS = 1/1024, RandomGenerator rng(1234) */
int orc_split_clusters(int d, int k, int64_t n, float* hassign, float* centroids) {
    const double EPS = 1 / 1024.;
    int nsplit = 0;
    orc_mt19937 g;
    orc_mt_seed(&g, 1234u);
    for (int ci = 0; ci < k; ci++) {
        if (hassign[ci] == 0) { /* need to redefine a centroid */
            int cj;
            for (cj = 0; 1; cj = (cj + 1) % k) {
                /* probability to pick this cluster for split */
                float p = (hassign[cj] - 1.0) / (float)(n - k);
                float r = orc_mt_next(&g) / (float)4294967295u; /* mt() / float(mt.max()) */
                if (r < p) break; /* found our cluster to be split */
            }
            memcpy(centroids + (size_t)ci * d, centroids + (size_t)cj * d, sizeof(*centroids) * (size_t)d);
            /* small symmetric pertubation */
            for (int j = 0; j < d; j++) {
                if (j % 2 == 0) {
                    centroids[(size_t)ci * d + j] *= 1 + EPS;
                    centroids[(size_t)cj * d + j] *= 1 - EPS;
                } else {
                    centroids[(size_t)ci * d + j] *= 1 - EPS;
                    centroids[(size_t)cj * d + j] *= 1 + EPS;
                }
            }
            /* assume even split of the cluster */
            hassign[ci] = hassign[cj] / 2;
            hassign[cj] -= hassign[ci];
            nsplit++;
        }
    }
    return nsplit;
}

/* faiss.Kmeans(d,k,niter=niter).train(x) + index.search(x,1).
 * x[n,d]; out_assign[n]; out_centroids[k,d]; out_obj[niter] (sum of the reported distances per iteration).
 * max_points_per_centroid = 256 (ClusteringParameters default), seed = 1234, nredo = 1. full_lloyd != 0
 * disables the subsampling. Returns 0, or -1 when n < k (faiss throws). */
int orc_kmeans(const float* x, int64_t n, int d, int k, int niter, int64_t seed, int scorer, int full_lloyd,
               int64_t* out_assign, float* out_centroids, float* out_obj) {
    if (n < k || k <= 0) return -1;
    const int max_points_per_centroid = 256;
    const float* xt = x;
    float* xsub = NULL;
    int64_t nx = n;
    if (!full_lloyd && nx > (int64_t)k * max_points_per_centroid) {
        /* subsample_training_set: perm = rand_perm(nx, seed); keep the first k*max_points */
        int32_t* perm = (int32_t*)malloc(sizeof(int32_t) * (size_t)nx);
        orc_rand_perm(perm, nx, seed);
        nx = (int64_t)k * max_points_per_centroid;
        xsub = (float*)malloc(sizeof(float) * (size_t)nx * d);
        for (int64_t i = 0; i < nx; ++i) memcpy(xsub + i * d, x + (int64_t)perm[i] * d, sizeof(float) * (size_t)d);
        free(perm);
        xt = xsub;
    }
    float* centroids = out_centroids;
    if (nx == k) {
        /* "Number of training points same as number of centroids, just copying" */
        memcpy(centroids, xt, sizeof(float) * (size_t)k * d);
    } else {
        int32_t* perm = (int32_t*)malloc(sizeof(int32_t) * (size_t)nx);
        orc_rand_perm(perm, nx, seed + 1); /* seed + 1 + redo * 15486557L, redo = 0 */
        for (int i = 0; i < k; ++i) memcpy(centroids + (size_t)i * d, xt + (int64_t)perm[i] * d, sizeof(float) * (size_t)d);
        free(perm);
        int64_t* assign = (int64_t*)malloc(sizeof(int64_t) * (size_t)nx);
        float* dis = (float*)malloc(sizeof(float) * (size_t)nx);
        float* hassign = (float*)malloc(sizeof(float) * (size_t)k);
        for (int it = 0; it < niter; ++it) {
            orc_knn(centroids, k, d, xt, nx, 1, ORC_L2, scorer, dis, assign);
            double obj = 0;
            for (int64_t i = 0; i < nx; ++i) obj += dis[i];
            if (out_obj) out_obj[it] = (float)obj;
            orc_compute_centroids(d, k, nx, xt, assign, hassign, centroids);
            orc_split_clusters(d, k, nx, hassign, centroids);
        }
        free(assign);
        free(dis);
        free(hassign);
    }
    free(xsub);
    /* lotus/utils.py:65 kmeans.index.search(vec_set, 1) over ALL points */
    float* dis_all = (float*)malloc(sizeof(float) * (size_t)n);
    orc_knn(centroids, k, d, x, n, 1, ORC_L2, scorer, dis_all, out_assign);
    free(dis_all);
    return 0;
}
