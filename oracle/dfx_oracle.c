/*
 * dfx_oracle.c -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * A plain-C restatement of the arithmetic that the reference
 * (facebookresearch/distributed-faiss) delegates to the third-party `faiss`
 * module on its search path.  faiss (setup.py:32 `faiss-cpu>=1.7.2`, unpinned,
 * un-vendored) is absent from /root/reference, from this container and from
 * the GPU box, so this file restates faiss's *published* algorithms
 * (IndexFlat, IndexIVFFlat, IndexIVFPQ with by_residual + precomputed tables,
 * IndexIVFScalarQuantizer QT_fp16, HeapArray<CMax<float,int64>>, Clustering,
 * ProductQuantizer) and is anchored on the reference's own call sites:
 *
 *   distributed_faiss/index.py:25-48,63-68,94   which faiss object each builder makes
 *   distributed_faiss/index.py:241-270           Index.search  -> faiss_index.search
 *   distributed_faiss/client.py:29-54            ResultHeap    -> float_maxheap_array_t
 *   distributed_faiss/client.py:265-310          _aggregate_results (merge across shards)
 *   tests/test_integration.py:181-203            golden vectors of the merge
 *
 * PARITY STATUS: the merge is pinned by the reference's golden vectors
 * (tests/golden/merge_golden.json).  flat / IVF-Flat / IVF-PQ / IVF-SQ search
 * arithmetic is "parity unpinned" against real faiss (no faiss anywhere, and no
 * reference test searches an IVF index); it is cross-checked against an
 * independent float64 numpy restatement (oracle/ref_numpy.py).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library.
 *
 * CANONICAL ARITHMETIC.  faiss's last-bit results depend on SIMD width, BLAS
 * and version; to make "bit-exact ids" a testable statement the summation
 * orders are pinned here and the CUDA kernels follow the same orders
 * (DESIGN.md "Canonical arithmetic"):
 *   seq-k   : acc = fmaf(a[k], b[k], acc), k ascending (GEMM-shaped paths:
 *             flat, coarse quantizer, PQ LUT / encode)
 *   warp-dot: 32 partial sums, lane j owns k = 128*i + 4*j + t (t=0..3),
 *             then a butterfly (xor 16,8,4,2,1) of plain adds (scan-shaped
 *             paths: IVF-Flat, IVF-SQ, exact ||q-c||^2)
 *   pq-sum  : the M table values reduced by a halving tree (s[i] += s[i+off], off = P/2..1,
 *             P = M rounded up to a power of two); dist = dis0 + (t + S)
 * Every result set is ordered by the TOTAL order (value asc, id asc) where
 * value = distance (L2) or -inner_product (IP).
 *
 * Build: see oracle/Makefile (gcc -O3 -fopenmp -ffp-contract=off -mfma).
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_METRIC_IP 0 /* faiss.METRIC_INNER_PRODUCT (index_cfg.py:47) */
#define ORC_METRIC_L2 1 /* faiss.METRIC_L2            (index_cfg.py:49) */

/* ------------------------------------------------------------------ */
/* scalar building blocks                                              */
/* ------------------------------------------------------------------ */

static inline float ip_seq(const float *a, const float *b, int d) {
    float acc = 0.f;
    for (int k = 0; k < d; k++) acc = fmaf(a[k], b[k], acc);
    return acc;
}

static inline float l2_seq(const float *a, const float *b, int d) {
    float acc = 0.f;
    for (int k = 0; k < d; k++) {
        float df = a[k] - b[k];
        acc = fmaf(df, df, acc);
    }
    return acc;
}

/* IEEE binary16 -> binary32, exact (same value as CUDA __half2float). */
static inline float half_to_float(uint16_t h) {
    uint32_t sign = (uint32_t)(h & 0x8000u) << 16;
    uint32_t exp = (h >> 10) & 0x1f;
    uint32_t man = h & 0x3ffu;
    uint32_t bits;
    if (exp == 0) {
        if (man == 0) {
            bits = sign;
        } else { /* subnormal: normalise */
            int e = -1;
            do {
                e++;
                man <<= 1;
            } while ((man & 0x400u) == 0);
            man &= 0x3ffu;
            bits = sign | ((uint32_t)(127 - 15 - e) << 23) | (man << 13);
        }
    } else if (exp == 31) {
        bits = sign | 0x7f800000u | (man << 13);
    } else {
        bits = sign | ((exp + 127 - 15) << 23) | (man << 13);
    }
    float f;
    memcpy(&f, &bits, 4);
    return f;
}

/* binary32 -> binary16, round-to-nearest-even (F16C vcvtps2ph / __float2half_rn). */
static inline uint16_t float_to_half(float f) {
    uint32_t x;
    memcpy(&x, &f, 4);
    uint32_t sign = (x >> 16) & 0x8000u;
    uint32_t absx = x & 0x7fffffffu;
    if (absx >= 0x7f800000u) { /* inf / nan */
        return (uint16_t)(sign | 0x7c00u | ((absx > 0x7f800000u) ? 0x200u : 0));
    }
    if (absx >= 0x477ff000u) { /* rounds to >= 65520 -> inf */
        return (uint16_t)(sign | 0x7c00u);
    }
    if (absx < 0x38800000u) { /* subnormal half or zero */
        if (absx < 0x33000000u) return (uint16_t)sign; /* < 2^-25 -> 0 */
        uint32_t e = absx >> 23;
        uint32_t man = (absx & 0x7fffffu) | 0x800000u;
        uint32_t shift = 126 - e; /* 14..24 */
        uint32_t hm = man >> shift;
        uint32_t rem = man & ((1u << shift) - 1);
        uint32_t halfway = 1u << (shift - 1);
        if (rem > halfway || (rem == halfway && (hm & 1))) hm++;
        return (uint16_t)(sign | hm);
    }
    uint32_t e = (absx >> 23) - 112;
    uint32_t man = absx & 0x7fffffu;
    uint32_t hm = (e << 10) | (man >> 13);
    uint32_t rem = man & 0x1fffu;
    if (rem > 0x1000u || (rem == 0x1000u && (hm & 1))) hm++;
    return (uint16_t)(sign | hm);
}

/* warp-dot canonical order.  mode 0: sum a*b ; mode 1: sum (a-b)^2.
 * b is float (b16 == NULL) or binary16 (b16 != NULL). d % 4 == 0. */
static float warp_dot(const float *a, const float *b, const uint16_t *b16, int d, int mode) {
    float acc[32];
    for (int j = 0; j < 32; j++) acc[j] = 0.f;
    for (int base0 = 0; base0 < d; base0 += 128) {
        for (int j = 0; j < 32; j++) {
            int base = base0 + 4 * j;
            if (base >= d) break;
            for (int t = 0; t < 4; t++) {
                float bv = b16 ? half_to_float(b16[base + t]) : b[base + t];
                if (mode == 0) {
                    acc[j] = fmaf(a[base + t], bv, acc[j]);
                } else {
                    float df = a[base + t] - bv;
                    acc[j] = fmaf(df, df, acc[j]);
                }
            }
        }
    }
    for (int off = 16; off >= 1; off >>= 1) {
        float nxt[32];
        for (int j = 0; j < 32; j++) nxt[j] = acc[j] + acc[j ^ off];
        for (int j = 0; j < 32; j++) acc[j] = nxt[j];
    }
    return acc[0];
}

float orc_warp_dot(const float *a, const float *b, int d, int mode) {
    return warp_dot(a, b, NULL, d, mode);
}
float orc_warp_dot_h(const float *a, const uint16_t *b16, int d, int mode) {
    return warp_dot(a, NULL, b16, d, mode);
}
void orc_float_to_half(const float *x, uint16_t *out, int64_t n) {
    for (int64_t i = 0; i < n; i++) out[i] = float_to_half(x[i]);
}
void orc_half_to_float(const uint16_t *x, float *out, int64_t n) {
    for (int64_t i = 0; i < n; i++) out[i] = half_to_float(x[i]);
}

/* ------------------------------------------------------------------ */
/* bounded result set under the total order (v asc, id asc)            */
/* ------------------------------------------------------------------ */

typedef struct {
    float v;
    int64_t id;
} cand_t;

static inline int cand_less(float v1, int64_t i1, float v2, int64_t i2) {
    return (v1 < v2) || (v1 == v2 && i1 < i2);
}

/* max-heap on (v,id): h[0] is the WORST of the kept k. */
typedef struct {
    cand_t *h;
    int64_t k, n;
} topk_t;

static void topk_init(topk_t *t, cand_t *buf, int64_t k) {
    t->h = buf;
    t->k = k;
    t->n = 0;
}

static void topk_sift_down(cand_t *h, int64_t n, int64_t i) {
    for (;;) {
        int64_t l = 2 * i + 1, r = l + 1, m = i;
        if (l < n && cand_less(h[m].v, h[m].id, h[l].v, h[l].id)) m = l;
        if (r < n && cand_less(h[m].v, h[m].id, h[r].v, h[r].id)) m = r;
        if (m == i) return;
        cand_t tmp = h[i];
        h[i] = h[m];
        h[m] = tmp;
        i = m;
    }
}

static inline void topk_push(topk_t *t, float v, int64_t id) {
    if (t->n < t->k) {
        int64_t i = t->n++;
        t->h[i].v = v;
        t->h[i].id = id;
        while (i > 0) {
            int64_t p = (i - 1) / 2;
            if (cand_less(t->h[p].v, t->h[p].id, t->h[i].v, t->h[i].id)) {
                cand_t tmp = t->h[p];
                t->h[p] = t->h[i];
                t->h[i] = tmp;
                i = p;
            } else
                break;
        }
    } else if (t->k > 0 && cand_less(v, id, t->h[0].v, t->h[0].id)) {
        t->h[0].v = v;
        t->h[0].id = id;
        topk_sift_down(t->h, t->n, 0);
    }
}

static inline float topk_worst(const topk_t *t) {
    return (t->n < t->k) ? FLT_MAX : t->h[0].v;
}

static int cand_cmp(const void *a, const void *b) {
    const cand_t *x = (const cand_t *)a, *y = (const cand_t *)b;
    if (x->v < y->v) return -1;
    if (x->v > y->v) return 1;
    if (x->id < y->id) return -1;
    if (x->id > y->id) return 1;
    return 0;
}

/* write sorted results; value_sign = -1 turns v=-ip back into ip.
 * missing: id -1, distance +FLT_MAX (L2) / -FLT_MAX (IP)  [faiss convention]. */
static void topk_emit(topk_t *t, float *D, int64_t *I, int metric) {
    qsort(t->h, (size_t)t->n, sizeof(cand_t), cand_cmp);
    for (int64_t j = 0; j < t->k; j++) {
        if (j < t->n) {
            D[j] = (metric == ORC_METRIC_IP) ? -t->h[j].v : t->h[j].v;
            I[j] = t->h[j].id;
        } else {
            D[j] = (metric == ORC_METRIC_IP) ? -FLT_MAX : FLT_MAX;
            I[j] = -1;
        }
    }
}

/* ------------------------------------------------------------------ */
/* IndexFlatIP / IndexFlatL2  (index.py:94 "flat" builder; index.py:25-33 */
/* coarse quantizer)                                                   */
/* ------------------------------------------------------------------ */

/* ranking value of row x for query q under the GEMM-shaped canonical form */
static inline float flat_value(int metric, const float *q, const float *x, float xnorm, int d) {
    float ip = ip_seq(q, x, d);
    if (metric == ORC_METRIC_IP) return -ip;
    return fmaf(-2.f, ip, xnorm); /* ||x||^2 - 2 q.x ; + ||q||^2 added on output */
}

void orc_row_norms(int d, int64_t n, const float *x, float *out) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; i++) out[i] = ip_seq(x + i * d, x + i * d, d);
}

int orc_flat_search(int metric, int d, int64_t n, const float *xb, int64_t nq, const float *xq,
                    int64_t k, float *D, int64_t *I) {
    float *xnorm = NULL;
    if (metric == ORC_METRIC_L2) {
        xnorm = (float *)malloc(sizeof(float) * (size_t)(n > 0 ? n : 1));
        orc_row_norms(d, n, xb, xnorm);
    }
#pragma omp parallel
    {
        cand_t *buf = (cand_t *)malloc(sizeof(cand_t) * (size_t)(k > 0 ? k : 1));
#pragma omp for schedule(dynamic, 1)
        for (int64_t qi = 0; qi < nq; qi++) {
            const float *q = xq + qi * d;
            topk_t t;
            topk_init(&t, buf, k);
            for (int64_t i = 0; i < n; i++) {
                float v = flat_value(metric, q, xb + i * d, xnorm ? xnorm[i] : 0.f, d);
                topk_push(&t, v, i);
            }
            topk_emit(&t, D + qi * k, I + qi * k, metric);
            if (metric == ORC_METRIC_L2) {
                float qn = ip_seq(q, q, d);
                for (int64_t j = 0; j < k; j++) {
                    if (I[qi * k + j] >= 0) {
                        float dd = D[qi * k + j] + qn;
                        D[qi * k + j] = dd < 0.f ? 0.f : dd; /* faiss clamps the BLAS form at 0 */
                    }
                }
            }
        }
        free(buf);
    }
    free(xnorm);
    return 0;
}

/* coarse quantizer: top-nprobe lists per query, best first.  keys[q][p] = list,
 * cval[q][p] = ranking value (L2: ||c||^2 - 2 q.c ; IP: -q.c). (IndexIVF::search
 * -> quantizer->search(nq, x, nprobe)) */
int orc_coarse(int metric, int d, int64_t nlist, const float *cent, int64_t nq, const float *xq,
               int64_t nprobe, int64_t *keys, float *cval) {
    float *cnorm = (float *)malloc(sizeof(float) * (size_t)nlist);
    orc_row_norms(d, nlist, cent, cnorm);
#pragma omp parallel
    {
        cand_t *buf = (cand_t *)malloc(sizeof(cand_t) * (size_t)nprobe);
#pragma omp for schedule(dynamic, 4)
        for (int64_t qi = 0; qi < nq; qi++) {
            const float *q = xq + qi * d;
            topk_t t;
            topk_init(&t, buf, nprobe);
            for (int64_t c = 0; c < nlist; c++)
                topk_push(&t, flat_value(metric, q, cent + c * d, cnorm[c], d), c);
            qsort(t.h, (size_t)t.n, sizeof(cand_t), cand_cmp);
            for (int64_t p = 0; p < nprobe; p++) {
                keys[qi * nprobe + p] = p < t.n ? t.h[p].id : -1;
                if (cval) cval[qi * nprobe + p] = p < t.n ? t.h[p].v : FLT_MAX;
            }
        }
        free(buf);
    }
    free(cnorm);
    return 0;
}

/* nearest centroid per row (quantizer.assign in IndexIVF::add) */
int orc_assign(int metric, int d, int64_t nlist, const float *cent, int64_t n, const float *x,
               int64_t *out) {
    return orc_coarse(metric, d, nlist, cent, n, x, 1, out, NULL);
}

/* ------------------------------------------------------------------ */
/* IndexIVFFlat (index.py:36-40, "ivf_simple")                          */
/* ------------------------------------------------------------------ */
int orc_ivfflat_search(int metric, int d, int64_t nlist, const float *cent,
                       const int64_t *list_off, const float *vecs, const int64_t *ids, int64_t nq,
                       const float *xq, int64_t nprobe, int64_t k, float *D, int64_t *I,
                       int64_t *ndis_out) {
    int64_t *keys = (int64_t *)malloc(sizeof(int64_t) * (size_t)(nq * nprobe));
    orc_coarse(metric, d, nlist, cent, nq, xq, nprobe, keys, NULL);
    int64_t ndis = 0;
#pragma omp parallel reduction(+ : ndis)
    {
        cand_t *buf = (cand_t *)malloc(sizeof(cand_t) * (size_t)(k > 0 ? k : 1));
#pragma omp for schedule(dynamic, 1)
        for (int64_t qi = 0; qi < nq; qi++) {
            const float *q = xq + qi * d;
            topk_t t;
            topk_init(&t, buf, k);
            for (int64_t p = 0; p < nprobe; p++) {
                int64_t l = keys[qi * nprobe + p];
                if (l < 0) continue;
                for (int64_t i = list_off[l]; i < list_off[l + 1]; i++) {
                    float v = (metric == ORC_METRIC_IP) ? -warp_dot(q, vecs + i * d, NULL, d, 0)
                                                        : warp_dot(q, vecs + i * d, NULL, d, 1);
                    topk_push(&t, v, ids[i]);
                }
                ndis += list_off[l + 1] - list_off[l];
            }
            topk_emit(&t, D + qi * k, I + qi * k, metric);
        }
        free(buf);
    }
    if (ndis_out) *ndis_out = ndis;
    free(keys);
    return 0;
}

/* ------------------------------------------------------------------ */
/* IndexIVFPQ (index.py:43-48, "knnlm"): METRIC_L2, by_residual, with    */
/* faiss's precomputed-table decomposition                              */
/*   d(q, c+p) = ||q-c||^2 + sum_m(||p_m||^2 + 2<c_m,p_m>) - 2 sum_m <q_m,p_m> */
/* term 2 is summed per stored vector at add time (tvals).              */
/* ------------------------------------------------------------------ */

/* lut[m][j] = -2 <q_m, P[m][j]> */
void orc_pq_query_lut(int d, int M, int ksub, const float *codebooks, const float *q, float *lut) {
    int dsub = d / M;
    for (int m = 0; m < M; m++)
        for (int j = 0; j < ksub; j++)
            lut[m * ksub + j] =
                -2.f * ip_seq(q + m * dsub, codebooks + ((size_t)m * ksub + j) * dsub, dsub);
}

/* tvals[i] = sum_m ( ||p_m||^2 + 2 <c_m, p_m> ), sequential in m */
void orc_pq_tvals(int d, int M, int ksub, const float *codebooks, const float *cent,
                  const int64_t *list_of, const uint8_t *codes, int64_t n, float *tvals) {
    int dsub = d / M;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; i++) {
        const float *c = cent + list_of[i] * d;
        float t = 0.f;
        for (int m = 0; m < M; m++) {
            const float *p = codebooks + ((size_t)m * ksub + codes[i * M + m]) * dsub;
            float tm = fmaf(2.f, ip_seq(c + m * dsub, p, dsub), ip_seq(p, p, dsub));
            t = t + tm;
        }
        tvals[i] = t;
    }
}

/* pq-sum canonical order: the M table values are padded with +0 to P = next power of two and
 * reduced by a halving tree: for off = P/2, P/4, ..., 1: s[i] = s[i] + s[i+off] (i < off).
 * (This is the order in which a warp butterfly adds them on the GPU.) */
static inline float pq_sum(const float *lut, const uint8_t *code, int M, int ksub) {
    float s[256];
    int P = 1;
    while (P < M) P <<= 1;
    for (int m = 0; m < P; m++) s[m] = (m < M) ? lut[m * ksub + code[m]] : 0.f;
    for (int off = P >> 1; off >= 1; off >>= 1)
        for (int i = 0; i < off; i++) s[i] = s[i] + s[i + off];
    return s[0];
}

int orc_ivfpq_search(int coarse_metric, int d, int64_t nlist, const float *cent, int M, int ksub,
                     const float *codebooks, const int64_t *list_off, const uint8_t *codes,
                     const float *tvals, const int64_t *ids, int64_t nq, const float *xq,
                     int64_t nprobe, int64_t k, float *D, int64_t *I, int64_t *ndis_out) {
    int64_t *keys = (int64_t *)malloc(sizeof(int64_t) * (size_t)(nq * nprobe));
    orc_coarse(coarse_metric, d, nlist, cent, nq, xq, nprobe, keys, NULL);
    int64_t ndis = 0;
#pragma omp parallel reduction(+ : ndis)
    {
        cand_t *buf = (cand_t *)malloc(sizeof(cand_t) * (size_t)(k > 0 ? k : 1));
        float *lut = (float *)malloc(sizeof(float) * (size_t)M * ksub);
#pragma omp for schedule(dynamic, 1)
        for (int64_t qi = 0; qi < nq; qi++) {
            const float *q = xq + qi * d;
            orc_pq_query_lut(d, M, ksub, codebooks, q, lut);
            topk_t t;
            topk_init(&t, buf, k);
            for (int64_t p = 0; p < nprobe; p++) {
                int64_t l = keys[qi * nprobe + p];
                if (l < 0) continue;
                float dis0 = warp_dot(q, cent + l * d, NULL, d, 1);
                for (int64_t i = list_off[l]; i < list_off[l + 1]; i++) {
                    float s = pq_sum(lut, codes + i * M, M, ksub);
                    float v = dis0 + (tvals[i] + s);
                    topk_push(&t, v, ids[i]);
                }
                ndis += list_off[l + 1] - list_off[l];
            }
            topk_emit(&t, D + qi * k, I + qi * k, ORC_METRIC_L2);
        }
        free(lut);
        free(buf);
    }
    if (ndis_out) *ndis_out = ndis;
    free(keys);
    return 0;
}

/* PQ encode of residuals r = x - c[list]: code_m = argmin_j ||r_m - P[m][j]||^2, ties -> smaller j */
void orc_pq_encode(int d, int M, int ksub, const float *codebooks, const float *cent,
                   const int64_t *list_of, const float *x, int64_t n, uint8_t *codes) {
    int dsub = d / M;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; i++) {
        const float *c = cent + list_of[i] * d;
        float r[4096];
        for (int kk = 0; kk < d; kk++) r[kk] = x[i * d + kk] - c[kk];
        for (int m = 0; m < M; m++) {
            float best = FLT_MAX;
            int bj = 0;
            for (int j = 0; j < ksub; j++) {
                float dd = l2_seq(r + m * dsub, codebooks + ((size_t)m * ksub + j) * dsub, dsub);
                if (dd < best) {
                    best = dd;
                    bj = j;
                }
            }
            codes[i * M + m] = (uint8_t)bj;
        }
    }
}

/* ------------------------------------------------------------------ */
/* IndexIVFScalarQuantizer QT_fp16, L2, by_residual (index.py:63-68)     */
/* ------------------------------------------------------------------ */
void orc_sq_encode(int d, const float *cent, const int64_t *list_of, const float *x, int64_t n,
                   uint16_t *codes) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; i++) {
        const float *c = cent + list_of[i] * d;
        for (int kk = 0; kk < d; kk++) codes[i * d + kk] = float_to_half(x[i * d + kk] - c[kk]);
    }
}

int orc_ivfsq_search(int coarse_metric, int d, int64_t nlist, const float *cent,
                     const int64_t *list_off, const uint16_t *codes, const int64_t *ids,
                     int64_t nq, const float *xq, int64_t nprobe, int64_t k, float *D, int64_t *I,
                     int64_t *ndis_out) {
    int64_t *keys = (int64_t *)malloc(sizeof(int64_t) * (size_t)(nq * nprobe));
    orc_coarse(coarse_metric, d, nlist, cent, nq, xq, nprobe, keys, NULL);
    int64_t ndis = 0;
#pragma omp parallel reduction(+ : ndis)
    {
        cand_t *buf = (cand_t *)malloc(sizeof(cand_t) * (size_t)(k > 0 ? k : 1));
        float *rq = (float *)malloc(sizeof(float) * (size_t)d);
#pragma omp for schedule(dynamic, 1)
        for (int64_t qi = 0; qi < nq; qi++) {
            const float *q = xq + qi * d;
            topk_t t;
            topk_init(&t, buf, k);
            for (int64_t p = 0; p < nprobe; p++) {
                int64_t l = keys[qi * nprobe + p];
                if (l < 0) continue;
                for (int kk = 0; kk < d; kk++) rq[kk] = q[kk] - cent[l * d + kk];
                for (int64_t i = list_off[l]; i < list_off[l + 1]; i++) {
                    float v = warp_dot(rq, NULL, codes + i * d, d, 1);
                    topk_push(&t, v, ids[i]);
                }
                ndis += list_off[l + 1] - list_off[l];
            }
            topk_emit(&t, D + qi * k, I + qi * k, ORC_METRIC_L2);
        }
        free(rq);
        free(buf);
    }
    if (ndis_out) *ndis_out = ndis;
    free(keys);
    return 0;
}

/* ------------------------------------------------------------------ */
/* float_maxheap_array_t as used by ResultHeap (client.py:29-54)         */
/*   heapify(): val = FLT_MAX, ids = -1                                 */
/*   addn_with_ids(k, D_s, I_s, k) per shard in sub_indexes order,      */
/*     replace the heap top iff top > v  (strict: earlier insert wins)  */
/*   reorder(): ascending, real entries first, pad (FLT_MAX, -1)        */
/* ids handed in by _aggregate_results are arange positions that grow   */
/* with insertion order (client.py:290), so "earlier insert wins" is     */
/* the total order (v asc, id asc) used here.                            */
/* Dall/Iall: [S][nq][k].                                               */
/* ------------------------------------------------------------------ */
int orc_merge(int64_t S, int64_t nq, int64_t k, const float *Dall, const int64_t *Iall,
              float *outD, int64_t *outI) {
#pragma omp parallel
    {
        cand_t *buf = (cand_t *)malloc(sizeof(cand_t) * (size_t)(k > 0 ? k : 1));
#pragma omp for schedule(static)
        for (int64_t qi = 0; qi < nq; qi++) {
            topk_t t;
            topk_init(&t, buf, k);
            for (int64_t s = 0; s < S; s++) {
                const float *Ds = Dall + (s * nq + qi) * k;
                const int64_t *Is = Iall + (s * nq + qi) * k;
                for (int64_t j = 0; j < k; j++) {
                    /* heap starts full of FLT_MAX: a value is only admitted if FLT_MAX > v */
                    if (Ds[j] < FLT_MAX) topk_push(&t, Ds[j], Is[j]);
                }
            }
            topk_emit(&t, outD + qi * k, outI + qi * k, ORC_METRIC_L2);
        }
        free(buf);
    }
    return 0;
}

/* ------------------------------------------------------------------ */
/* training (faiss Clustering / ProductQuantizer::train restated:      */
/* Lloyd iterations, random-subset init, empty clusters split from a   */
/* large one).  Not on the timed path; used to build CPU-side indexes. */
/* ------------------------------------------------------------------ */
static uint64_t splitmix64(uint64_t *s) {
    uint64_t z = (*s += 0x9e3779b97f4a7c15ull);
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
}

int orc_kmeans(int d, int64_t n, const float *x, int64_t k, int niter, uint64_t seed,
               float *cent) {
    if (n < k) return -1;
    /* init: random subset without replacement (partial Fisher-Yates) */
    int64_t *perm = (int64_t *)malloc(sizeof(int64_t) * (size_t)n);
    for (int64_t i = 0; i < n; i++) perm[i] = i;
    uint64_t s = seed;
    for (int64_t i = 0; i < k; i++) {
        int64_t j = i + (int64_t)(splitmix64(&s) % (uint64_t)(n - i));
        int64_t tmp = perm[i];
        perm[i] = perm[j];
        perm[j] = tmp;
        memcpy(cent + i * d, x + perm[i] * d, sizeof(float) * (size_t)d);
    }
    free(perm);
    int64_t *assign = (int64_t *)malloc(sizeof(int64_t) * (size_t)n);
    double *sum = (double *)malloc(sizeof(double) * (size_t)(k * d));
    int64_t *cnt = (int64_t *)malloc(sizeof(int64_t) * (size_t)k);
    for (int it = 0; it < niter; it++) {
        orc_assign(ORC_METRIC_L2, d, k, cent, n, x, assign);
        memset(sum, 0, sizeof(double) * (size_t)(k * d));
        memset(cnt, 0, sizeof(int64_t) * (size_t)k);
        for (int64_t i = 0; i < n; i++) {
            int64_t c = assign[i];
            cnt[c]++;
            for (int kk = 0; kk < d; kk++) sum[c * d + kk] += x[i * d + kk];
        }
        for (int64_t c = 0; c < k; c++)
            if (cnt[c] > 0)
                for (int kk = 0; kk < d; kk++) cent[c * d + kk] = (float)(sum[c * d + kk] / cnt[c]);
        /* empty clusters: split the currently largest one (faiss splits a big cluster
         * with a +-1/1024 perturbation) */
        for (int64_t c = 0; c < k; c++) {
            if (cnt[c] != 0) continue;
            int64_t big = 0;
            for (int64_t c2 = 1; c2 < k; c2++)
                if (cnt[c2] > cnt[big]) big = c2;
            for (int kk = 0; kk < d; kk++) {
                float v = cent[big * d + kk];
                float eps = 1.f / 1024.f;
                if (kk % 2 == 0) {
                    cent[c * d + kk] = v * (1 + eps);
                    cent[big * d + kk] = v * (1 - eps);
                } else {
                    cent[c * d + kk] = v * (1 - eps);
                    cent[big * d + kk] = v * (1 + eps);
                }
            }
            cnt[c] = cnt[big] / 2;
            cnt[big] -= cnt[c];
        }
    }
    free(assign);
    free(sum);
    free(cnt);
    return 0;
}

int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
