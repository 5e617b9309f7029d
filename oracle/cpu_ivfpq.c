/*
 * cpu_ivfpq.c -- TIMED CPU BASELINE of the IVF-PQ search path (bench infrastructure, NOT product
 * code, NOT the parity checker).
 *
 * The parity oracle (dfx_oracle.c) is a scalar, bit-exact CHECKER: dependent fmaf chains in the
 * GPU's canonical order, one heap push per centroid.  Timing it says nothing about what the
 * reference's CPU path costs.  This file is the throughput-oriented restatement of the same
 * faiss path (IndexIVFPQ::search as reached from reference distributed_faiss/index.py:257,
 * builder index.py:43-48) that bench.py's `cpu_baseline` leg and `--impl reference` arm time on
 * the GPU box's host cores (faiss itself is absent: setup.py:32 faiss-cpu, un-vendored):
 *
 *   coarse quantizer   ||c||^2 - 2 q.c from ONE sgemm per query block (the caller runs the BLAS
 *                      sgemm -- MKL through torch.mm, all host threads -- exactly what faiss's
 *                      IndexFlatL2 does for nq >= 20), then cpu_select_probes(): threaded
 *                      top-nprobe per row;
 *   table              lut[m][j] = -2 <q_m, P[m][j]>   (faiss's per-query inner-product table of
 *                      the precomputed-table decomposition), OpenMP over queries;
 *   scan               dis = dis0 + t[v] + sum_m lut[m][code[v][m]] over every code of every probed
 *                      list (row-major 32-byte codes, 4 independent accumulators, list-order
 *                      prefetch), bounded-heap top-k, OpenMP over queries (dynamic schedule).
 *
 * Summation order is whatever is fastest (pairwise, 4 accumulators): results are validated
 * against the oracle within 1e-4 relative on distances and on id agreement
 * (tests/test_oracle.py::test_cpu_baseline_matches_oracle), NOT bit for bit.
 *
 * Build: oracle/Makefile target libdfx_cpu_baseline.so (x86-64-v3, travels with the snapshot);
 * bench.py rebuilds it with -march=native on the box when gcc is there.
 * Only tests/ and bench.py's cpu_baseline / --impl reference legs may load this library.
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

int cpu_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

void cpu_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

/* top-nprobe smallest of vals[r][j] = cnorm[j] - 2 * G[r][j] per row (ties: smaller column).
 * G: [nrows][ld] inner products from sgemm.  keys: [nrows][nprobe] ascending by value.
 * A row keeps its candidates in a small sorted array; the scan over nlist is a compare against
 * the current worst, which the compiler vectorises into a mostly-not-taken branch. */
void cpu_select_probes(int64_t nrows, int64_t nlist, int64_t ld, const float *G, const float *cnorm,
                       int nprobe, int32_t *keys) {
#pragma omp parallel for schedule(static)
    for (int64_t r = 0; r < nrows; r++) {
        const float *g = G + r * ld;
        float bv[512];
        int32_t bi[512];
        int np = nprobe > 512 ? 512 : nprobe, cnt = 0;
        float worst = FLT_MAX;
        for (int64_t j = 0; j < nlist; j++) {
            const float v = cnorm[j] - 2.0f * g[j];
            if (cnt == np && !(v < worst)) continue;
            int p = cnt < np ? cnt++ : np - 1;
            while (p > 0 && bv[p - 1] > v) {
                bv[p] = bv[p - 1];
                bi[p] = bi[p - 1];
                p--;
            }
            bv[p] = v;
            bi[p] = (int32_t)j;
            if (cnt == np) worst = bv[np - 1];
        }
        for (int p = 0; p < nprobe; p++) keys[r * nprobe + p] = p < cnt ? bi[p] : -1;
    }
}

typedef struct {
    float v;
    int64_t id;
} hent;

static inline int hless(float v, int64_t id, const hent *e) { return v < e->v || (v == e->v && id < e->id); }

/* keys: [nq][nprobe] probed lists (from cpu_select_probes).  codes row-major [n][M], tvals [n],
 * ids [n] int64, list_off [nlist+1].  M must be a multiple of 4, ksub = 256.
 * outD/outI [nq][k] ascending, (FLT_MAX, -1) padded.  ndis_out: number of codes scanned. */
void cpu_ivfpq_scan(int64_t nq, int d, const float *xq, const float *cent, int M, const float *codebooks,
                    const int64_t *list_off, const uint8_t *codes, const float *tvals, const int64_t *ids,
                    const int32_t *keys, int nprobe, int k, float *outD, int64_t *outI, int64_t *ndis_out) {
    const int dsub = d / M;
    int64_t ndis_total = 0;
#pragma omp parallel reduction(+ : ndis_total)
    {
        float *lut = (float *)aligned_alloc(64, (size_t)M * 256 * sizeof(float));
        hent *heap = (hent *)malloc((size_t)k * sizeof(hent));
#pragma omp for schedule(dynamic, 4)
        for (int64_t q = 0; q < nq; q++) {
            const float *x = xq + q * d;
            /* per-query table */
            for (int m = 0; m < M; m++) {
                const float *qm = x + m * dsub;
                const float *P = codebooks + (size_t)m * 256 * dsub;
                float *lm = lut + m * 256;
                for (int j = 0; j < 256; j++) {
                    float acc = 0.f;
                    for (int t = 0; t < dsub; t++) acc += qm[t] * P[j * dsub + t];
                    lm[j] = -2.0f * acc;
                }
            }
            int cnt = 0;
            float worst = FLT_MAX;
            int64_t worst_id = INT64_MAX;
            for (int p = 0; p < nprobe; p++) {
                const int32_t l = keys[q * nprobe + p];
                if (l < 0) continue;
                const float *c = cent + (size_t)l * d;
                float dis0 = 0.f;
                for (int t = 0; t < d; t++) {
                    const float df = x[t] - c[t];
                    dis0 += df * df;
                }
                const int64_t b = list_off[l], e = list_off[l + 1];
                ndis_total += e - b;
                for (int64_t v = b; v < e; v++) {
                    const uint8_t *cd = codes + v * M;
                    __builtin_prefetch(cd + 4 * M, 0, 0);
                    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
                    for (int m = 0; m < M; m += 4) {
                        s0 += lut[(m + 0) * 256 + cd[m + 0]];
                        s1 += lut[(m + 1) * 256 + cd[m + 1]];
                        s2 += lut[(m + 2) * 256 + cd[m + 2]];
                        s3 += lut[(m + 3) * 256 + cd[m + 3]];
                    }
                    const float dist = dis0 + (tvals[v] + ((s0 + s1) + (s2 + s3)));
                    if (cnt == k && !(dist < worst || (dist == worst && ids[v] < worst_id))) continue;
                    /* sorted insertion (k is small: 10 in the benchmark) */
                    int pos = cnt < k ? cnt++ : k - 1;
                    const int64_t id = ids[v];
                    while (pos > 0 && hless(dist, id, &heap[pos - 1])) {
                        heap[pos] = heap[pos - 1];
                        pos--;
                    }
                    heap[pos].v = dist;
                    heap[pos].id = id;
                    if (cnt == k) {
                        worst = heap[k - 1].v;
                        worst_id = heap[k - 1].id;
                    }
                }
            }
            for (int j = 0; j < k; j++) {
                outD[q * k + j] = j < cnt ? heap[j].v : FLT_MAX;
                outI[q * k + j] = j < cnt ? heap[j].id : -1;
            }
        }
        free(lut);
        free(heap);
    }
    if (ndis_out) *ndis_out = ndis_total;
}
