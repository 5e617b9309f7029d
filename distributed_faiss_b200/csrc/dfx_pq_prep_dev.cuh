// dfx_pq_prep_dev.cuh -- K3 device code (included by dfx_search.cu; also compiled by the CPU
// emulator, tests/emu/).
#pragma once
#include "dfx_common.cuh"
#include "dfx_ptx.cuh"

// =====================================================================================
// K3: pq_prep.  lut[q][m][j] = -2 * ip_seq(q_m, P[m][j]);  dis0[q][p] = warp-dot ||q - c||^2
// =====================================================================================
// transposed != 0 (M == 32): the table is written as lut[q][j][m] (1) or with 64-column rows
// lut[q][j][c], m = c & 31 (2) for the block-layout scans (dfx_scan_il.cu / dfx_scan_il2.cu); it
// is staged through padded shared memory so both the codebook reads and the global writes stay
// coalesced.
__global__ void __launch_bounds__(256)
pq_prep_kernel(const float* __restrict__ Q, int d, int M, int ksub, int dsub,
               const float* __restrict__ codebooks, const float* __restrict__ cent,
               const int32_t* __restrict__ keys, int nprobe, float* __restrict__ lut,
               float* __restrict__ dis0, int transposed) {
    DFX_DYN_SMEM0(float, s_q);
    float* s_t = s_q + ((d + 3) / 4) * 4;  // transposed mode: [M][ksub + 1]
    const int64_t q = blockIdx.x;
    for (int i = threadIdx.x; i < d; i += blockDim.x) s_q[i] = Q[q * d + i];
    __syncthreads();
    if (lut) {
        const int tot = M * ksub;
#pragma unroll 4
        for (int idx = threadIdx.x; idx < tot; idx += blockDim.x) {
            const int m = (ksub == 256) ? (idx >> 8) : (idx / ksub);
            const float* p = codebooks + (size_t)idx * dsub;
            const float* qm = s_q + m * dsub;
            float acc = 0.f;
            if (dsub == 4) {  // same seq-k order, one 16-byte load
                const float4 pv = __ldg(reinterpret_cast<const float4*>(p));
                acc = __fmaf_rn(qm[0], pv.x, acc);
                acc = __fmaf_rn(qm[1], pv.y, acc);
                acc = __fmaf_rn(qm[2], pv.z, acc);
                acc = __fmaf_rn(qm[3], pv.w, acc);
            } else {
                for (int t = 0; t < dsub; t++) acc = __fmaf_rn(qm[t], p[t], acc);
            }
            if (transposed) s_t[m * (ksub + 1) + (idx - m * ksub)] = -2.f * acc;
            else lut[q * tot + idx] = -2.f * acc;
        }
        if (transposed == 1) {  // M == 32 here: [code][m]
            __syncthreads();
            for (int o = threadIdx.x; o < tot; o += blockDim.x) {
                const int j = o >> 5, m = o & 31;
                lut[q * tot + o] = s_t[m * (ksub + 1) + j];
            }
        } else if (transposed == 2) {  // M == 32, wide rows: [code][64], column c holds m = c & 31
            __syncthreads();           // (scan_pq_il2_kernel reads column lane + t without a wrap)
            for (int o = threadIdx.x; o < 2 * tot; o += blockDim.x) {
                const int j = o >> 6, m = o & 31;
                lut[q * 2 * tot + o] = s_t[m * (ksub + 1) + j];
            }
        }
    }
    if (dis0) {
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
        for (int p = warp; p < nprobe; p += nw) {
            int l = keys[q * nprobe + p];
            float acc = 0.f;
            if (l >= 0) {
                const float* c = cent + (size_t)l * d;
                for (int base = 4 * lane; base < d; base += 128) {
                    float4 cv = *reinterpret_cast<const float4*>(c + base);
                    float4 qv = *reinterpret_cast<const float4*>(s_q + base);
                    float df;
                    df = qv.x - cv.x; acc = __fmaf_rn(df, df, acc);
                    df = qv.y - cv.y; acc = __fmaf_rn(df, df, acc);
                    df = qv.z - cv.z; acc = __fmaf_rn(df, df, acc);
                    df = qv.w - cv.w; acc = __fmaf_rn(df, df, acc);
                }
            }
            acc = dfx_warp_butterfly(acc);
            if (lane == 0) dis0[q * nprobe + p] = acc;
        }
    }
}

// =====================================================================================
// K3 variant 2 (EXPERIMENTAL, dfx_set_param "prep_variant" = 2, off by default; M == 32,
// dsub == 4, block layouts only).  The profile of pq_prep_kernel at batch 4096 shows it bound by
// the 128 KB codebook every CTA pulls from L2 (537 MB per launch) plus the shared-memory
// transpose.  Here the codebook is read from a copy stored in the order the table is written,
// PT[j][m][4] (built once per index by cb_transpose_kernel), so table entries are produced in
// output order -- no shared-memory staging -- and one CTA serves QB queries with each codebook
// vector loaded once.  Same canonical arithmetic: acc = fma(q0,p0,0), fma(q1,p1,acc), ...; -2*acc.
// =====================================================================================
__global__ void cb_transpose_kernel(const float* __restrict__ cb, int M, int ksub, int dsub,
                                    float* __restrict__ cbT) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;  // over M * ksub * dsub
    if (i >= M * ksub * dsub) return;
    const int t = i % dsub, e = i / dsub, j = e % ksub, m = e / ksub;
    cbT[((size_t)j * M + m) * dsub + t] = cb[i];
}

// wide != 0: rows of 64 columns (block layout 2), else 32 (block layout 1)
template <int QB>
__global__ void __launch_bounds__(256)
pq_prep2_kernel(const float* __restrict__ Q, int64_t nq, int d, const float* __restrict__ cbT,
                const float* __restrict__ cent, const int32_t* __restrict__ keys, int nprobe,
                float* __restrict__ lut, float* __restrict__ dis0, int wide) {
    DFX_DYN_SMEM(float, s_q, 16);  // [QB][d]
    const int64_t q0 = (int64_t)blockIdx.x * QB;
    for (int i = threadIdx.x; i < QB * d; i += blockDim.x) {
        const int64_t q = q0 + i / d;
        s_q[i] = (q < nq) ? Q[q * d + (i % d)] : 0.f;
    }
    __syncthreads();
    if (lut) {
        const float4* pt4 = reinterpret_cast<const float4*>(cbT);
        for (int e = threadIdx.x; e < 32 * 256; e += blockDim.x) {  // e = j * 32 + m
            const int m = e & 31, j = e >> 5;
            const float4 pv = __ldg(pt4 + e);
#pragma unroll
            for (int qb = 0; qb < QB; qb++) {
                const int64_t q = q0 + qb;
                if (q >= nq) break;
                const float4 qm = *reinterpret_cast<const float4*>(s_q + qb * d + 4 * m);
                float acc = 0.f;
                acc = __fmaf_rn(qm.x, pv.x, acc);
                acc = __fmaf_rn(qm.y, pv.y, acc);
                acc = __fmaf_rn(qm.z, pv.z, acc);
                acc = __fmaf_rn(qm.w, pv.w, acc);
                const float val = -2.f * acc;
                if (wide) {
                    float* row = lut + q * 16384 + j * 64 + m;
                    row[0] = val;
                    row[32] = val;
                } else {
                    lut[q * 8192 + e] = val;
                }
            }
        }
    }
    if (dis0) {
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
        for (int pair = warp; pair < QB * nprobe; pair += nw) {
            const int qb = pair / nprobe, p = pair % nprobe;
            const int64_t q = q0 + qb;
            if (q >= nq) continue;
            const int l = keys[q * nprobe + p];
            float acc = 0.f;
            if (l >= 0) {
                const float* c = cent + (size_t)l * d;
                for (int base = 4 * lane; base < d; base += 128) {
                    float4 cv = *reinterpret_cast<const float4*>(c + base);
                    float4 qv = *reinterpret_cast<const float4*>(s_q + qb * d + base);
                    float df;
                    df = qv.x - cv.x; acc = __fmaf_rn(df, df, acc);
                    df = qv.y - cv.y; acc = __fmaf_rn(df, df, acc);
                    df = qv.z - cv.z; acc = __fmaf_rn(df, df, acc);
                    df = qv.w - cv.w; acc = __fmaf_rn(df, df, acc);
                }
            }
            acc = dfx_warp_butterfly(acc);
            if (lane == 0) dis0[q * nprobe + p] = acc;
        }
    }
}
