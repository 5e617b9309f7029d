// dfx_select.cuh -- exact k-selection of one row per CTA under the total order
// (value asc, secondary asc) on 64-bit composites (dfx_common.cuh).
//
// Used for: coarse top-nprobe (reference: quantizer.search inside IndexIVF::search),
// flat top-k, the per-query reduction of the per-list partial results, and the
// cross-shard merge K6 (reference client.py:265-310).
//
// Algorithm: MSB-first radix select (8-bit digits, early exit as soon as the chosen bucket
// is entirely needed) finds the k'-th smallest composite T (composites are unique within a
// row), elements <= T are collected into shared memory and bitonic-sorted.  Rows with at
// most `sort_cap` elements skip the radix passes and are sorted directly.
#pragma once
#include "dfx_common.cuh"
#include "dfx_ptx.cuh"

template <int THREADS>
__device__ __forceinline__ void dfx_block_bitonic_sort(uint64_t* s, int P) {
    for (int size = 2; size <= P; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            __syncthreads();
            for (int i = threadIdx.x; i < (P >> 1); i += THREADS) {
                int pos = 2 * i - (i & (stride - 1));
                int partner = pos + stride;
                bool up = ((pos & size) == 0);
                uint64_t a = s[pos], b = s[partner];
                if ((a > b) == up) {
                    s[pos] = b;
                    s[partner] = a;
                }
            }
        }
    }
    __syncthreads();
}

// Loader:  __device__ uint64_t operator()(int64_t row, int e) const   (DFX_COMP_NONE = no candidate)
// Writer:  __device__ void operator()(int64_t row, int j, uint64_t comp) const   (j in [0,k))
// dynamic smem: P * 8 bytes with P = pow2 >= max(k, min(n, sort_cap))
// One row by the whole CTA (every thread calls it with the same arguments; s_out = P * 8 bytes of
// shared memory).  A caller that selects several rows in a loop puts a __syncthreads() between them.
template <int THREADS, class Loader, class Writer>
__device__ __forceinline__ void dfx_select_row(const Loader& ld, const Writer& wr, int64_t row, int n, int k, int P,
                                               int sort_cap, uint64_t* s_out) {
    __shared__ int s_hist[256];
    __shared__ int s_cnt;
    __shared__ int s_valid;
    __shared__ uint64_t s_prefix;
    __shared__ int s_krem;
    __shared__ int s_done;

    const int tid = threadIdx.x;

    if (n <= sort_cap) {
        // small row: sort everything
        for (int e = tid; e < P; e += THREADS) s_out[e] = (e < n) ? ld(row, e) : DFX_COMP_NONE;
        dfx_block_bitonic_sort<THREADS>(s_out, P);
        for (int j = tid; j < k; j += THREADS) wr(row, j, (j < P) ? s_out[j] : DFX_COMP_NONE);
        return;
    }

    // ---- radix select of the k'-th smallest composite
    if (tid == 0) {
        s_prefix = 0;
        s_krem = k;
        s_done = 0;
        s_valid = 0;
    }
    uint64_t T = DFX_COMP_NONE;
    int kprime = k;
    for (int pass = 0; pass < 8; pass++) {
        const int shift = 56 - 8 * pass;
        for (int b = tid; b < 256; b += THREADS) s_hist[b] = 0;
        __syncthreads();
        const uint64_t prefix = s_prefix;
        int local_valid = 0;
        for (int e = tid; e < n; e += THREADS) {
            uint64_t c = ld(row, e);
            if (pass == 0) {
                local_valid += (c != DFX_COMP_NONE);
                atomicAdd(&s_hist[(int)(c >> 56)], 1);
            } else if ((c >> (shift + 8)) == prefix) {
                atomicAdd(&s_hist[(int)((c >> shift) & 255)], 1);
            }
        }
        if (pass == 0 && local_valid) atomicAdd(&s_valid, local_valid);
        __syncthreads();
        if (tid == 0) {
            if (pass == 0) {
                int kp = min(k, s_valid);
                s_krem = kp;
            }
            int krem = s_krem;
            if (krem <= 0) {
                s_done = 2;  // nothing to select
            } else {
                int cum = 0, b = 0;
                for (; b < 256; b++) {
                    int h = s_hist[b];
                    if (cum + h >= krem) break;
                    cum += h;
                }
                krem -= cum;
                uint64_t np_ = (prefix << 8) | (uint64_t)b;
                s_prefix = np_;
                s_krem = krem;
                if (s_hist[b] == krem || pass == 7) {
                    // the whole bucket is needed: threshold = bucket upper bound
                    s_done = 1;
                    s_prefix = (shift == 0) ? np_ : ((np_ << shift) | ((1ull << shift) - 1ull));
                }
            }
        }
        __syncthreads();
        if (s_done) break;
    }
    if (s_done == 2) {
        for (int j = tid; j < k; j += THREADS) wr(row, j, DFX_COMP_NONE);
        return;
    }
    T = s_prefix;
    kprime = min(k, s_valid);

    // ---- collect + sort
    if (tid == 0) s_cnt = 0;
    for (int e = tid; e < P; e += THREADS) s_out[e] = DFX_COMP_NONE;
    __syncthreads();
    for (int e = tid; e < n; e += THREADS) {
        uint64_t c = ld(row, e);
        if (c <= T && c != DFX_COMP_NONE) {
            int pos = atomicAdd(&s_cnt, 1);
            if (pos < P) s_out[pos] = c;
        }
    }
    __syncthreads();
    dfx_block_bitonic_sort<THREADS>(s_out, P);
    for (int j = tid; j < k; j += THREADS) wr(row, j, (j < kprime) ? s_out[j] : DFX_COMP_NONE);
}

template <int THREADS, class Loader, class Writer>
__global__ void __launch_bounds__(THREADS) dfx_select_rows_kernel(Loader ld, Writer wr, int n, int k,
                                                                 int P, int sort_cap) {
    DFX_DYN_SMEM(unsigned char, dfx_sel_smem, 16);
    dfx_select_row<THREADS>(ld, wr, (int64_t)blockIdx.x, n, k, P, sort_cap, reinterpret_cast<uint64_t*>(dfx_sel_smem));
}

// host-side launch helper
template <int THREADS, class Loader, class Writer>
static inline void dfx_launch_select(Loader ld, Writer wr, int64_t nrows, int n, int k, cudaStream_t st) {
    if (nrows <= 0) return;
    DFX_REQUIRE(k >= 1 && k <= 4096, "k-selection supports 1 <= k <= 4096");
    const int sort_cap = 2048;
    int base = (n <= sort_cap) ? (n > k ? n : k) : k;
    int P = dfx_next_pow2(base < 2 ? 2 : base);
    size_t smem = (size_t)P * 8;
    auto kern = dfx_select_rows_kernel<THREADS, Loader, Writer>;
    DFX_LAUNCH(kern, (unsigned)nrows, THREADS, smem, st, ld, wr, n, k, P, sort_cap);
}
