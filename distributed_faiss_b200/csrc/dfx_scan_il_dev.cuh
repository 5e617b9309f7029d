// dfx_scan_il_dev.cuh -- device code of dfx_scan_il.cu (also compiled by the CPU emulator, tests/emu/).
// K4 v2: inverted-list scan of PQ codes, lane-per-subquantizer, on an
// interleaved code layout (M == 32).
//
// Replaces the inner loop of faiss IndexIVFPQ::search (reached from reference
// distributed_faiss/index.py:257) -- `dis = dis0 + sum_m table[m][code[m]]` over every code of
// every probed list.
//
// Why a second layout: with one vector per lane (v1, dfx_search.cu) the 32 lanes of a warp read
// table[m][code] for the SAME m and 32 random codes -> random shared-memory bank conflicts
// (~3.4 wavefronts per lookup), which caps the scan at ~1/3 of the HBM roofline.  Here the table
// is stored transposed ([code][m], bank == m) and the 32 lanes always read 32 DIFFERENT m:
// 8 lanes share a vector, lane (u,i) owns subquantizers {i, i+8, i+16, i+24} of the 8 vectors of
// group u and looks them up in the rotated order j = (t + u) & 3, so at every step the warp
// touches m = i + 8*((t+u)&3): all 32 banks, conflict-free by construction.
// The canonical halving tree of oracle pq_sum (s[x] += s[x+off], off = 16,8,4,2,1) is evaluated
// as: in-lane (y0+y2)+(y1+y3)  [levels 16 and 8; the rotation only swaps commutative operands],
// then a transposed butterfly over the 8 lanes of the group (levels 4,2,1: 7 shuffles per 32
// vectors, static register indices because rows are stored pre-permuted, r = w ^ i).
// Lane 8u+i ends with the full sum for vector 8u+i of the block (layout: dfx_il_byte()).
// Lists are padded to whole blocks; padding carries t = +inf so it can never enter a result.
#pragma once
#include "dfx_internal.h"
#include "dfx_topk.cuh"
#include "dfx_ptx.cuh"

// ------------------------------------------------------------------ layout conversion
__device__ __forceinline__ int64_t il_list_of_block(const int64_t* __restrict__ blk_off, int64_t nlist, int64_t blk) {
    int64_t lo = 0, hi = nlist;
    while (hi - lo > 1) {
        int64_t mid = (lo + hi) >> 1;
        if (blk_off[mid] <= blk) lo = mid; else hi = mid;
    }
    return lo;
}

// row-major (list-sorted) -> interleaved blocks.  one CTA per block.
__global__ void __launch_bounds__(256)
pq_rm_to_il_kernel(const int64_t* __restrict__ list_off, const int64_t* __restrict__ blk_off, int64_t nlist,
                   const uint8_t* __restrict__ codes, const float* __restrict__ tvals,
                   const int32_t* __restrict__ ids, uint8_t* __restrict__ il_codes, float* __restrict__ il_tvals,
                   int32_t* __restrict__ il_ids, int layout) {
    const int64_t blk = blockIdx.x;
    const int64_t l = il_list_of_block(blk_off, nlist, blk);
    const int64_t base = list_off[l] + (blk - blk_off[l]) * 32, end = list_off[l + 1];
    for (int e = threadIdx.x; e < 1024; e += 256) {
        const int v = e >> 5, m = e & 31;
        const int64_t i = base + v;
        il_codes[blk * 1024 + dfx_il_byte_of(layout, v, m)] = (i < end) ? codes[i * 32 + m] : (uint8_t)0;
    }
    if (threadIdx.x < 32) {
        const int64_t i = base + threadIdx.x;
        il_tvals[blk * 32 + threadIdx.x] = (i < end) ? tvals[i] : __int_as_float(0x7f800000);
        il_ids[blk * 32 + threadIdx.x] = (i < end) ? ids[i] : -1;
    }
}

__global__ void __launch_bounds__(256)
pq_il_to_rm_kernel(const int64_t* __restrict__ list_off, const int64_t* __restrict__ blk_off, int64_t nlist,
                   const uint8_t* __restrict__ il_codes, const float* __restrict__ il_tvals,
                   const int32_t* __restrict__ il_ids, uint8_t* __restrict__ codes, float* __restrict__ tvals,
                   int32_t* __restrict__ ids, int layout) {
    const int64_t blk = blockIdx.x;
    const int64_t l = il_list_of_block(blk_off, nlist, blk);
    const int64_t base = list_off[l] + (blk - blk_off[l]) * 32, end = list_off[l + 1];
    for (int e = threadIdx.x; e < 1024; e += 256) {
        const int v = e >> 5, m = e & 31;
        const int64_t i = base + v;
        if (i < end) codes[i * 32 + m] = il_codes[blk * 1024 + dfx_il_byte_of(layout, v, m)];
    }
    if (threadIdx.x < 32) {
        const int64_t i = base + threadIdx.x;
        if (i < end) {
            tvals[i] = il_tvals[blk * 32 + threadIdx.x];
            ids[i] = il_ids[blk * 32 + threadIdx.x];
        }
    }
}

// ------------------------------------------------------------------ the scan
// lutT: [nq][256][32] (transposed table, written by pq_prep_kernel)
constexpr int IL_THREADS = 256;  // 8 warps: 4 CTAs/SM = 32 warps/SM (the 64-register limit)
// SPLIT: block layout 3 -- the same words as layout 1, but the two 16-byte halves of a lane's 32
// bytes are stored 512 bytes apart ([half][lane][16]) so that each 128-bit load of the warp is one
// contiguous 512-byte run (layout 1: lanes 32 bytes apart, every load touches all 8 lines of the
// block and each sector crosses the L2->L1 crossbar twice).
template <bool SPLIT>
__device__ __forceinline__ void
scan_pq_il_body(const float* __restrict__ lutT, const float* __restrict__ dis0, const int32_t* __restrict__ keys,
                int nprobe, int G, int ngroups, const int64_t* __restrict__ blk_off,
                const uint4* __restrict__ il_codes, const float* __restrict__ il_tvals,
                const int32_t* __restrict__ il_ids, int k, int cap, uint64_t* __restrict__ part) {
    DFX_DYN_SMEM(unsigned char, smem_raw, 16);
    float* s_lut = reinterpret_cast<float*>(smem_raw);                       // [256][32]
    uint64_t* s_buf = reinterpret_cast<uint64_t*>(smem_raw + 256 * 32 * 4);  // 8 warps x cap
    const int64_t q = blockIdx.x / ngroups;
    const int g = blockIdx.x % ngroups;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // the 32 KB table of this query arrives by ONE bulk async copy (TMA engine, mbarrier
    // completion): no LDG/STS traffic on the LSU pipe that the lookups need
    __shared__ __align__(8) uint64_t s_lut_bar;
    __shared__ unsigned int s_cta_key;  // CTA-wide admission bound (see WarpTopK::cta_key)
    if (tid == 0) {
        s_cta_key = 0xff800000u;  // order-preserving key of +inf: no bound yet
        dfx_bulk_init_one(&s_lut_bar);
    }
    __syncthreads();
    if (tid == 0) dfx_bulk_issue(s_lut, lutT + q * 8192, 32768u, &s_lut_bar);
    WarpTopK wt;
    wt.init(s_buf + (size_t)warp * cap, cap, k, &s_cta_key);
    dfx_bulk_wait(&s_lut_bar);
    // byte offsets (inside a 128-byte table row) of the 4 subquantizers this lane looks up, in
    // lookup order: m_t = i + 8*((t+u)&3)
    const uint32_t li = lane & 7, lu = lane >> 3;
    const uint32_t lut_base = dfx_smem_addr(s_lut);
    uint32_t moff[4];
#pragma unroll
    for (int t = 0; t < 4; t++) moff[t] = lut_base + (li + 8u * ((t + lu) & 3u)) * 4u;

    // one 32-vector block: 32 conflict-free table lookups, the in-lane part of the tree, the
    // 8-lane butterfly, then the admission test.  `w` = this lane's 8 code words.
    auto process = [&](const uint4& ca, const uint4& cb, float tv, uint32_t my_id, float d0) {
        const uint32_t w[8] = {ca.x, ca.y, ca.z, ca.w, cb.x, cb.y, cb.z, cb.w};
        float a[8];
#pragma unroll
        for (int r = 0; r < 8; r++) {
            float y[4];
#pragma unroll
            for (int t = 0; t < 4; t++) {
                // byte t of the word (one PRMT), times the 128-byte row pitch plus this lane's
                // column offset (one IMAD)
                const uint32_t code = __byte_perm(w[r], 0u, 0x4440u + (uint32_t)t);
                const uint32_t addr = code * 128u + moff[t];
                y[t] = dfx_lds_f32(addr);
            }
            a[r] = (y[0] + y[2]) + (y[1] + y[3]);  // tree levels 16 and 8
        }
#pragma unroll
        for (int off = 4; off >= 1; off >>= 1) {  // tree levels 4, 2, 1 across the 8 lanes
#pragma unroll
            for (int r = 0; r < off; r++) a[r] = a[r] + __shfl_xor_sync(0xffffffffu, a[r + off], off);
        }
        const float v = d0 + (tv + a[0]);  // vector `lane` of this block; padding has tv = +inf
        uint32_t sec = 0;
        const bool want = (v + 0.0f <= wt.cta_bound()) && wt.admits(v, [&] { return my_id; }, sec);
        wt.push_lanes(want, v, sec);
    };

    constexpr int S = IL_THREADS / 32;  // block stride of one warp
    const int p_end = min(nprobe, (g + 1) * G);
    for (int p = g * G; p < p_end; p++) {
        const int l = keys[q * nprobe + p];
        if (l < 0) continue;
        const float d0 = dis0[q * nprobe + p];
        const int64_t b0 = blk_off[l] + warp;
        const int64_t nb = (blk_off[l + 1] - b0 + S - 1) / S;  // blocks this warp owns in the list
        if (nb <= 0) continue;
        // this lane's slice of block i of the warp: codes 2 x 16 B, t 4 B, id 4 B, streamed one
        // block ahead (an id fetched only on admission would put a DRAM latency on the critical
        // path of every admission)
        const uint4* pc = il_codes + b0 * 64 + (SPLIT ? lane : lane * 2);
        constexpr int HALF = SPLIT ? 32 : 1;  // distance (in uint4) between a lane's two halves
        const float* pt = il_tvals + b0 * 32 + lane;
        const int32_t* pi = il_ids + b0 * 32 + lane;
        uint4 xa = dfx_ld_stream(pc), xb = dfx_ld_stream(pc + HALF);
        float xt = dfx_ld_stream_f(pt);
        uint32_t xi = dfx_ld_stream_u(pi);
        for (int64_t i = 0; i < nb; i++) {
            const uint4 ca = xa, cb = xb;
            const float ct = xt;
            const uint32_t ci = xi;
            if (i + 1 < nb) {  // next block of this warp, in flight while this one is processed
                pc += S * 64;
                pt += S * 32;
                pi += S * 32;
                xa = dfx_ld_stream(pc);
                xb = dfx_ld_stream(pc + HALF);
                xt = dfx_ld_stream_f(pt);
                xi = dfx_ld_stream_u(pi);
            }
            process(ca, cb, ct, ci, d0);
        }
    }
    cta_merge_and_write<IL_THREADS>(wt, s_buf, cap, k, part + ((int64_t)q * ngroups + g) * k);
}

__global__ void __launch_bounds__(IL_THREADS, 4)
scan_pq_il_kernel(const float* __restrict__ lutT, const float* __restrict__ dis0, const int32_t* __restrict__ keys,
                  int nprobe, int G, int ngroups, const int64_t* __restrict__ blk_off,
                  const uint4* __restrict__ il_codes, const float* __restrict__ il_tvals,
                  const int32_t* __restrict__ il_ids, int k, int cap, uint64_t* __restrict__ part) {
    scan_pq_il_body<false>(lutT, dis0, keys, nprobe, G, ngroups, blk_off, il_codes, il_tvals, il_ids, k, cap, part);
}
// EXPERIMENTAL (scan_variant = 3): the same kernel on block layout 3
__global__ void __launch_bounds__(IL_THREADS, 4)
scan_pq_il_split_kernel(const float* __restrict__ lutT, const float* __restrict__ dis0,
                        const int32_t* __restrict__ keys, int nprobe, int G, int ngroups,
                        const int64_t* __restrict__ blk_off, const uint4* __restrict__ il_codes,
                        const float* __restrict__ il_tvals, const int32_t* __restrict__ il_ids, int k, int cap,
                        uint64_t* __restrict__ part) {
    scan_pq_il_body<true>(lutT, dis0, keys, nprobe, G, ngroups, blk_off, il_codes, il_tvals, il_ids, k, cap, part);
}

