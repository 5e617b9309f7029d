// dfx_tc.cu -- K1: the coarse quantizer on 5th-generation tensor cores (tcgen05 + TMEM + TMA).
//
// Replaces the dense query x centroid contraction inside `quantizer.search(nq, x, nprobe)`
// (faiss IndexIVF::search, reached from reference distributed_faiss/index.py:257) and
// `quantizer.assign` (IndexIVF::add, index.py:425).
//
// Scheme ("screen on tensor cores, decide in canonical fp32"):
//   1. both operands are scaled by a power of two (queries per row, the table as a whole, so that
//      every row norm lands in [2^13, 2^14)) and rounded to fp16; q.c ~= (qh.ch) / (sq sc) is
//      accumulated in fp32 in TMEM by ONE tcgen05.mma per k-step.  The rounding error is bounded
//      rigorously: |q.c - screen| <= |q||c| (2^-10 + 2^-13)  (two 2^-11 relative roundings per
//      product, fp32 accumulation), see screen_tol.  (Round 1 used a 3-MMA bf16 hi/lo split with a
//      1e-5 error; the decide stage makes the result exact for ANY bounded error, and a model of the
//      bench data put the extra exact evaluations of the 100x looser bound at +0.5 per query, for
//      a third of the tensor-core work.)
//   2. the epilogue turns each 128x128 accumulator tile into ranking values
//      (L2: |c|^2 - 2 q.c, IP: -q.c) and keeps only the MINIMUM of every group of 32
//      consecutive centroids -> gmin[nq][nlist/32]  (32x less traffic than the full matrix).
//   3. the G = nprobe + margin groups with the smallest minima are selected exactly
//      (dfx_select.cuh).  Every true top-nprobe centroid lies in a group whose minimum is within
//      the tolerance of the nprobe-th smallest minimum; if the G-th selected group is still within
//      it (more than `margin` near-ties, e.g. duplicate centroids) the row is re-done exactly by
//      tc_exact_rows_kernel -- nothing is ever dropped silently.
//   4. the candidate centroids are re-evaluated in the CANONICAL fp32 order
//      (seq-k FMA, identical to oracle/dfx_oracle.c) and the final top-nprobe / argmin is
//      taken on those exact values -> the probe lists are bit-identical to the oracle's.
//
// One CTA = 6 warps: warp 0 TMA producer, warp 1 TMEM allocator + single-thread MMA issuer,
// warps 2..5 epilogue (one TMEM lane == one query row per thread).  d <= 256: A (the query tile)
// stays resident in shared memory and B (centroids) streams through a ring of 16 KB stages;
// larger d (config C4, d = 768): A and B k-atoms stream through the ring together.  Two TMEM
// accumulator buffers overlap the epilogue of tile i with the MMAs of tile i+1.
#include "dfx_internal.h"
#include <utility>
#include "dfx_select.cuh"
#include "dfx_topk.cuh"
#include "dfx_ptx.cuh"
#ifndef DFX_EMU
#include <cuda.h>
#include <cuda_bf16.h>
#endif

#ifndef DFX_EMU  // tcgen05 / TMA / mbarrier PTX: hardware only (the CPU emulator uses emu_tc_screen below)
// ------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
    } while (!done);
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* tm, uint64_t* bar, int c0,
                                            int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"((uint64_t)tm), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_alloc(uint32_t* smem_dst, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)),
                 "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tc_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_mma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                            uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                     smem_u32(bar))
                 : "memory");
}
// (bits(v) & mask) | J in one LOP3 (mask in a register, J an immediate)
template <int J>
__device__ __forceinline__ float tc_pack_col(float v, uint32_t mask) {
    uint32_t r;
    asm("lop3.b32 %0, %1, %2, %3, 0xEA;" : "=r"(r) : "r"(__float_as_uint(v)), "r"(mask), "n"(J));
    return __uint_as_float(r);
}
__device__ __forceinline__ float tc_min3(float a, float b, float c) {
    float r;
    asm("min.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
    return r;
}
// The two smallest of the 32 packed ranking values of one TMEM chunk, pairwise with the 3-input
// min of sm_100 (FMNMX3): per pair lo / hi, m2 = min(m2, max(m1, lo), hi), m1 = min(m1, lo) --
// 5 FMNMX + 2 LOP3 per pair on the ALU pipe, which bounds the FAST kernel (round 1: 6 + 4).
template <int J>
__device__ __forceinline__ void tc_pair_step(const uint32_t (&r)[32], const float* __restrict__ cn, float mult,
                                             uint32_t mask, float& m1, float& m2) {
    const float a = tc_pack_col<J>(fmaf(__uint_as_float(r[J]), mult, cn[J]), mask);
    const float b = tc_pack_col<J + 1>(fmaf(__uint_as_float(r[J + 1]), mult, cn[J + 1]), mask);
    const float lo = fminf(a, b), hi = fmaxf(a, b);
    m2 = tc_min3(m2, fmaxf(m1, lo), hi);
    m1 = fminf(m1, lo);
}
template <int... P>
__device__ __forceinline__ void tc_two_smallest(const uint32_t (&r)[32], const float* __restrict__ cn, float mult,
                                                uint32_t mask, float& m1, float& m2,
                                                std::integer_sequence<int, P...>) {
    (tc_pair_step<2 * P>(r, cn, mult, mask, m1, m2), ...);
}
// 32 consecutive fp32 columns of this thread's TMEM lane
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// K-major, SWIZZLE_128B shared-memory matrix descriptor (rows of 128 B, 8-row atoms of 1024 B)
__device__ __forceinline__ uint64_t make_kmajor_sw128_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3ffffu) >> 4);  // start address, 16-byte units
    d |= (uint64_t)(1024u >> 4) << 32;             // stride byte offset between 8-row atoms
    d |= (uint64_t)1 << 46;                        // descriptor version (sm_100)
    d |= (uint64_t)2 << 61;                        // SWIZZLE_128B
    return d;
}

// instruction descriptor: D=f32, A=B=f16 (format 0), both K-major, M=128, N = 128 or 256
static constexpr uint32_t tc_idesc(uint32_t n) { return (1u << 4) | ((n >> 3) << 17) | ((128u >> 4) << 24); }

#endif  // !DFX_EMU

// ------------------------------------------------------------------ the kernel
namespace tc {
constexpr int TILE = 128;           // rows of A and of B per tile
constexpr int KATOM = 64;           // fp16 elements per 128-byte swizzle row
constexpr int ATOM_BYTES = TILE * KATOM * 2;  // 16 KB
constexpr int THREADS = 192;        // TMA warp + MMA warp + 4 epilogue warps
constexpr int MAX_RESIDENT_KATOMS = 4;  // query tile resident in shared memory up to 4 atoms (64 KB): FAST d <= 256, PRECISE d <= 128
// Screening precision (NPL = planes per operand):
//   FAST    NPL 1: fp16(x s); one MMA per k-step; error <= |q||c| (2^-10 + 2^-13)
//   PRECISE NPL 2: hi = fp16(x s), lo = fp16(x s - hi); q.c ~= qh.ch + ql.ch + qh.cl, three MMAs per
//           k-step; error <= |q||c| 2^-17 (the dropped ql.cl term is 2^-22, hi + lo carries 22 bits)
// shared-memory plan (offsets inside the 1024-aligned dynamic block)
//   resident A (NPL x katoms <= 4): A = NPL x katoms x 16 KB, then a ring of 4 B stages
//                          (FAST: 16 KB stages, d <= 128: 97 KB, two CTAs per SM;
//                           PRECISE: 32 KB stages = 256 centroids, 192 KB)
//   streamed A (larger d): a ring of 6 stages of (A atom + B atom) = 32 KB
struct Smem {
    static constexpr bool resident(int katoms, int npl) { return npl * katoms <= MAX_RESIDENT_KATOMS; }
    // tile width in centroids: resident PRECISE uses N = 256 (one B stage = two 16 KB atoms side by
    // side): three SS-mode M = N = 128 MMAs per k-step read 24 KB per 192 cycles -- the SM's whole
    // 128 B/clk, before the TMA fill -- while N = 256 re-uses each A read over twice the columns
    // (36 KB per 384 cycles); everything else keeps N = 128 (two accumulators of N = 256 are all
    // of TMEM, which excludes the two-CTAs-per-SM shape of FAST)
    static constexpr int tile_n(bool res, int npl) { return (res && npl == 2) ? 256 : 128; }
    static constexpr int nstage(bool res, int npl) { return res ? 4 : 6; }
    static constexpr int stage_bytes(bool res, int npl) { return res ? (tile_n(res, npl) / TILE) * ATOM_BYTES : 2 * ATOM_BYTES; }
    static constexpr int RING(bool res, int katoms, int npl) { return res ? npl * katoms * ATOM_BYTES : 0; }
    static constexpr int BARS(bool res, int katoms, int npl) {
        return RING(res, katoms, npl) + nstage(res, npl) * stage_bytes(res, npl);
    }
    static constexpr int total(bool res, int katoms, int npl) {
        return BARS(res, katoms, npl) + 512 + 2 * tile_n(res, npl) * 4;
    }
};
}  // namespace tc

// the power of two that brings a norm into [2^13, 2^14) (1 for a zero / non-finite norm; the
// exponent is clamped to +-45, norms beyond 2^+-45 are outside what the screening supports)
__host__ __device__ __forceinline__ float dfx_pow2_scale(float nrm) {
    if (!(nrm > 0.f) || !(nrm < 3.0e38f)) return 1.f;
    int e;
    frexpf(nrm, &e);  // nrm = m 2^e, m in [0.5, 1)
    e = e < -45 ? -45 : (e > 45 ? 45 : e);
    return ldexpf(1.f, 14 - e);
}

// fp32 rows -> fp16 rows scaled by a power of two, rows >= n zero filled; one warp per row.
//   row_scaled != 0 (queries): the row is scaled so that its norm lands in [2^13, 2^14) and
//       mult[row] = factor / (row scale * table_scale) -- what the screening epilogue multiplies
//       the accumulator with (factor = -2 for L2, -1 for inner product; all powers of two, exact);
//   row_scaled == 0 (a centroid table / the rows of a FLAT index): every row times table_scale.
// Elements more than 2^28 below the row norm fall into fp16's subnormal range and lose relative
// (not absolute) precision: an error of at most 2^-39 of the norm, far inside screen_tol.
// npl == 2 also writes the residual plane lo = fp16(x s - hi) at out + n_pad * d.
__global__ void __launch_bounds__(256)
f16_rows_kernel(const float* __restrict__ x, int64_t n, int64_t n_pad, int d, int row_scaled, float table_scale,
                float factor, int npl, __half* __restrict__ out, float* __restrict__ mult, float* __restrict__ norm2) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= n_pad) return;
    const bool real = row < n;
    float scale = table_scale;
    if (row_scaled) {
        float part = 0.f;
        if (real)
            for (int k = lane; k < d; k += 32) {
                const float v = x[row * d + k];
                part = __fmaf_rn(v, v, part);
            }
        const float n2 = dfx_warp_butterfly(part);
        scale = dfx_pow2_scale(sqrtf(n2));
        if (lane == 0) {
            mult[row] = factor / (scale * table_scale);
            if (norm2) norm2[row] = n2;
        }
    }
    for (int k = lane; k < d; k += 32) {
        const float v = real ? x[row * d + k] * scale : 0.f;
        const __half hi = __float2half_rn(v);
        out[row * d + k] = hi;
        if (npl == 2) out[n_pad * d + row * d + k] = __float2half_rn(v - __half2float(hi));
    }
}

#ifdef DFX_EMU
// ---- CPU emulator stand-in (tests/emu/): the screening kernel cannot be emulated instruction
// by instruction, so its RESULT is restated: the fp16 operands the product code prepared
// (f16_rows_kernel), fp32 accumulation (order not specified by the hardware either), the
// epilogue's ranking value fma(acc, mult[row], cn) and its packed (min, runner-up, arg-min) per
// 32 centroids.  Everything downstream (group selection, exact canonical re-evaluation, the
// overflow fallback, the drivers) is the product code.
static void emu_tc_screen(const __half* qh, const float* qmult, int64_t nq, int64_t nq_pad, const __half* ch,
                          int64_t nlist, int64_t nl_pad, int d, int npl, const float* cnorm, int metric, float* gmin,
                          float* gmin2, uint8_t* gargc, float* tmin, int ng) {
    const float big = 3.0e38f;
    std::vector<float> cf((size_t)npl * nl_pad * d), qf((size_t)npl * d);
    for (size_t i = 0; i < cf.size(); i++) cf[i] = __half2float(ch[i]);
    for (int64_t row = 0; row < nq; row++) {
        for (int pl = 0; pl < npl; pl++)
            for (int k = 0; k < d; k++) qf[(size_t)pl * d + k] = __half2float(qh[(size_t)pl * nq_pad * d + row * d + k]);
        for (int g = 0; g < ng; g++) {
            float m1 = big, m2 = big;
            for (int j = 0; j < 32; j++) {
                const int64_t col = (int64_t)g * 32 + j;
                float acc = 0.f;
                const float* h = &cf[(size_t)col * d];
                for (int k = 0; k < d; k++) acc += qf[k] * h[k];
                if (npl == 2) {
                    const float* l = &cf[(size_t)nl_pad * d + (size_t)col * d];
                    for (int k = 0; k < d; k++) acc += qf[d + k] * h[k] + qf[k] * l[k];
                }
                float cn = big;
                if (col < nlist) cn = (metric == DFX_METRIC_L2) ? cnorm[col] : 0.f;
                const float v = fmaf(acc, qmult[row], cn);
                uint32_t u;
                memcpy(&u, &v, 4);
                u = (u & ~31u) | (uint32_t)j;
                float vj;
                memcpy(&vj, &u, 4);
                m2 = fminf(m2, fmaxf(m1, vj));
                m1 = fminf(m1, vj);
            }
            uint32_t u1;
            memcpy(&u1, &m1, 4);
            gmin[row * ng + g] = m1;
            gmin2[row * ng + g] = m2;
            gargc[row * ng + g] = (uint8_t)(u1 & 31u);
        }
        for (int t = 0; t < ng / 4; t++) {  // the minimum of every tile of 4 groups
            const float* gm = gmin + row * ng + 4 * t;
            tmin[row * (ng / 4) + t] = fminf(fminf(gm[0], gm[1]), fminf(gm[2], gm[3]));
        }
    }
}
#else
// tmQ: fp16 [NPL * nq_pad, d] (scaled query rows; NPL == 2: hi plane then lo plane), tmC: fp16
// [NPL * nl_pad, d] (scaled table rows), d = 64 KATOMS
// qmult: float [nq_pad], the per-row multiplier of the accumulator (f16_rows_kernel)
// gmin / gmin2: float [nq][ng], gargc: u8 [nq][ng], ng = nl_pad/32
// KT: d / 64 at compile time for the resident shapes (1, 2, 4); 0 = streamed, k-atoms at run time
template <int KT, int METRIC, int NPL>
__global__ void __launch_bounds__(tc::THREADS, (KT != 0 && NPL == 1) ? 2 : 1)
tc_coarse_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmC, int nq,
                 int nq_pad, int nlist, int nl_pad, int katoms_rt, const float* __restrict__ cnorm,
                 const float* __restrict__ qmult, int ctiles_per_cta, float* __restrict__ gmin,
                 float* __restrict__ gmin2, uint8_t* __restrict__ gargc, float* __restrict__ tmin, int ng) {
    using namespace tc;
    constexpr bool RES = KT != 0;
    const int KATOMS = KT ? KT : katoms_rt;
    constexpr int NSTAGE = Smem::nstage(RES, NPL);
    constexpr int STAGE_BYTES = Smem::stage_bytes(RES, NPL);
    constexpr int TN = Smem::tile_n(RES, NPL);   // centroids per tile
    constexpr int TMEM_COLS = 2 * TN;            // two accumulator buffers
    constexpr uint32_t IDESC = tc_idesc(TN);
    // operand pairs accumulated per k-atom: FAST (qh, ch); PRECISE (qh, ch), (ql, ch), (qh, cl)
    constexpr int NCOMBO = NPL == 1 ? 1 : 3;
    extern __shared__ unsigned char smem_raw_tc[];
    unsigned char* smem = reinterpret_cast<unsigned char*>(
        (reinterpret_cast<uintptr_t>(smem_raw_tc) + 1023) & ~(uintptr_t)1023);  // SWIZZLE_128B atoms
    unsigned char* sA = smem;                            // RES: [plane][katom] x 16 KB
    unsigned char* sR = smem + Smem::RING(RES, KATOMS, NPL);  // ring: B atom (RES) or A atom + B atom
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Smem::BARS(RES, KATOMS, NPL));
    uint64_t* full = bars;                 // [NSTAGE]
    uint64_t* empty = bars + NSTAGE;       // [NSTAGE]
    uint64_t* a_full = bars + 2 * NSTAGE;  // [1]
    uint64_t* t_full = a_full + 1;         // [2]
    uint64_t* t_empty = t_full + 2;        // [2]
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(t_empty + 2);
    float* s_cn = reinterpret_cast<float*>(smem + Smem::BARS(RES, KATOMS, NPL) + 512);  // [2][TN]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int qt = blockIdx.y;
    const int ctiles = (nl_pad + TN - 1) / TN;  // (TN = 256: the last tile may be half outside; zero / masked)
    const int ct0 = blockIdx.x * ctiles_per_cta;
    const int ct1 = min(ctiles, ct0 + ctiles_per_cta);
    const int ntiles = ct1 - ct0;
    if (ntiles <= 0) return;

    if (threadIdx.x == 0) {
        for (int i = 0; i < NSTAGE; i++) {
            mbar_init(&full[i], 1);
            mbar_init(&empty[i], 1);
        }
        mbar_init(a_full, 1);
        for (int b = 0; b < 2; b++) {
            mbar_init(&t_full[b], 1);
            mbar_init(&t_empty[b], 4);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 1) tc_alloc(tmem_ptr, TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    if (warp == 0) {
        // ===================== TMA producer =====================
        // RES: per tile, the B atoms of the hi plane then (PRECISE) of the lo plane.
        // streamed: per tile and k-atom, the (A atom, B atom) pair of every combo.
        if (lane == 0) {
            if (RES) {
                mbar_expect_tx(a_full, NPL * KATOMS * ATOM_BYTES);
                for (int pl = 0; pl < NPL; pl++)
                    for (int ka = 0; ka < KATOMS; ka++)
                        tma_load_2d(sA + (pl * KATOMS + ka) * ATOM_BYTES, &tmQ, a_full, ka * KATOM,
                                    pl * nq_pad + qt * TILE);
            }
            int stage = 0, phase = 0;
            for (int t = 0; t < ntiles; t++) {
                const int crow = (ct0 + t) * TN;
                // RES: c = B plane; streamed: c = combo
                for (int c = 0; c < (RES ? NPL : NCOMBO); c++)
                for (int ka = 0; ka < KATOMS; ka++) {
                    mbar_wait(&empty[stage], phase ^ 1);
                    mbar_expect_tx(&full[stage], STAGE_BYTES);
                    unsigned char* dst = sR + stage * STAGE_BYTES;
                    if (!RES) {
                        tma_load_2d(dst, &tmQ, &full[stage], ka * KATOM, (c == 1 ? nq_pad : 0) + qt * TILE);
                        dst += ATOM_BYTES;
                    }
                    const int bpl = RES ? c : (c == 2 ? 1 : 0);
                    tma_load_2d(dst, &tmC, &full[stage], ka * KATOM, bpl * nl_pad + crow);
                    if (++stage == NSTAGE) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (one thread) =====================
        if (lane == 0) {
            if (RES) {
                mbar_wait(a_full, 0);
                tc_fence_after();
            }
            int stage = 0, phase = 0;
            const uint32_t a_base = smem_u32(sA);
            const uint32_t r_base = smem_u32(sR);
            for (int t = 0; t < ntiles; t++) {
                const int buf = t & 1;
                mbar_wait(&t_empty[buf], ((t >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + buf * TN;
#pragma unroll
                for (int c = 0; c < (RES ? NPL : NCOMBO); c++)
#pragma unroll
                for (int ka = 0; ka < KATOMS; ka++) {
                    mbar_wait(&full[stage], phase);
                    tc_fence_after();
                    const uint32_t st_addr = r_base + stage * STAGE_BYTES;
                    const uint32_t b_addr = RES ? st_addr : st_addr + ATOM_BYTES;
                    const uint32_t ah_addr = RES ? a_base + ka * ATOM_BYTES : st_addr;
                    const uint32_t al_addr = a_base + (KATOMS + ka) * ATOM_BYTES;  // RES && NPL == 2 only
#pragma unroll
                    for (int kk = 0; kk < KATOM / 16; kk++) {  // UMMA_K = 16 fp16 = 32 bytes
                        const uint64_t bd = make_kmajor_sw128_desc(b_addr + kk * 32);
                        tc_mma_f16(d_tmem, make_kmajor_sw128_desc(ah_addr + kk * 32), bd, IDESC,
                                   (c > 0 || ka > 0 || kk > 0) ? 1u : 0u);
                        // resident PRECISE: the hi-plane B stage also takes ql . ch
                        if (RES && NPL == 2 && c == 0)
                            tc_mma_f16(d_tmem, make_kmajor_sw128_desc(al_addr + kk * 32), bd, IDESC, 1u);
                    }
                    tc_commit(&empty[stage]);  // frees the stage once these MMAs have read it
                    if (++stage == NSTAGE) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
                tc_commit(&t_full[buf]);  // accumulator of this tile complete
            }
        }
    } else {
        // ===================== epilogue: TMEM -> per-group (min, runner-up, argmin) =============
        // The epilogue is bound by the half-rate ALU pipe (and, with one MMA per k-step, it is the
        // longest stage of the kernel), so it is kept to 3.5 ALU instructions per element
        // (tc_two_smallest): the column index (0..31) replaces the 5 low mantissa bits of the
        // screening value (a 2^-18 relative perturbation, inside the screening tolerance), which
        // makes the arg-min fall out of the minimum itself.
        const int quad = warp & 3;             // TMEM lane quadrant this warp may access
        const int row = quad * 32 + lane;      // query row inside the tile == TMEM lane
        const int64_t grow = (int64_t)qt * TILE + row;
        const int et = threadIdx.x - 64;       // 0..127
        const float big = 3.0e38f;             // out-of-range columns: finite, never selected
        const float mult = qmult[grow];        // padded rows carry a multiplier too
        uint32_t idx_mask;                     // ~31 in a REGISTER: (v & mask) | j is then one LOP3
        asm volatile("mov.u32 %0, 0xffffffe0;" : "=r"(idx_mask));
        for (int t = 0; t < ntiles; t++) {
            const int buf = t & 1;
            const int col0 = (ct0 + t) * TN;
#pragma unroll
            for (int e = et; e < TN; e += 128) {
                const int c = col0 + e;
                float cn = 0.f;
                if (METRIC == DFX_METRIC_L2 && c < nlist) cn = cnorm[c];
                s_cn[buf * TN + e] = (c < nlist) ? cn : big;
            }
            asm volatile("bar.sync 1, 128;" ::: "memory");
            mbar_wait(&t_full[buf], (t >> 1) & 1);
            tc_fence_after();
#pragma unroll
            for (int half = 0; half < TN / 128; half++) {  // 128 columns = 4 groups = one selection tile
                float gm[4], gm2[4];
                uint32_t ga = 0;
#pragma unroll
                for (int ch = 0; ch < 4; ch++) {
                    uint32_t r[32];
                    tc_ld32(tmem_base + buf * TN + half * 128 + ch * 32 + ((uint32_t)(quad * 32) << 16), r);
                    float m1 = big, m2 = big;
                    tc_two_smallest(r, s_cn + buf * TN + half * 128 + ch * 32, mult, idx_mask, m1, m2,
                                    std::make_integer_sequence<int, 16>{});
                    gm[ch] = m1;
                    gm2[ch] = m2;
                    ga |= (__float_as_uint(m1) & 31u) << (8 * ch);
                }
                if (half == TN / 128 - 1) {  // the accumulator has been read: hand it back
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&t_empty[buf]);
                }
                const int colh = col0 + half * 128;
                if (grow < nq && colh < nl_pad) {
                    const int64_t o = grow * ng + (colh >> 5);
                    *reinterpret_cast<float4*>(gmin + o) = make_float4(gm[0], gm[1], gm[2], gm[3]);
                    *reinterpret_cast<float4*>(gmin2 + o) = make_float4(gm2[0], gm2[1], gm2[2], gm2[3]);
                    *reinterpret_cast<uint32_t*>(gargc + o) = ga;
                    // the tile's minimum: the first level of the group selection (topg_collect_kernel)
                    tmin[grow * (ng >> 2) + (colh >> 7)] = fminf(fminf(gm[0], gm[1]), fminf(gm[2], gm[3]));
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tc_dealloc(tmem_base, TMEM_COLS);
    }
}
#endif  // DFX_EMU

// Which columns of the selected groups can still matter?  With t = value of the nprobe-th
// smallest group minimum and tol >= twice the screening error, every true top-nprobe centroid c
// has approx(c) <= t + tol; inside its group it is either the arg-min column or has
// approx(c) >= gmin2(group).  So a group is expanded to all 32 columns only if
// gmin2 <= t + tol, otherwise its arg-min column is the only candidate.
//
// Screening error, s = |q| max|c| (Cauchy-Schwarz bounds sum |q_k c_k| by it):
//   FAST (fp16 operands): every product carries two relative roundings of at most 2^-11
//     (2^-10 + 2^-22 together), the fp32 accumulation of the d products at most d 2^-24:
//       |q.c - screen| <= s (2^-10 + 2^-22 + d 2^-24)
//   PRECISE (hi + lo planes, three partial products): hi + lo represents an operand to 2^-22, the
//     dropped ql.cl term is 2^-22, 3d products are accumulated:
//       |q.c - screen| <= s (3 2^-22 + 3 d 2^-24)
// The L2 ranking value |c|^2 - 2 q.c doubles it and the tolerance must be twice the error of a
// value (threshold and candidate are both approximations): tol = 4 err (+5 % slack).  The packed
// column index perturbs a value by 2^-18 |v|, |v| <= max|c|^2 + 2 s: the 1e-5 term.
static float screen_tol_rel(int d, bool fast) {
    const double err = fast ? (1.0 / 1024 + 1.0 / 4194304 + d / 16777216.0) : (3.0 / 4194304 + 3.0 * d / 16777216.0);
    return (float)(4.2 * err);
}
__device__ __forceinline__ float screen_tol(float qn2, float cmax2, float rel) {
    const float s = sqrtf(qn2 * cmax2);
    return rel * s + 1e-5f * (cmax2 + 2.f * s) + 1e-30f;
}
// "the selected groups may not be all that matter": the G-th (last) selected group is itself
// within the tolerance, so an unselected group could be too -> the row is re-done exactly
__device__ __forceinline__ bool screen_overflow(const int32_t* __restrict__ groups, int64_t row, int G, int ng,
                                                const float* __restrict__ gmin, float thr) {
    if (G >= ng) return false;  // every group is selected
    const int gl = groups[row * G + G - 1];
    return gl >= 0 && gmin[row * ng + gl] <= thr;
}
// ovf = [overflow rows, rows that the FAST tolerance would overflow (counted by PRECISE launches
// for the AUTO mode), row list ...].  t = the nprobe-th smallest group minimum (inf: no pruning).
struct ScreenTol {
    float rel;       // of this launch's precision
    float rel_fast;  // > 0: also count the rows FAST would overflow
};
__device__ __forceinline__ float screen_threshold(const int32_t* __restrict__ groups, int64_t row, int G, int ng,
                                                  const float* __restrict__ gmin, float t, float qn2, float cmax2,
                                                  ScreenTol tol, int32_t* __restrict__ ovf) {
    const float thr = t + screen_tol(qn2, cmax2, tol.rel);
    if (screen_overflow(groups, row, G, ng, gmin, thr)) ovf[2 + atomicAdd(&ovf[0], 1)] = (int32_t)row;
    if (tol.rel_fast > 0.f && screen_overflow(groups, row, G, ng, gmin, t + screen_tol(qn2, cmax2, tol.rel_fast)))
        atomicAdd(&ovf[1], 1);
    return thr;
}

// The G (<= 32) smallest group minima of a row, ascending by (value, group): one warp per row.
//   1. every lane takes the minimum of its n/32 values; the G-th smallest of those 32 lane
//      minima, B, bounds the answer from above (G different lanes hold a value <= B);
//   2. the values <= B are collected (a few dozen at most in practice) into a per-warp buffer;
//   3. the buffer is bitonic-sorted by the warp and the first G entries are kept.
// Exact for any input; if more than CAP values are <= B (massive ties) the row falls back to
// G rounds of warp arg-min.  ~10x fewer instructions than G x n compare rounds.
// Two levels (fine != nullptr): `coarse` holds the minimum of every FINE_PER = 4 consecutive
// groups (one 128-column tile of the screening kernel, written by its epilogue) -- steps 1-3 run
// on the n / 4 tile minima, then the 4 G groups of the G smallest tiles are sorted and the first
// G kept.  The G smallest groups always lie in the G smallest tiles: a group g outside them
// would have G tiles, each holding a group with (minimum, index) below g's, before it.
template <int CAP>
__device__ __forceinline__ void warp_sort_buf(uint64_t* buf, int cnt, int lane, int& P) {
    P = 32;
    while (P < cnt) P <<= 1;
    for (int e = cnt + lane; e < P; e += 32) buf[e] = DFX_COMP_NONE;
    __syncwarp();
    for (int size = 2; size <= P; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int i = lane; i < (P >> 1); i += 32) {
                const int pos = 2 * i - (i & (stride - 1));
                const int partner = pos + stride;
                const bool up = ((pos & size) == 0);
                const uint64_t a = buf[pos], bb = buf[partner];
                if ((a > bb) == up) {
                    buf[pos] = bb;
                    buf[partner] = a;
                }
            }
            __syncwarp();
        }
    }
}
template <int CAP>
__global__ void __launch_bounds__(256)
topg_collect_kernel(const float* __restrict__ coarse, int64_t nq, int n, int G, const float* __restrict__ fine,
                    int n_fine, int32_t* __restrict__ groups) {
    constexpr int FINE_PER = 4;
    static_assert(CAP >= 32 * FINE_PER, "the expansion of 32 tiles must fit the buffer");
    __shared__ uint64_t s_buf[8][CAP];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * 8 + warp;
    if (row >= nq) return;
    const float* g = coarse + row * n;
    uint64_t* buf = s_buf[warp];
    // (n <= 512, the usual case -- 65 536 lists = 512 tiles: the row is loaded ONCE, 16 independent
    // loads per lane, and both passes run on registers; the loops below were bound by one L2
    // latency per iteration)
    constexpr int REGV = 16;
    const bool in_regs = n <= 32 * REGV;
    float rv[REGV];
    if (in_regs) {
#pragma unroll
        for (int i = 0; i < REGV; i++) {
            const int j = lane + 32 * i;
            rv[i] = (j < n) ? g[j] : 0.f;
        }
    }
    // 1. lane minima -> bound
    uint64_t lmin = DFX_COMP_NONE;
    if (in_regs) {
#pragma unroll
        for (int i = 0; i < REGV; i++) {
            const int j = lane + 32 * i;
            const uint64_t c = (j < n) ? dfx_comp(rv[i], (uint32_t)j) : DFX_COMP_NONE;
            lmin = c < lmin ? c : lmin;
        }
    } else {
        for (int j = lane; j < n; j += 32) {
            const uint64_t c = dfx_comp(g[j], (uint32_t)j);
            lmin = c < lmin ? c : lmin;
        }
    }
    // bitonic sort of the 32 lane minima across the warp (ascending by lane)
    uint64_t x = lmin;
#pragma unroll
    for (int size = 2; size <= 32; size <<= 1) {
#pragma unroll
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            const uint64_t y = __shfl_xor_sync(0xffffffffu, x, stride);
            const bool up = ((lane & size) == 0);
            const bool lower = ((lane & stride) == 0);
            const bool take_min = (lower == up);
            x = take_min ? (x < y ? x : y) : (x < y ? y : x);
        }
    }
    const int gth = min(G, 32) - 1;
    const uint64_t bound = __shfl_sync(0xffffffffu, x, gth);  // G-th smallest lane minimum
    // 2. collect everything <= bound
    int cnt = 0;
    bool overflow = false;
    for (int j0 = 0, it = 0; j0 < n; j0 += 32, it++) {
        const int j = j0 + lane;
        float gv = 0.f;
        if (in_regs) {
#pragma unroll
            for (int i = 0; i < REGV; i++) gv = (i == it) ? rv[i] : gv;  // rv[it] without local memory
        } else if (j < n) {
            gv = g[j];
        }
        const uint64_t c = (j < n) ? dfx_comp(gv, (uint32_t)j) : DFX_COMP_NONE;
        const bool want = c <= bound && c != DFX_COMP_NONE;
        const unsigned mask = __ballot_sync(0xffffffffu, want);
        if (mask) {
            const int pos = cnt + __popc(mask & ((1u << lane) - 1u));
            if (want && pos < CAP) buf[pos] = c;
            cnt += __popc(mask);
            if (cnt > CAP) overflow = true;
        }
    }
    int P = 32;
    if (!overflow) {
        warp_sort_buf<CAP>(buf, cnt, lane, P);  // 3. sort the candidates
    } else {
        // fallback: G rounds of "smallest composite above the previous pick", left in buf[0..G)
        uint64_t prev = 0;
        for (int r = 0; r < G; r++) {
            uint64_t best = DFX_COMP_NONE;
            for (int j = lane; j < n; j += 32) {
                const uint64_t c = dfx_comp(g[j], (uint32_t)j);
                if ((r == 0 || c > prev) && c < best) best = c;
            }
#pragma unroll
            for (int off = 16; off >= 1; off >>= 1) {
                const uint64_t o = __shfl_xor_sync(0xffffffffu, best, off);
                best = o < best ? o : best;
            }
            if (lane == 0) buf[r] = best;
            prev = best;
            if (best == DFX_COMP_NONE) prev = DFX_COMP_NONE - 1;
        }
        __syncwarp();
    }
    if (fine) {  // second level: the groups of the G smallest tiles
        const uint64_t tc_ = (lane < G && (overflow || lane < P)) ? buf[lane] : DFX_COMP_NONE;
        __syncwarp();
        const float* f = fine + row * n_fine;
#pragma unroll
        for (int j = 0; j < FINE_PER; j++) {
            const int64_t gi = (int64_t)(uint32_t)tc_ * FINE_PER + j;
            buf[lane * FINE_PER + j] =
                (tc_ != DFX_COMP_NONE && gi < n_fine) ? dfx_comp(f[gi], (uint32_t)gi) : DFX_COMP_NONE;
        }
        __syncwarp();
        warp_sort_buf<CAP>(buf, 32 * FINE_PER, lane, P);
        overflow = false;
    }
    for (int r = lane; r < G; r += 32) {
        const uint64_t c = (overflow || r < P) ? buf[r] : DFX_COMP_NONE;
        groups[row * G + r] = (c == DFX_COMP_NONE) ? -1 : (int32_t)(uint32_t)c;
    }
}

// exact canonical fp32 values of the candidates.
//   MODE 0: out[row][c] = comp(value, centroid) for every candidate slot (NONE for unused slots)
//   MODE 1: write the arg-min centroid to assign[row]
//   MODE 2: select the nprobe smallest in shared memory and write keys[row][0..nprobe) (sorted)
// rows: optional indirection (the CTA for list entry b handles row rows[b]); nrows_dev: its length.
template <int MODE>
__global__ void __launch_bounds__(128)
rerank_kernel(const float* __restrict__ Q, int d, const float* __restrict__ cent, const float* __restrict__ cnorm,
              int64_t nlist, int metric, const int32_t* __restrict__ groups, int G, int nprobe,
              const float* __restrict__ gmin, const float* __restrict__ gmin2, const uint8_t* __restrict__ gargc,
              int ng, float cmax2, ScreenTol tol, uint64_t* __restrict__ out, int32_t* __restrict__ assign,
              const int32_t* __restrict__ rows, const int32_t* __restrict__ nrows_dev, int64_t nrows,
              int32_t* __restrict__ keys, int32_t* __restrict__ ovf) {
    DFX_DYN_SMEM(unsigned char, rr_smem, 16);
    float* s_q = reinterpret_cast<float*>(rr_smem);
    uint64_t* s_c = reinterpret_cast<uint64_t*>(rr_smem + ((size_t)d * 4 + 15) / 16 * 16);  // MODE 2
    __shared__ unsigned long long s_best[4];
    __shared__ float s_thr;
    __shared__ int s_cnt;
    const int ncand = G * 32;
    const int64_t limit = nrows_dev ? (int64_t)*nrows_dev : nrows;
    for (int64_t b = blockIdx.x; b < limit; b += gridDim.x) {
        const int64_t row = rows ? rows[b] : b;
        __syncthreads();
        for (int i = threadIdx.x; i < d; i += 128) s_q[i] = Q[row * d + i];
        if (threadIdx.x == 0) s_cnt = 0;
        __syncthreads();
        if (threadIdx.x == 0) {
            float qn2 = 0.f;
            for (int i = 0; i < d; i++) qn2 += s_q[i] * s_q[i];
            // the bound needs nprobe distinct groups: with fewer groups than probes nothing is pruned
            const int gk = (nprobe <= G) ? groups[row * G + nprobe - 1] : -1;
            // (more near-ties than selected groups: the row is also queued for tc_exact_rows_kernel)
            s_thr = screen_threshold(groups, row, G, ng, gmin,
                                     gk >= 0 ? gmin[row * ng + gk] : __int_as_float(0x7f800000), qn2, cmax2, tol, ovf);
        }
        __syncthreads();
        const float thr = s_thr;
        unsigned long long best = DFX_COMP_NONE;
        for (int c = threadIdx.x; c < ncand; c += 128) {
            const int gid = groups[row * G + (c >> 5)];
            unsigned long long comp = DFX_COMP_NONE;
            if (gid >= 0) {
                const int64_t o = row * ng + gid;
                const bool expand = gmin2[o] <= thr;
                const bool live = gmin[o] <= thr && (expand || (c & 31) == (int)gargc[o]);
                const int64_t col = (int64_t)gid * 32 + (c & 31);
                if (live && col < nlist) {
                    const float* x = cent + col * d;
                    float acc = 0.f;
#pragma unroll 8
                    for (int k = 0; k < d; k += 4) {  // (unrolled: 8 row loads in flight, same FMA chain)
                        const float4 xv = *reinterpret_cast<const float4*>(x + k);
                        acc = __fmaf_rn(s_q[k + 0], xv.x, acc);
                        acc = __fmaf_rn(s_q[k + 1], xv.y, acc);
                        acc = __fmaf_rn(s_q[k + 2], xv.z, acc);
                        acc = __fmaf_rn(s_q[k + 3], xv.w, acc);
                    }
                    const float v = (metric == DFX_METRIC_IP) ? -acc : __fmaf_rn(-2.f, acc, cnorm[col]);
                    comp = dfx_comp(v, (uint32_t)col);
                }
            }
            if (MODE == 1) best = comp < best ? comp : best;
            else if (MODE == 0) out[row * ncand + c] = comp;
            else if (comp != DFX_COMP_NONE) s_c[atomicAdd(&s_cnt, 1)] = comp;
        }
        if (MODE == 1) {
#pragma unroll
            for (int off = 16; off >= 1; off >>= 1) {
                unsigned long long o = __shfl_xor_sync(0xffffffffu, best, off);
                best = o < best ? o : best;
            }
            if ((threadIdx.x & 31) == 0) s_best[threadIdx.x >> 5] = best;
            __syncthreads();
            if (threadIdx.x == 0) {
                for (int w = 1; w < 4; w++) best = s_best[w] < best ? s_best[w] : best;
                assign[row] = (best == DFX_COMP_NONE) ? -1 : (int32_t)(uint32_t)best;
            }
        }
        if (MODE == 2) {
            __syncthreads();
            const int cnt = s_cnt;
            int P = 32;
            while (P < cnt) P <<= 1;
            for (int e = cnt + threadIdx.x; e < P; e += 128) s_c[e] = DFX_COMP_NONE;
            dfx_block_bitonic_sort<128>(s_c, P);
            for (int j = threadIdx.x; j < nprobe; j += 128) {
                const uint64_t c = (j < P) ? s_c[j] : DFX_COMP_NONE;
                keys[row * nprobe + j] = (c == DFX_COMP_NONE) ? -1 : (int32_t)(uint32_t)c;
            }
        }
    }
}

// Search path, nprobe <= 32 and G <= 64 (the default there since round 2; measured on B200: the
// decide stage of a 4096-query launch drops from 62 us to ~20 us): the same decision as
// rerank_kernel<2>, one WARP per query instead of one 128-thread CTA.  rerank_kernel<2> costs 62 us per 4096-query launch although only ~9 candidates per query
// survive the screening: one thread sums |q|^2 serially, 128 threads walk 512 candidate slots of
// which a handful are live, then a block-wide sort.  Here: lanes own the selected groups, live
// candidates are compacted into a per-warp list (arg-min column, or all 32 columns of a group
// whose runner-up is within the tolerance), each lane evaluates one candidate in the canonical
// seq-k order, and the nprobe smallest come out of a register bitonic network.  The candidate set
// is a superset of the true top-nprobe for any tolerance >= the screening error, so the keys are
// identical to rerank_kernel<2>'s (|q|^2 is summed in a different order: only the tolerance moves,
// by one ulp).
constexpr int RR2_WARPS = 4;
__global__ void __launch_bounds__(RR2_WARPS * 32)
rerank2_kernel(const float* __restrict__ Q, int d, const float* __restrict__ cent, const float* __restrict__ cnorm,
               int64_t nlist, int metric, const int32_t* __restrict__ groups, int G, int nprobe,
               const float* __restrict__ gmin, const float* __restrict__ gmin2, const uint8_t* __restrict__ gargc,
               int ng, float cmax2, ScreenTol tol, int64_t nrows, int32_t* __restrict__ keys,
               int32_t* __restrict__ ovf) {
    DFX_DYN_SMEM(unsigned char, rr2_smem, 16);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * RR2_WARPS + warp;
    if (row >= nrows) return;  // whole warps leave; no block-wide barrier below
    const int dq = (d + 3) / 4 * 4;
    float* s_q = reinterpret_cast<float*>(rr2_smem) + (size_t)warp * dq;
    int32_t* s_cand = reinterpret_cast<int32_t*>(rr2_smem + (size_t)RR2_WARPS * dq * 4) + (size_t)warp * G * 32;
    float part = 0.f;
#pragma unroll 4
    for (int i = lane; i < d; i += 32) {
        const float v = Q[row * d + i];
        s_q[i] = v;
        part = __fmaf_rn(v, v, part);
    }
    const float qn2 = dfx_warp_butterfly(part);
    __syncwarp();
    const int gk = (nprobe <= G) ? groups[row * G + nprobe - 1] : -1;
    float thr = 0.f;
    if (lane == 0)
        thr = screen_threshold(groups, row, G, ng, gmin, gk >= 0 ? gmin[row * ng + gk] : __int_as_float(0x7f800000),
                               qn2, cmax2, tol, ovf);
    thr = __shfl_sync(0xffffffffu, thr, 0);
    // ---- candidate list: one entry for a live group's arg-min column, 32 for an expanded group
    int cnt = 0;
    for (int g0 = 0; g0 < G; g0 += 32) {
        const int gi = g0 + lane;
        int gid = -1;
        bool live = false, expand = false;
        int argc = 0;
        if (gi < G) {
            gid = groups[row * G + gi];
            if (gid >= 0) {
                const int64_t o = row * ng + gid;
                live = gmin[o] <= thr;
                expand = live && gmin2[o] <= thr;
                argc = (int)gargc[o];
            }
        }
        const bool single = live && !expand && ((int64_t)gid * 32 + argc) < nlist;
        const unsigned ms = __ballot_sync(0xffffffffu, single);
        if (single) s_cand[cnt + __popc(ms & ((1u << lane) - 1u))] = gid * 32 + argc;
        cnt += __popc(ms);
        unsigned me = __ballot_sync(0xffffffffu, expand);
        while (me) {  // rare: a group whose runner-up cannot be ruled out contributes all its columns
            const int src = __ffs((int)me) - 1;
            me &= me - 1;
            const int eg = __shfl_sync(0xffffffffu, gid, src);
            const int64_t col = (int64_t)eg * 32 + lane;
            const bool ok = col < nlist;
            const unsigned mo = __ballot_sync(0xffffffffu, ok);
            if (ok) s_cand[cnt + __popc(mo & ((1u << lane) - 1u))] = (int32_t)col;
            cnt += __popc(mo);
        }
    }
    __syncwarp();
    // ---- exact canonical values, 32 candidates at a time; keep the 32 smallest composites
    uint64_t kept = DFX_COMP_NONE;
    for (int c0 = 0; c0 < cnt; c0 += 32) {
        uint64_t comp = DFX_COMP_NONE;
        if (c0 + lane < cnt) {
            const int64_t col = s_cand[c0 + lane];
            const float* x = cent + col * d;
            float acc = 0.f;
#pragma unroll 8
            for (int k = 0; k < d; k += 4) {  // (unrolled: 8 row loads in flight, same FMA chain)
                const float4 xv = *reinterpret_cast<const float4*>(x + k);
                acc = __fmaf_rn(s_q[k + 0], xv.x, acc);
                acc = __fmaf_rn(s_q[k + 1], xv.y, acc);
                acc = __fmaf_rn(s_q[k + 2], xv.z, acc);
                acc = __fmaf_rn(s_q[k + 3], xv.w, acc);
            }
            const float v = (metric == DFX_METRIC_IP) ? -acc : __fmaf_rn(-2.f, acc, cnorm[col]);
            comp = dfx_comp(v, (uint32_t)col);
        }
        kept = dfx_warp_merge_sorted32(kept, dfx_warp_sort32_asc(comp, lane), lane);
    }
    if (lane < nprobe) keys[row * nprobe + lane] = (kept == DFX_COMP_NONE) ? -1 : (int32_t)(uint32_t)kept;
}

// assign fast path: one thread per row decides from the screening summary alone when the best
// group's arg-min column is unambiguous (no other column within tol); other rows are queued
// for the exact evaluation above.
__global__ void assign_resolve_kernel(const float* __restrict__ qnorm2, const int32_t* __restrict__ groups,
                                      const float* __restrict__ gmin, const float* __restrict__ gmin2,
                                      const uint8_t* __restrict__ gargc, int ng, int G, float cmax2, float tol_rel,
                                      float tol_rel_fast, int64_t n, int32_t* __restrict__ assign,
                                      int32_t* __restrict__ amb_rows, int32_t* __restrict__ amb_count,
                                      int32_t* __restrict__ ovf) {
    const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= n) return;
    const int g0 = groups[row * G + 0], g1 = (G > 1) ? groups[row * G + 1] : -1;
    const int64_t o0 = row * ng + g0;
    const float thr = gmin[o0] + screen_tol(qnorm2[row], cmax2, tol_rel);
    if (tol_rel_fast > 0.f &&
        screen_overflow(groups, row, G, ng, gmin, gmin[o0] + screen_tol(qnorm2[row], cmax2, tol_rel_fast)))
        atomicAdd(&ovf[1], 1);  // AUTO mode statistic: FAST would send this row to the exact fallback
    const bool clear = gmin2[o0] > thr && (g1 < 0 || gmin[row * ng + g1] > thr);
    if (clear) {
        assign[row] = g0 * 32 + (int)gargc[o0];
    } else {
        assign[row] = -1;
        amb_rows[atomicAdd(amb_count, 1)] = (int32_t)row;
    }
}

// ---- exact fallback for the rows the decide stage flagged (screen_overflow): the K smallest
// composites over ALL columns, canonical values computed on the fly (seq-k FMA, the same chain as
// rerank_kernel), selected by the radix select of dfx_select.cuh.  Rare by construction (more than
// `margin` group minima within the tolerance: duplicate centroids / duplicate rows); slow but exact.
struct ExactLoader {
    const float* q;  // the row's query, staged in shared memory by the kernel
    const float* X;
    const float* xnorm;
    int d, metric;   // d % 4 == 0 (the tensor-core path requires d % 64 == 0)
    __device__ __forceinline__ uint64_t operator()(int64_t /*row*/, int e) const {
        const float* x = X + (int64_t)e * d;
        float acc = 0.f;
#pragma unroll 8
        for (int k = 0; k < d; k += 4) {  // 8 row loads in flight, the canonical seq-k FMA chain
            const float4 xv = *reinterpret_cast<const float4*>(x + k);
            acc = __fmaf_rn(q[k + 0], xv.x, acc);
            acc = __fmaf_rn(q[k + 1], xv.y, acc);
            acc = __fmaf_rn(q[k + 2], xv.z, acc);
            acc = __fmaf_rn(q[k + 3], xv.w, acc);
        }
        const float v = (metric == DFX_METRIC_IP) ? -acc : __fmaf_rn(-2.f, acc, xnorm[e]);
        return dfx_comp(v, (uint32_t)e);
    }
};
struct ExactWriter {  // mode 2: keys[row][K]; mode 1: assign[row]; mode 0: out[row][K] composites
    int mode, K;
    int32_t* keys;
    int32_t* assign;
    uint64_t* out;
    __device__ __forceinline__ void operator()(int64_t row, int j, uint64_t c) const {
        const int32_t col = (c == DFX_COMP_NONE) ? -1 : (int32_t)(uint32_t)c;
        if (mode == 2) keys[row * K + j] = col;
        else if (mode == 1) assign[row] = col;
        else out[row * K + j] = c;
    }
};
// dynamic smem: P composites, then the row's query (d floats)
__global__ void __launch_bounds__(256)
tc_exact_rows_kernel(const float* __restrict__ Q, ExactLoader ld, ExactWriter wr, const int32_t* __restrict__ ovf,
                     int n, int k, int P, int sort_cap) {
    DFX_DYN_SMEM(unsigned char, ex_smem, 16);
    uint64_t* s_out = reinterpret_cast<uint64_t*>(ex_smem);
    float* s_q = reinterpret_cast<float*>(ex_smem + (size_t)P * 8);
    ld.q = s_q;
    const int count = ovf[0];
    for (int b = blockIdx.x; b < count; b += gridDim.x) {
        const int64_t row = ovf[2 + b];
        for (int i = threadIdx.x; i < ld.d; i += 256) s_q[i] = Q[row * ld.d + i];
        __syncthreads();
        dfx_select_row<256>(ld, wr, row, n, k, P, sort_cap, s_out);
        __syncthreads();
    }
}
// mode / K as in ExactWriter; Q rows are indexed by the values in ovf (chunk-relative)
static void launch_exact_rows(const float* Q, const float* X, const float* xnorm, int64_t ncols, int d, int metric,
                              const int32_t* ovf, int64_t nrows, int mode, int K, int32_t* keys, int32_t* assign,
                              uint64_t* out, cudaStream_t st) {
    DFX_REQUIRE(K >= 1 && K <= 4096 && ncols < (1ll << 31), "exact fallback: bad K / column count");
    const int sort_cap = 2048, n = (int)ncols;
    const int base = (n <= sort_cap) ? (n > K ? n : K) : K;
    const int P = dfx_next_pow2(base < 2 ? 2 : base);
    ExactLoader ld{nullptr, X, xnorm, d, metric};
    ExactWriter wr{mode, K, keys, assign, out};
    const unsigned grid = (unsigned)std::min<int64_t>(nrows, 2 * 148);
    const size_t smem = (size_t)P * 8 + (size_t)d * 4;
    DFX_CUDA(cudaFuncSetAttribute(tc_exact_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    DFX_LAUNCH(tc_exact_rows_kernel, grid, 256, smem, st, Q, ld, wr, ovf, n, K, P, sort_cap);
}

__global__ void max_reduce_kernel(const float* __restrict__ x, int64_t n, float* __restrict__ out) {
    float m = 0.f;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        m = fmaxf(m, x[i]);
    for (int off = 16; off >= 1; off >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, off));
    if ((threadIdx.x & 31) == 0) atomicMax(reinterpret_cast<int*>(out), __float_as_int(m));  // m >= 0
}

// ------------------------------------------------------------------ host side
#ifndef DFX_EMU
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode_fn() {
    static PFN_encodeTiled fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        DFX_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres));
        DFX_REQUIRE(p != nullptr && qres == cudaDriverEntryPointSuccess, "cuTensorMapEncodeTiled not available");
        fn = (PFN_encodeTiled)p;
    }
    return fn;
}

// fp16 matrix [rows, d] row-major, box = 128 rows x 64 columns, 128-byte swizzle
static void make_tmap(CUtensorMap* tm, const void* base, int64_t rows, int d, int box_rows = tc::TILE) {
    cuuint64_t gdim[2] = {(cuuint64_t)d, (cuuint64_t)rows};
    cuuint64_t gstr[1] = {(cuuint64_t)d * 2};
    cuuint32_t box[2] = {(cuuint32_t)tc::KATOM, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = get_encode_fn()(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), gdim, gstr, box,
                                 estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                 CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    DFX_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with code " + std::to_string((int)r));
}

#endif  // !DFX_EMU

// d = a whole number of 64-element k-atoms (d <= 256: query tile resident; above: streamed)
bool dfx_tc_supported(int d) { return d >= 64 && d % 64 == 0 && d <= 2048; }

// max |c|^2 of a centroid table (for the screening tolerance), synchronises
static float max_norm2(dfx_index* idx, const float* cnorm, int64_t nlist, cudaStream_t st) {
    idx->w_misc.reserve(8);
    DFX_CUDA(cudaMemsetAsync(idx->w_misc.p, 0, 4, st));
    DFX_LAUNCH(max_reduce_kernel, 64, 256, 0, st, cnorm, nlist, idx->w_misc.as<float>());
    float h = 0.f;
    DFX_CUDA(cudaMemcpyAsync(&h, idx->w_misc.p, 4, cudaMemcpyDeviceToHost, st));
    DFX_CUDA(cudaStreamSynchronize(st));
    return h;
}

// fp16 copy of a table (centroids / FLAT rows), npl planes, scaled so that the largest row norm
// lands in [2^13, 2^14); returns the scale
static float tc_prepare_table(const float* x, int64_t n, int d, float cmax2, int npl, DevBuf& out, cudaStream_t st) {
    const int64_t n_pad = dfx_ceil_div(n, tc::TILE) * tc::TILE;
    out.reserve((size_t)npl * n_pad * d * 2);
    const float scale = dfx_pow2_scale(sqrtf(cmax2));
    DFX_LAUNCH(f16_rows_kernel, (unsigned)dfx_ceil_div(n_pad, 8), 256, 0, st, x, n, n_pad, d, 0, scale, 0.f, npl,
               out.as<__half>(), (float*)nullptr, (float*)nullptr);
    return scale;
}

// (re)build the fp16 copy of the centroids (both planes: either precision can use it); call after
// training / import
void dfx_tc_prepare_centroids(dfx_index* idx, cudaStream_t st) {
    const int d = idx->cfg.d;
    if (!dfx_tc_supported(d)) return;
    const int64_t nlist = idx->cfg.nlist;
    idx->tc_cmax2 = max_norm2(idx, idx->cnorm.as<float>(), nlist, st);
    idx->tc_cscale = tc_prepare_table(idx->centroids.as<float>(), nlist, d, idx->tc_cmax2, 2, idx->tc_cent, st);
    idx->tc_cent_npl = 2;
    idx->tc_ready = true;
    if (idx->tc_mode == 0) idx->tc_fast = false;  // AUTO: a new table starts PRECISE
    idx->tc_stat_pending = false;
    idx->tc_acc_rows = idx->tc_acc_bad = 0;
}

// ---- AUTO precision: statistics of a past launch come back through pinned memory, no sync
static bool tc_stream_capturing(cudaStream_t st) {
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(st, &cs) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return cs != cudaStreamCaptureStatusNone;
}
// AUTO rule.  A row that overflows costs an exact pass over ALL columns by one CTA (milliseconds
// at nlist = 65 536), more than FAST saves on a whole launch, so FAST is only worth it where
// overflows practically never happen: PRECISE -> FAST after at least 16 384 observed rows none of
// which would have overflowed under FAST's tolerance; FAST -> PRECISE as soon as more than one row
// in 16 384 did overflow.  Counts accumulate across launches (a batch of one query says nothing
// on its own).  Either way the results are exact; only the cost moves.  A switch bumps
// `generation` (captured graphs are stale).
static void tc_auto_update(dfx_index* idx, bool was_fast, int64_t rows, int64_t overflowed, int64_t fast_would) {
    if (idx->tc_mode != 0 || rows <= 0 || was_fast != idx->tc_fast) return;
    idx->tc_acc_rows += rows;
    idx->tc_acc_bad += was_fast ? overflowed : fast_would;
    bool fast = idx->tc_fast;
    const int64_t window = idx->tc_auto_window;  // 16 384 rows unless dfx_set_param("tc_auto_window") says otherwise
    if (!fast && idx->tc_acc_rows >= window) {
        fast = idx->tc_acc_bad == 0;
        idx->tc_acc_rows = idx->tc_acc_bad = 0;
    } else if (fast && idx->tc_acc_bad > 0 && idx->tc_acc_bad * window > idx->tc_acc_rows) {
        fast = false;
        idx->tc_acc_rows = idx->tc_acc_bad = 0;
    } else if (fast && idx->tc_acc_rows >= (1 << 20)) {
        idx->tc_acc_rows = idx->tc_acc_bad = 0;  // a fresh window
    }
    if (fast != idx->tc_fast) {
        idx->tc_fast = fast;
        idx->generation++;
    }
}
static void tc_stats_poll(dfx_index* idx, cudaStream_t st) {  // before choosing the precision of a launch
    if (!idx->tc_stat_pending || tc_stream_capturing(st)) return;  // (no event queries inside a capture)
    if (cudaEventQuery(idx->tc_stat_ev) != cudaSuccess) {
        cudaGetLastError();  // not ready yet: keep the current precision
        return;
    }
    idx->tc_stat_pending = false;
    idx->tc_last_rows = idx->tc_stat_rows;
    idx->tc_last_overflow = idx->tc_stat_h[0];
    idx->tc_last_fast_would = idx->tc_stat_h[1];
    tc_auto_update(idx, idx->tc_stat_fast, idx->tc_stat_rows, idx->tc_stat_h[0], idx->tc_stat_h[1]);
}
void dfx_tc_stats_sync(dfx_index* idx) {
    if (!idx->tc_stat_pending) return;
    cudaEventSynchronize(idx->tc_stat_ev);
    tc_stats_poll(idx, nullptr);
}
static void tc_stats_post(dfx_index* idx, const int32_t* ovf, int64_t rows, bool fast, cudaStream_t st) {
    if (idx->tc_mode != 0 || idx->tc_stat_pending || tc_stream_capturing(st)) return;
    if (!idx->tc_stat_h) {
        DFX_CUDA(cudaMallocHost(reinterpret_cast<void**>(&idx->tc_stat_h), 8));
        DFX_CUDA(cudaEventCreateWithFlags(&idx->tc_stat_ev, cudaEventDisableTiming));
    }
    DFX_CUDA(cudaMemcpyAsync(idx->tc_stat_h, ovf, 8, cudaMemcpyDeviceToHost, st));
    DFX_CUDA(cudaEventRecord(idx->tc_stat_ev, st));
    idx->tc_stat_pending = true;
    idx->tc_stat_fast = fast;
    idx->tc_stat_rows = rows;
}
static ScreenTol tc_tol(const dfx_index* idx, int d, bool fast) {
    ScreenTol t;
    t.rel = screen_tol_rel(d, fast);
    t.rel_fast = (!fast && idx->tc_mode == 0) ? screen_tol_rel(d, true) : 0.f;
    return t;
}

// screening pass: gmin[nq][ng] for a batch of rows against a table prepared by tc_prepare_table
// (npl planes used; the table must hold at least npl); also leaves |q|^2 in idx->tc_qn
static void tc_screen(dfx_index* idx, int d, const float* d_x, int64_t nq, const void* table_h, float table_scale,
                      int npl, const float* cnorm, int64_t nlist, int metric, float* gmin, float* gmin2,
                      uint8_t* gargc, float* tmin, cudaStream_t st) {
    using namespace tc;
    const int64_t nl_pad = dfx_ceil_div(nlist, TILE) * TILE;
    const int64_t nq_pad = dfx_ceil_div(nq, TILE) * TILE;
    const int ng = (int)(nl_pad / 32);
    idx->tc_q.reserve((size_t)npl * nq_pad * d * 2);
    idx->tc_qmult.reserve((size_t)nq_pad * 4);
    DFX_LAUNCH(f16_rows_kernel, (unsigned)dfx_ceil_div(nq_pad, 8), 256, 0, st, d_x, nq, nq_pad, d, 1, table_scale,
               metric == DFX_METRIC_IP ? -1.f : -2.f, npl, idx->tc_q.as<__half>(), idx->tc_qmult.as<float>(),
               (float*)nullptr);
#ifdef DFX_EMU
    emu_tc_screen(idx->tc_q.as<__half>(), idx->tc_qmult.as<float>(), nq, nq_pad, static_cast<const __half*>(table_h),
                  nlist, nl_pad, d, npl, cnorm, metric, gmin, gmin2, gargc, tmin, ng);
    return;
#else
    const int katoms = d / KATOM;
    // resident only for the k-atom counts the kernel is instantiated for (1, 2, 4); others stream
    const bool res = Smem::resident(katoms, npl) && (katoms == 1 || katoms == 2 || katoms == 4);
    const int tn = Smem::tile_n(res, npl);  // centroids per tile (and per TMA box of the table)
    CUtensorMap tmQ, tmC;
    make_tmap(&tmQ, idx->tc_q.p, npl * nq_pad, d);
    make_tmap(&tmC, table_h, npl * nl_pad, d, tn);
    const int qtiles = (int)(nq_pad / TILE), ctiles = (int)dfx_ceil_div(nl_pad, tn);
    const size_t smem = (size_t)Smem::total(res, katoms, npl) + 1024;
    const int slots = 148 * ((npl == 1 && smem <= 113 * 1024) ? 2 : 1);  // CTAs resident at once
    // split the centroid tiles so that the grid is close to a whole number of waves while every
    // CTA keeps enough tiles to amortise its start (TMEM allocation, query tile, pipeline fill):
    // cost model = waves x (tiles per CTA + ~3 tiles of fixed overhead)
    int csplit = 1, per = ctiles;
    {
        double best_cost = 1e30;
        for (int cs = 1; cs <= ctiles && cs <= 4096; cs++) {
            const int pr = (int)dfx_ceil_div(ctiles, cs);
            const int cs_eff = (int)dfx_ceil_div(ctiles, pr);
            if (cs_eff != cs) continue;
            const int64_t ctas = (int64_t)qtiles * cs;
            const double waves = (double)dfx_ceil_div(ctas, slots);
            const double cost = waves * (pr + 3.0);
            if (cost < best_cost - 1e-9) {
                best_cost = cost;
                csplit = cs;
                per = pr;
            }
        }
    }
    dim3 grid((unsigned)csplit, (unsigned)qtiles);
    DFX_REQUIRE(qtiles <= 65535, "too many query tiles in one screening launch");
#define DFX_TC_LAUNCH(KT_, MT, NP)                                                                       \
    do {                                                                                                 \
        auto kern = tc_coarse_kernel<KT_, MT, NP>;                                                       \
        DFX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));    \
        DFX_LAUNCH(kern, grid, THREADS, smem, st, tmQ, tmC, (int)nq, (int)nq_pad, (int)nlist, (int)nl_pad, katoms, \
                   cnorm, idx->tc_qmult.as<float>(), per, gmin, gmin2, gargc, tmin, ng);                 \
    } while (0)
#define DFX_TC_LAUNCH_M(KT_, NP)                                          \
    do {                                                                  \
        if (metric == DFX_METRIC_L2) DFX_TC_LAUNCH(KT_, DFX_METRIC_L2, NP); \
        else DFX_TC_LAUNCH(KT_, DFX_METRIC_IP, NP);                       \
    } while (0)
    // resident shapes get their k-atom count at compile time (npl * katoms <= 4)
    if (npl == 1) {
        if (res && katoms == 1) DFX_TC_LAUNCH_M(1, 1);
        else if (res && katoms == 2) DFX_TC_LAUNCH_M(2, 1);
        else if (res && katoms == 4) DFX_TC_LAUNCH_M(4, 1);
        else DFX_TC_LAUNCH_M(0, 1);
    } else {
        if (res && katoms == 1) DFX_TC_LAUNCH_M(1, 2);
        else if (res && katoms == 2) DFX_TC_LAUNCH_M(2, 2);
        else DFX_TC_LAUNCH_M(0, 2);
    }
#undef DFX_TC_LAUNCH_M
#undef DFX_TC_LAUNCH
#endif  // DFX_EMU
}

// the G groups of every row with the smallest minima, ascending: two-level warp selection for
// G <= 32 (tile minima first), the generic radix select otherwise
static void tc_select_groups(const float* gmin, const float* tmin, int64_t nrows, int ng, int G, int32_t* groups,
                             cudaStream_t st) {
    if (G <= 32 && ng >= 128) {
        auto tk = topg_collect_kernel<256>;
        DFX_LAUNCH(tk, (unsigned)dfx_ceil_div(nrows, 8), 256, 0, st, tmin, nrows, ng / 4, G, gmin, ng, groups);
    } else if (G <= 32 && ng >= 32) {
        auto tk = topg_collect_kernel<256>;
        DFX_LAUNCH(tk, (unsigned)dfx_ceil_div(nrows, 8), 256, 0, st, gmin, nrows, ng, G, (const float*)nullptr, 0,
                   groups);
    } else {
        dfx_launch_select_cols(gmin, nrows, ng, ng, G, 0, groups, nullptr, nullptr, 0, st);
    }
}

// the overflow record of a decide launch: [overflow rows, FAST-would-overflow rows, row list ...]
static int32_t* tc_ovf_reset(dfx_index* idx, int64_t nrows, cudaStream_t st) {
    idx->tc_ovf.reserve((size_t)(nrows + 2) * 4);
    DFX_CUDA(cudaMemsetAsync(idx->tc_ovf.p, 0, 8, st));
    return idx->tc_ovf.as<int32_t>();
}

// top-nprobe lists per query -> keys int32 [nq, nprobe] (exactly the oracle's coarse result)
void dfx_tc_coarse_search(dfx_index* idx, const float* d_x, int64_t nq, int nprobe, int32_t* keys, cudaStream_t st) {
    const int d = idx->cfg.d;
    const int64_t nlist = idx->cfg.nlist;
    const int metric = idx->cfg.metric;
    const int64_t nl_pad = dfx_ceil_div(nlist, tc::TILE) * tc::TILE;
    const int ng = (int)(nl_pad / 32);
    // groups kept per query: nprobe + 8; with few groups in all (nlist <= 2048) every one is kept,
    // so that no row can overflow (at d = 768, nprobe 32 of 64 groups, most rows did)
    int G = nprobe + 8;
    if (G > ng || ng <= 64) G = ng;
    const int64_t QC = std::max<int64_t>(tc::TILE, ((64ll << 20) / ((int64_t)ng * 4)) / tc::TILE * tc::TILE);
    const int64_t qmax = std::min<int64_t>(nq, QC);
    idx->tc_gmin.reserve((size_t)qmax * ng * 4);
    idx->tc_gmin2.reserve((size_t)qmax * ng * 4);
    idx->tc_gargc.reserve((size_t)qmax * ng);
    idx->tc_groups.reserve((size_t)qmax * G * 4);
    idx->tc_tmin.reserve((size_t)qmax * (ng / 4 + 1) * 4);
    float* tmin = idx->tc_tmin.as<float>();
    const float* cent = idx->centroids.as<float>();
    const float* cnorm = idx->cnorm.as<float>();
    float* gmin = idx->tc_gmin.as<float>();
    float* gmin2 = idx->tc_gmin2.as<float>();
    uint8_t* gargc = idx->tc_gargc.as<uint8_t>();
    int32_t* groups = idx->tc_groups.as<int32_t>();
    tc_stats_poll(idx, st);
    const bool fast = idx->tc_fast;
    const ScreenTol tol = tc_tol(idx, d, fast);
    for (int64_t q0 = 0; q0 < nq; q0 += QC) {
        const int64_t qc = std::min(QC, nq - q0);
        const float* xq = d_x + q0 * d;
        int32_t* kq = keys + q0 * nprobe;
        tc_screen(idx, d, xq, qc, idx->tc_cent.p, idx->tc_cscale, fast ? 1 : 2, cnorm, nlist, metric, gmin, gmin2,
                  gargc, tmin, st);
        tc_select_groups(gmin, tmin, qc, ng, G, groups, st);
        int32_t* ovf = tc_ovf_reset(idx, qc, st);
        const int P_cand = dfx_next_pow2(G * 32 < 32 ? 32 : G * 32);
        if (nprobe <= 32 && G <= 64 && d % 4 == 0) {  // warp per query
            const size_t smem = (size_t)RR2_WARPS * ((d + 3) / 4 * 4) * 4 + (size_t)RR2_WARPS * G * 32 * 4;
            if (smem > 48 * 1024)
                DFX_CUDA(cudaFuncSetAttribute(rerank2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            DFX_LAUNCH(rerank2_kernel, (unsigned)dfx_ceil_div(qc, RR2_WARPS), RR2_WARPS * 32, smem, st, xq, d, cent,
                       cnorm, nlist, metric, groups, G, nprobe, gmin, gmin2, gargc, ng, idx->tc_cmax2, tol, qc, kq, ovf);
        } else if (P_cand <= 4096) {  // fused: candidates never leave the SM
            auto kern = rerank_kernel<2>;
            const size_t smem = ((size_t)d * 4 + 15) / 16 * 16 + (size_t)P_cand * 8;
            DFX_LAUNCH(kern, (unsigned)qc, 128, smem, st, xq, d, cent, cnorm, nlist, metric, groups, G, nprobe, gmin,
                       gmin2, gargc, ng, idx->tc_cmax2, tol, (uint64_t*)nullptr, (int32_t*)nullptr,
                       (const int32_t*)nullptr, (const int32_t*)nullptr, qc, kq, ovf);
        } else {
            idx->tc_cand.reserve((size_t)qmax * G * 32 * 8);
            auto kern = rerank_kernel<0>;
            DFX_LAUNCH(kern, (unsigned)qc, 128, (size_t)d * 4, st, xq, d, cent, cnorm, nlist, metric, groups, G,
                       nprobe, gmin, gmin2, gargc, ng, idx->tc_cmax2, tol, idx->tc_cand.as<uint64_t>(),
                       (int32_t*)nullptr, (const int32_t*)nullptr, (const int32_t*)nullptr, qc, (int32_t*)nullptr, ovf);
            dfx_launch_select_comp(idx->tc_cand.as<uint64_t>(), qc, G * 32, (int64_t)G * 32, nprobe, kq, st);
        }
        // rows with more near-ties than selected groups (screen_overflow): exact, over all lists
        launch_exact_rows(xq, cent, cnorm, nlist, d, metric, ovf, qc, 2, nprobe, kq, nullptr, nullptr, st);
        if (q0 + QC >= nq) tc_stats_post(idx, ovf, qc, fast, st);
    }
}

// Flat search on tensor cores (default for d in {64, 128}, N >= 1024, k <= 100; dfx_set_param
// "flat_tensor_cores" = 0 selects the FFMA GEMM; measured on B200, config C1 100 k x 1 k queries:
// 13 -> 161 TFLOP/s algorithmic, bit-exact): flat search (reference
// index.py:94 IndexFlatIP, and IndexFlatL2 of the C-ABI) through the same machinery, with the
// database rows in place of the centroids and k in place of nprobe.  Screening on tensor cores
// over bf16 planes of the rows (built lazily, rebuilt after an add), the k + 8 groups with the
// smallest minima, exact canonical values of the columns the screening cannot rule out
// (rerank_kernel<0>) -- the values gemm_values_kernel would have produced for them, bit for bit.
// Leaves nq x ncand composites in idx->tc_cand; the caller selects and writes (D, I).
// Returns ncand, or 0 when this shape is not handled (caller falls back to the FFMA GEMM).
int dfx_tc_flat_candidates(dfx_index* idx, const float* d_x, int64_t nq, int k, cudaStream_t st) {
    const int d = idx->cfg.d;
    const int64_t N = idx->n_sorted;
    if (!dfx_tc_supported(d) || N < 1024 || k > 100) return 0;
    const int metric = idx->cfg.metric;
    const int64_t nl_pad = dfx_ceil_div(N, tc::TILE) * tc::TILE;
    const int ng = (int)(nl_pad / 32);
    if (idx->tc_flat_rows != N) {  // (re)build the fp16 copy of the rows and the norm bound
        const float* norms = idx->xnorm.as<float>();
        if (metric != DFX_METRIC_L2) {  // inner product keeps no row norms: the tolerance needs their maximum
            idx->cnorm.reserve((size_t)N * 4);
            dfx_launch_row_norms(idx->payload.as<float>(), N, d, idx->cnorm.as<float>(), st);
            norms = idx->cnorm.as<float>();
        }
        idx->tc_cmax2 = max_norm2(idx, norms, N, st);
        idx->tc_cscale = tc_prepare_table(idx->payload.as<float>(), N, d, idx->tc_cmax2, 2, idx->tc_cent, st);
        idx->tc_cent_npl = 2;
        idx->tc_flat_rows = N;
        if (idx->tc_mode == 0) idx->tc_fast = false;  // AUTO: new rows start PRECISE
        idx->tc_stat_pending = false;
        idx->tc_acc_rows = idx->tc_acc_bad = 0;
    }
    tc_stats_poll(idx, st);
    const bool fast = idx->tc_fast;
    const ScreenTol tol = tc_tol(idx, d, fast);
    int G = k + 8;
    if (G > ng) G = ng;
    const int ncand = G * 32;
    idx->tc_gmin.reserve((size_t)nq * ng * 4);
    idx->tc_gmin2.reserve((size_t)nq * ng * 4);
    idx->tc_gargc.reserve((size_t)nq * ng);
    idx->tc_groups.reserve((size_t)nq * G * 4);
    idx->tc_cand.reserve((size_t)nq * ncand * 8);
    idx->tc_tmin.reserve((size_t)nq * (ng / 4 + 1) * 4);
    tc_screen(idx, d, d_x, nq, idx->tc_cent.p, idx->tc_cscale, fast ? 1 : 2, idx->xnorm.as<float>(), N, metric,
              idx->tc_gmin.as<float>(), idx->tc_gmin2.as<float>(), idx->tc_gargc.as<uint8_t>(),
              idx->tc_tmin.as<float>(), st);
    tc_select_groups(idx->tc_gmin.as<float>(), idx->tc_tmin.as<float>(), nq, ng, G, idx->tc_groups.as<int32_t>(), st);
    int32_t* ovf = tc_ovf_reset(idx, nq, st);
    auto kern = rerank_kernel<0>;
    DFX_LAUNCH(kern, (unsigned)nq, 128, (size_t)d * 4, st, d_x, d, idx->payload.as<float>(), idx->xnorm.as<float>(), N,
               metric, idx->tc_groups.as<int32_t>(), G, k, idx->tc_gmin.as<float>(), idx->tc_gmin2.as<float>(),
               idx->tc_gargc.as<uint8_t>(), ng, idx->tc_cmax2, tol, idx->tc_cand.as<uint64_t>(), (int32_t*)nullptr,
               (const int32_t*)nullptr, (const int32_t*)nullptr, nq, (int32_t*)nullptr, ovf);
    // rows with more near-ties than selected groups: their ncand slots become the exact top-ncand
    launch_exact_rows(d_x, idx->payload.as<float>(), idx->xnorm.as<float>(), N, d, metric, ovf, nq, 0, ncand, nullptr,
                      nullptr, idx->tc_cand.as<uint64_t>(), st);
    tc_stats_post(idx, ovf, nq, fast, st);
    return ncand;
}

// nearest centroid per row (build path): screening, then the answer straight from the screening
// summary when it is unambiguous, exact canonical evaluation for the (rare) ambiguous rows
void dfx_tc_assign(dfx_index* idx, int d, const float* d_cent, const float* d_cnorm, int64_t nlist, int metric,
                   int64_t n, const float* d_x, int32_t* d_assign, cudaStream_t st) {
    const int64_t nl_pad = dfx_ceil_div(nlist, tc::TILE) * tc::TILE;
    const int ng = (int)(nl_pad / 32);
    // groups kept per row: the best one decides, the others tell whether it does so unambiguously;
    // a row whose G-th group is still within the tolerance goes to the exact fallback
    const int G = ng < 6 ? ng : 6;
    const int64_t RC = std::max<int64_t>(tc::TILE, ((256ll << 20) / ((int64_t)ng * 4)) / tc::TILE * tc::TILE);
    const int64_t rmax = std::min<int64_t>(n, RC);
    idx->tc_gmin.reserve((size_t)rmax * ng * 4);
    idx->tc_gmin2.reserve((size_t)rmax * ng * 4);
    idx->tc_gargc.reserve((size_t)rmax * ng);
    idx->tc_groups.reserve((size_t)rmax * G * 4);
    idx->tc_tmin.reserve((size_t)rmax * (ng / 4 + 1) * 4);
    idx->tc_qn.reserve((size_t)rmax * 4);
    idx->tc_amb.reserve((size_t)(rmax + 1) * 4);
    // fp16 copy of this centroid table (k-means changes it every iteration; converting is cheap).
    // Precision: tc_mode 1 / 2 as set; AUTO decides per chunk from the counts of the chunk before
    // (read back with a sync -- this is the build path), starting PRECISE on every call.
    const float cmax2 = max_norm2(idx, d_cnorm, nlist, st);
    const float cscale = tc_prepare_table(d_cent, nlist, d, cmax2, 2, idx->tc_cent_tmp, st);
    bool fast = idx->tc_mode == 1;
    int64_t acc_rows = 0, acc_bad = 0;
    for (int64_t r0 = 0, rc = 0; r0 < n; r0 += rc) {
        // the first AUTO chunk is a small PRECISE probe (16 384 rows), so that most of the rows run
        // at the precision their statistics call for
        rc = (idx->tc_mode == 0 && r0 == 0) ? std::min<int64_t>(n, dfx_ceil_div(idx->tc_auto_window, tc::TILE) * tc::TILE)
                                            : std::min(RC, n - r0);
        const float* xr = d_x + r0 * d;
        ScreenTol tol;
        tol.rel = screen_tol_rel(d, fast);
        tol.rel_fast = 0.f;  // counted once per row by assign_resolve_kernel
        const float rel_fast = (!fast && idx->tc_mode == 0) ? screen_tol_rel(d, true) : 0.f;
        tc_screen(idx, d, xr, rc, idx->tc_cent_tmp.p, cscale, fast ? 1 : 2, d_cnorm, nlist, metric,
                  idx->tc_gmin.as<float>(), idx->tc_gmin2.as<float>(), idx->tc_gargc.as<uint8_t>(),
                  idx->tc_tmin.as<float>(), st);
        tc_select_groups(idx->tc_gmin.as<float>(), idx->tc_tmin.as<float>(), rc, ng, G, idx->tc_groups.as<int32_t>(),
                         st);
        dfx_launch_row_norms(xr, rc, d, idx->tc_qn.as<float>(), st);
        int32_t* amb_count = idx->tc_amb.as<int32_t>();
        int32_t* amb_rows = amb_count + 1;
        DFX_CUDA(cudaMemsetAsync(amb_count, 0, 4, st));
        int32_t* ovf = tc_ovf_reset(idx, rc, st);
        DFX_LAUNCH(assign_resolve_kernel, (unsigned)dfx_ceil_div(rc, 256), 256, 0, st, idx->tc_qn.as<float>(),
                   idx->tc_groups.as<int32_t>(), idx->tc_gmin.as<float>(), idx->tc_gmin2.as<float>(),
                   idx->tc_gargc.as<uint8_t>(), ng, G, cmax2, tol.rel, rel_fast, rc, d_assign + r0, amb_rows,
                   amb_count, ovf);
        auto kern = rerank_kernel<1>;
        const unsigned grid = (unsigned)std::min<int64_t>(rc, 148 * 16);
        DFX_LAUNCH(kern, grid, 128, (size_t)d * 4, st, xr, d, d_cent, d_cnorm, nlist, metric,
                   idx->tc_groups.as<int32_t>(), G, 1, idx->tc_gmin.as<float>(), idx->tc_gmin2.as<float>(),
                   idx->tc_gargc.as<uint8_t>(), ng, cmax2, tol, (uint64_t*)nullptr, d_assign + r0,
                   (const int32_t*)amb_rows, (const int32_t*)amb_count, rc, (int32_t*)nullptr, ovf);
        launch_exact_rows(xr, d_cent, d_cnorm, nlist, d, metric, ovf, rc, 1, 1, nullptr, d_assign + r0, nullptr, st);
        if (idx->tc_mode == 0 && r0 + rc < n) {  // AUTO: the next chunk's precision (rule of tc_auto_update)
            int32_t h[2] = {0, 0};
            DFX_CUDA(cudaMemcpyAsync(h, ovf, 8, cudaMemcpyDeviceToHost, st));
            DFX_CUDA(cudaStreamSynchronize(st));
            acc_rows += rc;
            acc_bad += fast ? h[0] : h[1];
            if (!fast && acc_rows >= idx->tc_auto_window) {
                fast = acc_bad == 0;
                acc_rows = acc_bad = 0;
            } else if (fast && acc_bad > 0 && acc_bad * idx->tc_auto_window > acc_rows) {
                fast = false;
                acc_rows = acc_bad = 0;
            }
        }
    }
}
