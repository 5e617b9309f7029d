// dfx_tc.cu -- K1: the coarse quantizer on 5th-generation tensor cores (tcgen05 + TMEM + TMA).
//
// Replaces the dense query x centroid contraction inside `quantizer.search(nq, x, nprobe)`
// (faiss IndexIVF::search, reached from reference distributed_faiss/index.py:257) and
// `quantizer.assign` (IndexIVF::add, index.py:425).
//
// Scheme ("screen on tensor cores, decide in canonical fp32"):
//   1. fp32 operands are split into bf16 hi + lo planes; q.c ~= qh.ch + ql.ch + qh.cl is
//      accumulated in fp32 in TMEM by tcgen05.mma (relative error ~1e-6, the ql.cl term
//      ~2^-18 is dropped).
//   2. the epilogue turns each 128x128 accumulator tile into ranking values
//      (L2: |c|^2 - 2 q.c, IP: -q.c) and keeps only the MINIMUM of every group of 32
//      consecutive centroids -> gmin[nq][nlist/32]  (32x less traffic than the full matrix).
//   3. the G = nprobe + margin groups with the smallest minima are selected exactly
//      (dfx_select.cuh).  Every true top-nprobe centroid lies in one of the nprobe groups
//      with the smallest group minimum; `margin` absorbs the 1e-6 screening error.
//   4. the G*32 candidate centroids are re-evaluated in the CANONICAL fp32 order
//      (seq-k FMA, identical to oracle/dfx_oracle.c) and the final top-nprobe / argmin is
//      taken on those exact values -> the probe lists are bit-identical to the oracle's.
//
// One CTA = 6 warps: warp 0 TMA producer, warp 1 TMEM allocator + single-thread MMA issuer,
// warps 2..5 epilogue (one TMEM lane == one query row per thread).  A (query planes) stays
// resident in shared memory, B (centroid planes) streams through a ring of 16 KB stages, two
// TMEM accumulator buffers overlap the epilogue of tile i with the MMAs of tile i+1.
#include "dfx_internal.h"
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
__device__ __forceinline__ void tc_mma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
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

// instruction descriptor: D=f32, A=B=bf16, both K-major, M=128, N=128
static constexpr uint32_t TC_IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((128u >> 3) << 17) | ((128u >> 4) << 24);

#endif  // !DFX_EMU

// ------------------------------------------------------------------ the kernel
namespace tc {
constexpr int TILE = 128;           // rows of A and of B per tile
constexpr int KATOM = 64;           // bf16 elements per 128-byte swizzle row
constexpr int ATOM_BYTES = TILE * KATOM * 2;  // 16 KB
constexpr int NSTAGE = 8;
constexpr int THREADS = 192;        // TMA warp + MMA warp + 4 epilogue warps
constexpr int TMEM_COLS = 256;      // two 128-column accumulator buffers
struct Smem {
    // offsets inside the 1024-aligned dynamic shared memory block
    static constexpr int A = 0;                                   // [plane hi/lo][katom] x 16 KB
    static constexpr int B(int katoms) { return 2 * katoms * ATOM_BYTES; }
    static constexpr int BARS(int katoms) { return B(katoms) + NSTAGE * ATOM_BYTES; }
    static constexpr int total(int katoms) { return BARS(katoms) + 512 + 2 * TILE * 4; }
};
}  // namespace tc

#ifdef DFX_EMU
// ---- CPU emulator stand-ins (tests/emu/): the screening kernel cannot be emulated instruction
// by instruction, so its RESULT is restated: the same bf16 hi/lo split of both operands, the
// same three partial products, fp32 accumulation (order not specified by the hardware either),
// and the epilogue's packed (min, runner-up, arg-min) per 32 centroids.  Everything downstream
// (group selection, exact canonical re-evaluation, the drivers) is the product code.
struct __nv_bfloat16 { unsigned short x; };
static inline float emu_bf16_rn(float v) {  // round to nearest even bf16, returned as fp32
    uint32_t u;
    memcpy(&u, &v, 4);
    if ((u & 0x7f800000u) == 0x7f800000u) return v;
    u += 0x7fffu + ((u >> 16) & 1u);
    u &= 0xffff0000u;
    float r;
    memcpy(&r, &u, 4);
    return r;
}
// "planes" hold a padded fp32 copy in the emulator build (same byte size as the two bf16 planes)
__global__ void split_bf16_kernel(const float* __restrict__ x, int64_t n, int64_t n_pad, int d,
                                  __nv_bfloat16* __restrict__ out) {
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_pad * d) return;
    reinterpret_cast<float*>(out)[t] = (t / d < n) ? x[t] : 0.f;
}
static void emu_tc_screen(const float* q, int64_t nq, const float* c, int64_t nlist, int64_t nl_pad, int d,
                          const float* cnorm, int metric, float* gmin, float* gmin2, uint8_t* gargc, int ng) {
    const float big = 3.0e38f;
    std::vector<float> ch((size_t)nl_pad * d), cl((size_t)nl_pad * d), qh(d), ql(d);
    for (size_t i = 0; i < ch.size(); i++) {
        ch[i] = emu_bf16_rn(c[i]);
        cl[i] = emu_bf16_rn(c[i] - ch[i]);
    }
    for (int64_t row = 0; row < nq; row++) {
        for (int k = 0; k < d; k++) {
            qh[k] = emu_bf16_rn(q[row * d + k]);
            ql[k] = emu_bf16_rn(q[row * d + k] - qh[k]);
        }
        for (int g = 0; g < ng; g++) {
            float m1 = big, m2 = big;
            for (int j = 0; j < 32; j++) {
                const int64_t col = (int64_t)g * 32 + j;
                float ip = 0.f;
                const float* h = &ch[(size_t)col * d];
                const float* l = &cl[(size_t)col * d];
                for (int k = 0; k < d; k++) ip += qh[k] * h[k] + ql[k] * h[k] + qh[k] * l[k];
                float cn = big;
                if (col < nlist) cn = (metric == DFX_METRIC_L2) ? cnorm[col] : 0.f;
                const float v = (metric == DFX_METRIC_IP) ? (cn - ip) : fmaf(-2.f, ip, cn);
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
    }
}
#else
// tmQ: bf16 [2*nq_pad, d]  (rows [0,nq_pad) = hi plane, [nq_pad, 2 nq_pad) = lo plane)
// tmC: bf16 [2*nl_pad, d]
// gmin: float [nq][ng], ng = nl_pad/32
template <int KATOMS, int METRIC>
__global__ void __launch_bounds__(tc::THREADS, 1)
tc_coarse_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmC, int nq,
                 int nq_pad, int nlist, int nl_pad, const float* __restrict__ cnorm, int metric, int ctiles_per_cta,
                 float* __restrict__ gmin, float* __restrict__ gmin2, uint8_t* __restrict__ gargc, int ng) {
    using namespace tc;
    extern __shared__ unsigned char smem_raw_tc[];
    unsigned char* smem = reinterpret_cast<unsigned char*>(
        (reinterpret_cast<uintptr_t>(smem_raw_tc) + 1023) & ~(uintptr_t)1023);  // SWIZZLE_128B atoms
    unsigned char* sA = smem + Smem::A;
    unsigned char* sB = smem + Smem::B(KATOMS);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Smem::BARS(KATOMS));
    uint64_t* full = bars;                 // [NSTAGE]
    uint64_t* empty = bars + NSTAGE;       // [NSTAGE]
    uint64_t* a_full = bars + 2 * NSTAGE;  // [1]
    uint64_t* t_full = a_full + 1;         // [2]
    uint64_t* t_empty = t_full + 2;        // [2]
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(t_empty + 2);
    float* s_cn = reinterpret_cast<float*>(smem + Smem::BARS(KATOMS) + 512);  // [2][TILE]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int qt = blockIdx.y;
    const int ctiles = nl_pad / TILE;
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

    constexpr int STAGES_PER_TILE = 2 * KATOMS;  // ch atoms then cl atoms

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            mbar_expect_tx(a_full, 2 * KATOMS * ATOM_BYTES);
            for (int pl = 0; pl < 2; pl++)
                for (int ka = 0; ka < KATOMS; ka++)
                    tma_load_2d(sA + (pl * KATOMS + ka) * ATOM_BYTES, &tmQ, a_full, ka * KATOM,
                                pl * nq_pad + qt * TILE);
            int stage = 0, phase = 0;
            for (int t = 0; t < ntiles; t++) {
                const int crow = (ct0 + t) * TILE;
                for (int s = 0; s < STAGES_PER_TILE; s++) {
                    const int pl = s / KATOMS, ka = s % KATOMS;
                    mbar_wait(&empty[stage], phase ^ 1);
                    mbar_expect_tx(&full[stage], ATOM_BYTES);
                    tma_load_2d(sB + stage * ATOM_BYTES, &tmC, &full[stage], ka * KATOM, pl * nl_pad + crow);
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
            mbar_wait(a_full, 0);
            tc_fence_after();
            int stage = 0, phase = 0;
            const uint32_t a_base = smem_u32(sA);
            const uint32_t b_base = smem_u32(sB);
            for (int t = 0; t < ntiles; t++) {
                const int buf = t & 1;
                mbar_wait(&t_empty[buf], ((t >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + buf * TILE;
                for (int s = 0; s < STAGES_PER_TILE; s++) {
                    const int pl = s / KATOMS, ka = s % KATOMS;
                    mbar_wait(&full[stage], phase);
                    tc_fence_after();
                    const uint32_t b_addr = b_base + stage * ATOM_BYTES;
                    const uint32_t ah_addr = a_base + (0 * KATOMS + ka) * ATOM_BYTES;
                    const uint32_t al_addr = a_base + (1 * KATOMS + ka) * ATOM_BYTES;
#pragma unroll
                    for (int kk = 0; kk < KATOM / 16; kk++) {  // UMMA_K = 16 bf16 = 32 bytes
                        const uint64_t bd = make_kmajor_sw128_desc(b_addr + kk * 32);
                        // qh . (ch | cl)
                        tc_mma_bf16(d_tmem, make_kmajor_sw128_desc(ah_addr + kk * 32), bd, TC_IDESC,
                                    (s > 0 || kk > 0) ? 1u : 0u);
                        // ql . ch
                        if (pl == 0) tc_mma_bf16(d_tmem, make_kmajor_sw128_desc(al_addr + kk * 32), bd, TC_IDESC, 1u);
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
        // The epilogue is bound by the half-rate ALU pipe, so it is kept to 3 FMNMX + 1 LOP3 per
        // element: the column index (0..31) replaces the 5 low mantissa bits of the screening
        // value (a 2^-18 relative perturbation, far below the screening tolerance), which makes
        // the arg-min fall out of the minimum itself; the runner-up is min(m2, max(m1, v)).
        const int quad = warp & 3;             // TMEM lane quadrant this warp may access
        const int row = quad * 32 + lane;      // query row inside the tile == TMEM lane
        const int64_t grow = (int64_t)qt * TILE + row;
        const int et = threadIdx.x - 64;       // 0..127
        const float big = 3.0e38f;             // out-of-range columns: finite, never selected
        for (int t = 0; t < ntiles; t++) {
            const int buf = t & 1;
            const int col0 = (ct0 + t) * TILE;
            {
                const int c = col0 + et;
                float cn = 0.f;
                if (METRIC == DFX_METRIC_L2 && c < nlist) cn = cnorm[c];
                s_cn[buf * TILE + et] = (c < nlist) ? cn : big;
            }
            asm volatile("bar.sync 1, 128;" ::: "memory");
            mbar_wait(&t_full[buf], (t >> 1) & 1);
            tc_fence_after();
            float gm[4], gm2[4];
            uint32_t ga = 0;
#pragma unroll
            for (int ch = 0; ch < 4; ch++) {
                uint32_t r[32];
                tc_ld32(tmem_base + buf * TILE + ch * 32 + ((uint32_t)(quad * 32) << 16), r);
                float m1 = big, m2 = big;
#pragma unroll
                for (int j = 0; j < 32; j++) {
                    const float ip = __uint_as_float(r[j]);
                    const float cn = s_cn[buf * TILE + ch * 32 + j];
                    const float v = (METRIC == DFX_METRIC_IP) ? (cn - ip) : fmaf(-2.f, ip, cn);
                    const float vj = __uint_as_float((__float_as_uint(v) & ~31u) | (uint32_t)j);
                    m2 = fminf(m2, fmaxf(m1, vj));
                    m1 = fminf(m1, vj);
                }
                gm[ch] = m1;
                gm2[ch] = m2;
                ga |= (__float_as_uint(m1) & 31u) << (8 * ch);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&t_empty[buf]);
            if (grow < nq) {
                const int64_t o = grow * ng + (col0 >> 5);
                *reinterpret_cast<float4*>(gmin + o) = make_float4(gm[0], gm[1], gm[2], gm[3]);
                *reinterpret_cast<float4*>(gmin2 + o) = make_float4(gm2[0], gm2[1], gm2[2], gm2[3]);
                *reinterpret_cast<uint32_t*>(gargc + o) = ga;
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

// ------------------------------------------------------------------ helpers around it
// fp32 [n, d] -> bf16 hi/lo planes [2][n_pad][d] (rows >= n zero filled)
__global__ void split_bf16_kernel(const float* __restrict__ x, int64_t n, int64_t n_pad, int d,
                                  __nv_bfloat16* __restrict__ out) {
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_pad * d) return;
    int64_t i = t / d;
    float v = (i < n) ? x[t] : 0.f;
    __nv_bfloat16 hi = __float2bfloat16_rn(v);
    __nv_bfloat16 lo = __float2bfloat16_rn(v - __bfloat162float(hi));
    out[t] = hi;
    out[n_pad * d + t] = lo;
}

#endif  // DFX_EMU

// the G smallest group minima of a row, G <= 8: one warp per row, G rounds of warp arg-min
template <int G>
__global__ void topg_small_kernel(const float* __restrict__ gmin, int64_t nq, int ng,
                                  int32_t* __restrict__ groups) {
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= nq) return;
    const float* g = gmin + row * ng;
    uint64_t taken[G];
#pragma unroll
    for (int r = 0; r < G; r++) taken[r] = DFX_COMP_NONE;
#pragma unroll
    for (int r = 0; r < G; r++) {
        uint64_t best = DFX_COMP_NONE;
        const uint64_t prev = (r == 0) ? 0 : taken[r - 1];
        for (int j = lane; j < ng; j += 32) {
            uint64_t c = dfx_comp(g[j], (uint32_t)j);
            // composites are unique: the r-th smallest is the smallest one above the (r-1)-th
            if ((r == 0 || c > prev) && c < best) best = c;
        }
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) {
            uint64_t o = __shfl_xor_sync(0xffffffffu, best, off);
            best = o < best ? o : best;
        }
        taken[r] = best;
        if (lane == 0) groups[row * G + r] = (best == DFX_COMP_NONE) ? -1 : (int32_t)(uint32_t)best;
    }
}

// Which columns of the selected groups can still matter?  With t = value of the nprobe-th
// smallest group minimum and tol >= twice the screening error, every true top-nprobe centroid c
// has approx(c) <= t + tol; inside its group it is either the arg-min column or has
// approx(c) >= gmin2(group).  So a group is expanded to all 32 columns only if
// gmin2 <= t + tol, otherwise its arg-min column is the only candidate.
__device__ __forceinline__ float screen_tol(float qn2, float cmax2) {
    // bf16-split error ~1e-5 |q||c|, plus the 2^-18 |v| perturbation of the packed column index
    // (|v| <= |c|^2 + 2|q||c|); both with a wide margin
    return 1e-4f * (sqrtf(qn2 * cmax2) + cmax2) + 1e-30f;
}

// The G (<= 32) smallest group minima of a row, ascending by (value, group): one warp per row.
//   1. every lane takes the minimum of its ng/32 values; the G-th smallest of those 32 lane
//      minima, B, bounds the answer from above (G different lanes hold a value <= B);
//   2. the values <= B are collected (a few dozen at most in practice) into a per-warp buffer;
//   3. the buffer is bitonic-sorted by the warp and the first G entries are written.
// Exact for any input; if more than CAP values are <= B (massive ties) the row falls back to
// G rounds of warp arg-min.  ~10x fewer instructions than G x ng compare rounds.
template <int CAP>
__global__ void __launch_bounds__(256)
topg_collect_kernel(const float* __restrict__ gmin, int64_t nq, int ng, int G, int32_t* __restrict__ groups) {
    __shared__ uint64_t s_buf[8][CAP];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * 8 + warp;
    if (row >= nq) return;
    const float* g = gmin + row * ng;
    uint64_t* buf = s_buf[warp];
    // 1. lane minima -> bound
    uint64_t lmin = DFX_COMP_NONE;
    for (int j = lane; j < ng; j += 32) {
        const uint64_t c = dfx_comp(g[j], (uint32_t)j);
        lmin = c < lmin ? c : lmin;
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
    for (int j0 = 0; j0 < ng; j0 += 32) {
        const int j = j0 + lane;
        const uint64_t c = (j < ng) ? dfx_comp(g[j], (uint32_t)j) : DFX_COMP_NONE;
        const bool want = c <= bound && c != DFX_COMP_NONE;
        const unsigned mask = __ballot_sync(0xffffffffu, want);
        if (mask) {
            const int pos = cnt + __popc(mask & ((1u << lane) - 1u));
            if (want && pos < CAP) buf[pos] = c;
            cnt += __popc(mask);
            if (cnt > CAP) overflow = true;
        }
    }
    if (!overflow) {
        // 3. sort the candidates
        int P = 32;
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
        for (int r = lane; r < G; r += 32) {
            const uint64_t c = (r < P) ? buf[r] : DFX_COMP_NONE;
            groups[row * G + r] = (c == DFX_COMP_NONE) ? -1 : (int32_t)(uint32_t)c;
        }
        return;
    }
    // fallback: G rounds of "smallest composite above the previous pick"
    uint64_t prev = 0;
    for (int r = 0; r < G; r++) {
        uint64_t best = DFX_COMP_NONE;
        for (int j = lane; j < ng; j += 32) {
            const uint64_t c = dfx_comp(g[j], (uint32_t)j);
            if ((r == 0 || c > prev) && c < best) best = c;
        }
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) {
            const uint64_t o = __shfl_xor_sync(0xffffffffu, best, off);
            best = o < best ? o : best;
        }
        if (lane == 0) groups[row * G + r] = (best == DFX_COMP_NONE) ? -1 : (int32_t)(uint32_t)best;
        prev = best;
        if (best == DFX_COMP_NONE) prev = DFX_COMP_NONE - 1;
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
              int ng, float cmax2, uint64_t* __restrict__ out, int32_t* __restrict__ assign,
              const int32_t* __restrict__ rows, const int32_t* __restrict__ nrows_dev, int64_t nrows,
              int32_t* __restrict__ keys) {
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
            s_thr = (gk >= 0 ? gmin[row * ng + gk] : __int_as_float(0x7f800000)) + screen_tol(qn2, cmax2);
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
                    for (int k = 0; k < d; k += 4) {
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
               int ng, float cmax2, int64_t nrows, int32_t* __restrict__ keys) {
    DFX_DYN_SMEM(unsigned char, rr2_smem, 16);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * RR2_WARPS + warp;
    if (row >= nrows) return;  // whole warps leave; no block-wide barrier below
    const int dq = (d + 3) / 4 * 4;
    float* s_q = reinterpret_cast<float*>(rr2_smem) + (size_t)warp * dq;
    int32_t* s_cand = reinterpret_cast<int32_t*>(rr2_smem + (size_t)RR2_WARPS * dq * 4) + (size_t)warp * G * 32;
    float part = 0.f;
    for (int i = lane; i < d; i += 32) {
        const float v = Q[row * d + i];
        s_q[i] = v;
        part = __fmaf_rn(v, v, part);
    }
    const float qn2 = dfx_warp_butterfly(part);
    __syncwarp();
    const int gk = (nprobe <= G) ? groups[row * G + nprobe - 1] : -1;
    const float thr = (gk >= 0 ? gmin[row * ng + gk] : __int_as_float(0x7f800000)) + screen_tol(qn2, cmax2);
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
            for (int k = 0; k < d; k += 4) {
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
                                      const uint8_t* __restrict__ gargc, int ng, float cmax2, int64_t n,
                                      int32_t* __restrict__ assign, int32_t* __restrict__ amb_rows,
                                      int32_t* __restrict__ amb_count) {
    const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= n) return;
    const int g0 = groups[row * 2 + 0], g1 = groups[row * 2 + 1];
    const int64_t o0 = row * ng + g0;
    const float thr = gmin[o0] + screen_tol(qnorm2[row], cmax2);
    const bool clear = gmin2[o0] > thr && (g1 < 0 || gmin[row * ng + g1] > thr);
    if (clear) {
        assign[row] = g0 * 32 + (int)gargc[o0];
    } else {
        assign[row] = -1;
        amb_rows[atomicAdd(amb_count, 1)] = (int32_t)row;
    }
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

// bf16 matrix [rows, d] row-major, box = 128 rows x 64 columns, 128-byte swizzle
static void make_tmap(CUtensorMap* tm, const void* base, int64_t rows, int d) {
    cuuint64_t gdim[2] = {(cuuint64_t)d, (cuuint64_t)rows};
    cuuint64_t gstr[1] = {(cuuint64_t)d * 2};
    cuuint32_t box[2] = {(cuuint32_t)tc::KATOM, (cuuint32_t)tc::TILE};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = get_encode_fn()(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstr, box,
                                 estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                 CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    DFX_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with code " + std::to_string((int)r));
}

#endif  // !DFX_EMU

bool dfx_tc_supported(int d) { return d == 64 || d == 128; }

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

// (re)build the bf16 planes of the centroids; call after training / import
void dfx_tc_prepare_centroids(dfx_index* idx, cudaStream_t st) {
    const int d = idx->cfg.d;
    if (!dfx_tc_supported(d)) return;
    const int64_t nlist = idx->cfg.nlist;
    const int64_t nl_pad = dfx_ceil_div(nlist, tc::TILE) * tc::TILE;
    idx->tc_cent.reserve((size_t)2 * nl_pad * d * 2);
    DFX_LAUNCH(split_bf16_kernel, (unsigned)dfx_ceil_div(nl_pad * d, 256), 256, 0, st, idx->centroids.as<float>(),
               nlist, nl_pad, d, idx->tc_cent.as<__nv_bfloat16>());
    idx->tc_cmax2 = max_norm2(idx, idx->cnorm.as<float>(), nlist, st);
    idx->tc_ready = true;
}

// screening pass: gmin[nq][ng] for a batch of rows against any centroid table with bf16 planes
static void tc_screen(dfx_index* idx, int d, const float* d_x, int64_t nq, const void* cent_planes,
                      const float* cnorm, int64_t nlist, int metric, float* gmin, float* gmin2, uint8_t* gargc,
                      cudaStream_t st) {
    using namespace tc;
    const int64_t nl_pad = dfx_ceil_div(nlist, TILE) * TILE;
    const int64_t nq_pad = dfx_ceil_div(nq, TILE) * TILE;
    const int ng = (int)(nl_pad / 32);
    idx->tc_q.reserve((size_t)2 * nq_pad * d * 2);
    DFX_LAUNCH(split_bf16_kernel, (unsigned)dfx_ceil_div(nq_pad * d, 256), 256, 0, st, d_x, nq, nq_pad, d,
               idx->tc_q.as<__nv_bfloat16>());
#ifdef DFX_EMU
    emu_tc_screen(idx->tc_q.as<float>(), nq, static_cast<const float*>(cent_planes), nlist, nl_pad, d, cnorm, metric,
                  gmin, gmin2, gargc, ng);
    return;
#else
    CUtensorMap tmQ, tmC;
    make_tmap(&tmQ, idx->tc_q.p, 2 * nq_pad, d);
    make_tmap(&tmC, cent_planes, 2 * nl_pad, d);
    const int qtiles = (int)(nq_pad / TILE), ctiles = (int)(nl_pad / TILE);
    // enough CTAs to fill the machine a few times over, each walking a contiguous range of
    // centroid tiles with its query tile resident
    // split the centroid tiles so that the grid is close to a whole number of 148-CTA waves
    // (one CTA per SM) while every CTA keeps enough tiles to amortise loading its query tile:
    // cost model = waves x (tiles per CTA + ~3 tiles of fixed overhead)
    int csplit = 1, per = ctiles;
    {
        double best_cost = 1e30;
        for (int cs = 1; cs <= ctiles && cs <= 4096; cs++) {
            const int pr = (int)dfx_ceil_div(ctiles, cs);
            const int cs_eff = (int)dfx_ceil_div(ctiles, pr);
            if (cs_eff != cs) continue;
            const int64_t ctas = (int64_t)qtiles * cs;
            const double waves = (double)dfx_ceil_div(ctas, 148);
            const double cost = waves * (pr + 3.0);
            if (cost < best_cost - 1e-9) {
                best_cost = cost;
                csplit = cs;
                per = pr;
            }
        }
    }
    const int katoms = d / KATOM;
    const size_t smem = (size_t)Smem::total(katoms) + 1024;
    dim3 grid((unsigned)csplit, (unsigned)qtiles);
    DFX_REQUIRE(qtiles <= 65535, "too many query tiles in one screening launch");
#define DFX_TC_LAUNCH(KA, MT)                                                                            \
    do {                                                                                                 \
        auto kern = tc_coarse_kernel<KA, MT>;                                                            \
        DFX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));    \
        DFX_LAUNCH(kern, grid, THREADS, smem, st, tmQ, tmC, (int)nq, (int)nq_pad, (int)nlist, (int)nl_pad, \
                   cnorm, metric, per, gmin, gmin2, gargc, ng);                                          \
    } while (0)
    if (katoms == 2 && metric == DFX_METRIC_L2) DFX_TC_LAUNCH(2, DFX_METRIC_L2);
    else if (katoms == 2) DFX_TC_LAUNCH(2, DFX_METRIC_IP);
    else if (metric == DFX_METRIC_L2) DFX_TC_LAUNCH(1, DFX_METRIC_L2);
    else DFX_TC_LAUNCH(1, DFX_METRIC_IP);
#undef DFX_TC_LAUNCH
#endif  // DFX_EMU
}

// top-nprobe lists per query -> keys int32 [nq, nprobe] (exactly the oracle's coarse result)
void dfx_tc_coarse_search(dfx_index* idx, const float* d_x, int64_t nq, int nprobe, int32_t* keys, cudaStream_t st) {
    const int d = idx->cfg.d;
    const int64_t nlist = idx->cfg.nlist;
    const int64_t nl_pad = dfx_ceil_div(nlist, tc::TILE) * tc::TILE;
    const int ng = (int)(nl_pad / 32);
    int G = nprobe + 8;
    if (G > ng) G = ng;
    const int64_t QC = std::max<int64_t>(tc::TILE, ((64ll << 20) / ((int64_t)ng * 4)) / tc::TILE * tc::TILE);
    const int64_t qmax = std::min<int64_t>(nq, QC);
    idx->tc_gmin.reserve((size_t)qmax * ng * 4);
    idx->tc_gmin2.reserve((size_t)qmax * ng * 4);
    idx->tc_gargc.reserve((size_t)qmax * ng);
    idx->tc_groups.reserve((size_t)qmax * G * 4);
    for (int64_t q0 = 0; q0 < nq; q0 += QC) {
        const int64_t qc = std::min(QC, nq - q0);
        tc_screen(idx, d, d_x + q0 * d, qc, idx->tc_cent.p, idx->cnorm.as<float>(), nlist, idx->cfg.metric,
                  idx->tc_gmin.as<float>(), idx->tc_gmin2.as<float>(), idx->tc_gargc.as<uint8_t>(), st);
        if (G <= 32 && ng >= 32) {
            auto tk = topg_collect_kernel<256>;
            DFX_LAUNCH(tk, (unsigned)dfx_ceil_div(qc, 8), 256, 0, st, idx->tc_gmin.as<float>(), qc, ng, G,
                       idx->tc_groups.as<int32_t>());
        } else {
            dfx_launch_select_cols(idx->tc_gmin.as<float>(), qc, ng, ng, G, 0, idx->tc_groups.as<int32_t>(),
                                   nullptr, nullptr, 0, st);
        }
        const int P_cand = dfx_next_pow2(G * 32 < 32 ? 32 : G * 32);
        if (nprobe <= 32 && G <= 64 && d % 4 == 0) {  // warp per query
            const size_t smem = (size_t)RR2_WARPS * ((d + 3) / 4 * 4) * 4 + (size_t)RR2_WARPS * G * 32 * 4;
            DFX_LAUNCH(rerank2_kernel, (unsigned)dfx_ceil_div(qc, RR2_WARPS), RR2_WARPS * 32, smem, st, d_x + q0 * d, d,
                       idx->centroids.as<float>(), idx->cnorm.as<float>(), nlist, idx->cfg.metric,
                       idx->tc_groups.as<int32_t>(), G, nprobe, idx->tc_gmin.as<float>(), idx->tc_gmin2.as<float>(),
                       idx->tc_gargc.as<uint8_t>(), ng, idx->tc_cmax2, qc, keys + q0 * nprobe);
        } else if (P_cand <= 4096) {  // fused: candidates never leave the SM
            auto kern = rerank_kernel<2>;
            const size_t smem = ((size_t)d * 4 + 15) / 16 * 16 + (size_t)P_cand * 8;
            DFX_LAUNCH(kern, (unsigned)qc, 128, smem, st, d_x + q0 * d, d, idx->centroids.as<float>(),
                       idx->cnorm.as<float>(), nlist, idx->cfg.metric, idx->tc_groups.as<int32_t>(), G, nprobe,
                       idx->tc_gmin.as<float>(), idx->tc_gmin2.as<float>(), idx->tc_gargc.as<uint8_t>(), ng,
                       idx->tc_cmax2, (uint64_t*)nullptr, (int32_t*)nullptr, (const int32_t*)nullptr,
                       (const int32_t*)nullptr, qc, keys + q0 * nprobe);
        } else {
            idx->tc_cand.reserve((size_t)qmax * G * 32 * 8);
            auto kern = rerank_kernel<0>;
            DFX_LAUNCH(kern, (unsigned)qc, 128, (size_t)d * 4, st, d_x + q0 * d, d, idx->centroids.as<float>(),
                       idx->cnorm.as<float>(), nlist, idx->cfg.metric, idx->tc_groups.as<int32_t>(), G, nprobe,
                       idx->tc_gmin.as<float>(), idx->tc_gmin2.as<float>(), idx->tc_gargc.as<uint8_t>(), ng,
                       idx->tc_cmax2, idx->tc_cand.as<uint64_t>(), (int32_t*)nullptr, (const int32_t*)nullptr,
                       (const int32_t*)nullptr, qc, (int32_t*)nullptr);
            dfx_launch_select_comp(idx->tc_cand.as<uint64_t>(), qc, G * 32, (int64_t)G * 32, nprobe,
                                   keys + q0 * nprobe, st);
        }
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
    if (idx->tc_flat_rows != N) {  // (re)build the planes and the norm bound
        idx->tc_cent.reserve((size_t)2 * nl_pad * d * 2);
        DFX_LAUNCH(split_bf16_kernel, (unsigned)dfx_ceil_div(nl_pad * d, 256), 256, 0, st, idx->payload.as<float>(), N,
                   nl_pad, d, idx->tc_cent.as<__nv_bfloat16>());
        const float* norms = idx->xnorm.as<float>();
        if (metric != DFX_METRIC_L2) {  // inner product keeps no row norms: the tolerance needs their maximum
            idx->cnorm.reserve((size_t)N * 4);
            dfx_launch_row_norms(idx->payload.as<float>(), N, d, idx->cnorm.as<float>(), st);
            norms = idx->cnorm.as<float>();
        }
        idx->tc_cmax2 = max_norm2(idx, norms, N, st);
        idx->tc_flat_rows = N;
    }
    int G = k + 8;
    if (G > ng) G = ng;
    const int ncand = G * 32;
    idx->tc_gmin.reserve((size_t)nq * ng * 4);
    idx->tc_gmin2.reserve((size_t)nq * ng * 4);
    idx->tc_gargc.reserve((size_t)nq * ng);
    idx->tc_groups.reserve((size_t)nq * G * 4);
    idx->tc_cand.reserve((size_t)nq * ncand * 8);
    tc_screen(idx, d, d_x, nq, idx->tc_cent.p, idx->xnorm.as<float>(), N, metric, idx->tc_gmin.as<float>(),
              idx->tc_gmin2.as<float>(), idx->tc_gargc.as<uint8_t>(), st);
    if (G <= 32 && ng >= 32) {
        auto tk = topg_collect_kernel<256>;
        DFX_LAUNCH(tk, (unsigned)dfx_ceil_div(nq, 8), 256, 0, st, idx->tc_gmin.as<float>(), nq, ng, G,
                   idx->tc_groups.as<int32_t>());
    } else {
        dfx_launch_select_cols(idx->tc_gmin.as<float>(), nq, ng, ng, G, 0, idx->tc_groups.as<int32_t>(), nullptr,
                               nullptr, 0, st);
    }
    auto kern = rerank_kernel<0>;
    DFX_LAUNCH(kern, (unsigned)nq, 128, (size_t)d * 4, st, d_x, d, idx->payload.as<float>(), idx->xnorm.as<float>(), N,
               metric, idx->tc_groups.as<int32_t>(), G, k, idx->tc_gmin.as<float>(), idx->tc_gmin2.as<float>(),
               idx->tc_gargc.as<uint8_t>(), ng, idx->tc_cmax2, idx->tc_cand.as<uint64_t>(), (int32_t*)nullptr,
               (const int32_t*)nullptr, (const int32_t*)nullptr, nq, (int32_t*)nullptr);
    return ncand;
}

// nearest centroid per row (build path): screening, then the answer straight from the screening
// summary when it is unambiguous, exact canonical evaluation for the (rare) ambiguous rows
void dfx_tc_assign(dfx_index* idx, int d, const float* d_cent, const float* d_cnorm, int64_t nlist, int metric,
                   int64_t n, const float* d_x, int32_t* d_assign, cudaStream_t st) {
    const int64_t nl_pad = dfx_ceil_div(nlist, tc::TILE) * tc::TILE;
    const int ng = (int)(nl_pad / 32);
    constexpr int G = 2;
    const int64_t RC = std::max<int64_t>(tc::TILE, ((256ll << 20) / ((int64_t)ng * 4)) / tc::TILE * tc::TILE);
    const int64_t rmax = std::min<int64_t>(n, RC);
    idx->tc_gmin.reserve((size_t)rmax * ng * 4);
    idx->tc_gmin2.reserve((size_t)rmax * ng * 4);
    idx->tc_gargc.reserve((size_t)rmax * ng);
    idx->tc_groups.reserve((size_t)rmax * G * 4);
    idx->tc_qn.reserve((size_t)rmax * 4);
    idx->tc_amb.reserve((size_t)(rmax + 1) * 4);
    // bf16 planes of this centroid table (k-means changes it every iteration; splitting is cheap)
    idx->tc_cent_tmp.reserve((size_t)2 * nl_pad * d * 2);
    DFX_LAUNCH(split_bf16_kernel, (unsigned)dfx_ceil_div(nl_pad * d, 256), 256, 0, st, d_cent, nlist, nl_pad, d,
               idx->tc_cent_tmp.as<__nv_bfloat16>());
    const void* cent_planes = idx->tc_cent_tmp.p;
    const float cmax2 = max_norm2(idx, d_cnorm, nlist, st);
    for (int64_t r0 = 0; r0 < n; r0 += RC) {
        const int64_t rc = std::min(RC, n - r0);
        const float* xr = d_x + r0 * d;
        tc_screen(idx, d, xr, rc, cent_planes, d_cnorm, nlist, metric, idx->tc_gmin.as<float>(),
                  idx->tc_gmin2.as<float>(), idx->tc_gargc.as<uint8_t>(), st);
        auto topg = topg_small_kernel<G>;
        DFX_LAUNCH(topg, (unsigned)dfx_ceil_div(rc, 8), 256, 0, st, idx->tc_gmin.as<float>(), rc, ng,
                   idx->tc_groups.as<int32_t>());
        dfx_launch_row_norms(xr, rc, d, idx->tc_qn.as<float>(), st);
        int32_t* amb_count = idx->tc_amb.as<int32_t>();
        int32_t* amb_rows = amb_count + 1;
        DFX_CUDA(cudaMemsetAsync(amb_count, 0, 4, st));
        DFX_LAUNCH(assign_resolve_kernel, (unsigned)dfx_ceil_div(rc, 256), 256, 0, st, idx->tc_qn.as<float>(),
                   idx->tc_groups.as<int32_t>(), idx->tc_gmin.as<float>(), idx->tc_gmin2.as<float>(),
                   idx->tc_gargc.as<uint8_t>(), ng, cmax2, rc, d_assign + r0, amb_rows, amb_count);
        auto kern = rerank_kernel<1>;
        const unsigned grid = (unsigned)std::min<int64_t>(rc, 148 * 16);
        DFX_LAUNCH(kern, grid, 128, (size_t)d * 4, st, xr, d, d_cent, d_cnorm, nlist, metric,
                   idx->tc_groups.as<int32_t>(), G, 1, idx->tc_gmin.as<float>(), idx->tc_gmin2.as<float>(),
                   idx->tc_gargc.as<uint8_t>(), ng, cmax2, (uint64_t*)nullptr, d_assign + r0,
                   (const int32_t*)amb_rows, (const int32_t*)amb_count, rc, (int32_t*)nullptr);
    }
}
