// emu_pq.cpp -- CPU build of the IVF-PQ block-layout pipeline's DEVICE code: layout conversion,
// K3 pq_prep, K4 v2 (scan_pq_il_kernel) and K4 v3 (scan_pq_il2_kernel), compiled by g++ from the
// same headers nvcc compiles (distributed_faiss_b200/csrc/*_dev.cuh) and run on the fiber SIMT
// runtime (simt.h).  Lets the transcription of a kernel be checked against the oracle without a
// GPU (tests/test_emu_kernels.py).  Test infrastructure only; nothing here is shipped.
#define SIMT_IMPLEMENTATION
#include <cuda_runtime.h>  // the shim in tests/emu/shim

std::atomic<long long> g_dfx_launches{0};
void dfx_set_error(const std::string&) {}

#include "dfx_scan_il_dev.cuh"
#include "dfx_scan_il2_dev.cuh"
#include "dfx_pq_prep_dev.cuh"

extern "C" {

void emu_seed(uint64_t seed) { simt::S().rng = seed; }
static uint64_t g_seed = 0;
void emu_set_seed(uint64_t seed) { g_seed = seed; }

int emu_rm_to_il(int layout, int64_t nlist, const int64_t* list_off, const int64_t* blk_off, const uint8_t* codes,
                 const float* tvals, const int32_t* ids, int64_t nblk, uint8_t* il_codes, float* il_tvals,
                 int32_t* il_ids) {
    simt::launch((unsigned)nblk, 256, 0, [=] {
        pq_rm_to_il_kernel(list_off, blk_off, nlist, codes, tvals, ids, il_codes, il_tvals, il_ids, layout);
    }, g_seed);
    return 0;
}

int emu_il_to_rm(int layout, int64_t nlist, const int64_t* list_off, const int64_t* blk_off, const uint8_t* il_codes,
                 const float* il_tvals, const int32_t* il_ids, int64_t nblk, uint8_t* codes, float* tvals,
                 int32_t* ids) {
    simt::launch((unsigned)nblk, 256, 0, [=] {
        pq_il_to_rm_kernel(list_off, blk_off, nlist, il_codes, il_tvals, il_ids, codes, tvals, ids, layout);
    }, g_seed);
    return 0;
}

// transposed: 0 lut[q][m][j], 1 lut[q][j][m], 2 lut[q][j][64]
int emu_pq_prep(const float* Q, int64_t nq, int d, int M, int ksub, int dsub, const float* codebooks,
                const float* cent, const int32_t* keys, int nprobe, float* lut, float* dis0, int transposed) {
    const size_t smem = (size_t)((d + 3) / 4) * 16 + (transposed ? (size_t)M * (ksub + 1) * 4 : 0);
    simt::launch((unsigned)nq, 256, smem, [=] {
        pq_prep_kernel(Q, d, M, ksub, dsub, codebooks, cent, keys, nprobe, lut, dis0, transposed);
    }, g_seed);
    return 0;
}

// experimental K3 variant 2: transposed codebook + QB = 8 queries per CTA
int emu_pq_prep2(const float* Q, int64_t nq, int d, const float* codebooks, const float* cent, const int32_t* keys,
                 int nprobe, float* lut, float* dis0, int wide) {
    std::vector<float> cbT((size_t)32 * 256 * 4);
    float* pT = cbT.data();
    simt::launch((32 * 256 * 4 + 255) / 256, 256, 0, [=] { cb_transpose_kernel(codebooks, 32, 256, 4, pT); }, g_seed);
    constexpr int QB = 8;
    simt::launch((unsigned)((nq + QB - 1) / QB), 256, (size_t)QB * d * 4, [=] {
        pq_prep2_kernel<QB>(Q, nq, d, pT, cent, keys, nprobe, lut, dis0, wide);
    }, g_seed);
    return 0;
}

int emu_scan_v2(const float* lutT, const float* dis0, const int32_t* keys, int64_t nq, int nprobe, int G, int ngroups,
                const int64_t* blk_off, const uint8_t* il_codes, const float* il_tvals, const int32_t* il_ids, int k,
                int cap, uint64_t* part, int split) {
    const size_t smem = (size_t)256 * 32 * 4 + (size_t)(IL_THREADS / 32) * cap * 8;
    const uint4* c4 = reinterpret_cast<const uint4*>(il_codes);
    if (split)  // block layout 3
        simt::launch((unsigned)(nq * ngroups), IL_THREADS, smem, [=] {
            scan_pq_il_split_kernel(lutT, dis0, keys, nprobe, G, ngroups, blk_off, c4, il_tvals, il_ids, k, cap, part);
        }, g_seed);
    else
        simt::launch((unsigned)(nq * ngroups), IL_THREADS, smem, [=] {
            scan_pq_il_kernel(lutT, dis0, keys, nprobe, G, ngroups, blk_off, c4, il_tvals, il_ids, k, cap, part);
        }, g_seed);
    return 0;
}

int emu_scan_v3(const float* lutW, const float* dis0, const int32_t* keys, int64_t nq, int nprobe, int G, int ngroups,
                const int64_t* blk_off, const uint8_t* il_codes, const float* il_tvals, const int32_t* il_ids, int k,
                int cap, uint64_t* part) {
    if (G > IL2_MAXG) return 1;
    const bool reg = k <= 32;
    const size_t smem = (size_t)IL2_LUT_BYTES + (reg ? (size_t)(IL2_THREADS / 32) * IL2_QCAP * 8 : (size_t)(IL2_THREADS / 32) * cap * 8);
    const uint4* c4 = reinterpret_cast<const uint4*>(il_codes);
    if (reg)
        simt::launch((unsigned)(nq * ngroups), IL2_THREADS, smem, [=] {
            scan_pq_il2_kernel<true>(lutW, dis0, keys, nprobe, G, ngroups, blk_off, c4, il_tvals, il_ids, k, cap, part);
        }, g_seed);
    else
        simt::launch((unsigned)(nq * ngroups), IL2_THREADS, smem, [=] {
            scan_pq_il2_kernel<false>(lutW, dis0, keys, nprobe, G, ngroups, blk_off, c4, il_tvals, il_ids, k, cap, part);
        }, g_seed);
    return 0;
}

}  // extern "C"
