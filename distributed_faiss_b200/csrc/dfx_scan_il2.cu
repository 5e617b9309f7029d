// dfx_scan_il2.cu -- launcher of the experimental lane-per-vector IVF-PQ scan (K4 v3).
// Kernel: dfx_scan_il2_dev.cuh.
#include "dfx_scan_il2_dev.cuh"
#include <cstdlib>

template <bool REG, int THREADS, bool RING = false>
static void launch_il2(dfx_index* idx, int64_t qc, const int32_t* keys, int nprobe, int G, int ngroups, int k, int cap,
                       uint64_t* part, cudaStream_t st) {
    constexpr int NW = THREADS / 32;
    const size_t topk = REG ? (size_t)NW * IL2_QCAP * 8 : (size_t)NW * cap * 8;
    const size_t smem = (size_t)IL2_LUT_BYTES + (RING ? (topk + 15) / 16 * 16 + (size_t)NW * IL2_RING * IL2_SLOT_BYTES : topk);
    auto kern = scan_pq_il2_kernel<REG, THREADS, RING>;
    DFX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    DFX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    DFX_LAUNCH(kern, (unsigned)(qc * ngroups), THREADS, smem, st, idx->w_lut.as<float>(), idx->w_dis0.as<float>(), keys,
               nprobe, G, ngroups, idx->blk_off.as<int64_t>(), idx->il_codes.as<uint4>(), idx->il_tvals.as<float>(),
               idx->il_ids.as<int32_t>(), k, cap, part);
}

void dfx_launch_scan_pq_il2(dfx_index* idx, int64_t qc, const int32_t* keys, int nprobe, int G, int ngroups, int k,
                            int cap, uint64_t* part, cudaStream_t st) {
    DFX_REQUIRE(G <= IL2_MAXG, "scan_pq_il2: more than 16 probes per CTA");
    static const int threads = [] {  // experiments: DFX_IL2_THREADS=384 -> 2 CTAs/SM of 12 warps
        const char* e = getenv("DFX_IL2_THREADS");
        return (e && atoi(e) == 384) ? 384 : 256;
    }();
    const bool reg = k <= 32;
    if (idx->il2_ring && reg) {  // experimental: code blocks through per-warp cp.async.bulk rings
        launch_il2<true, 256, true>(idx, qc, keys, nprobe, G, ngroups, k, cap, part, st);
        return;
    }
    // (the shared-memory top-k path sorts NW * cap entries with a bitonic network: NW must be a
    // power of two, so 12-warp CTAs are only used with the register top-k)
    if (threads == 384 && reg) {
        launch_il2<true, 384>(idx, qc, keys, nprobe, G, ngroups, k, cap, part, st);
    } else {
        if (reg) launch_il2<true, 256>(idx, qc, keys, nprobe, G, ngroups, k, cap, part, st);
        else launch_il2<false, 256>(idx, qc, keys, nprobe, G, ngroups, k, cap, part, st);
    }
}
