// dfx_scan_il2.cu -- launcher of the experimental lane-per-vector IVF-PQ scan (K4 v3).
// Kernel: dfx_scan_il2_dev.cuh.
#include "dfx_scan_il2_dev.cuh"

void dfx_launch_scan_pq_il2(dfx_index* idx, int64_t qc, const int32_t* keys, int nprobe, int G, int ngroups, int k,
                            int cap, uint64_t* part, cudaStream_t st) {
    DFX_REQUIRE(G <= IL2_MAXG, "scan_pq_il2: more than 16 probes per CTA");
    const bool reg = k <= 32;
    const size_t smem = (size_t)IL2_LUT_BYTES +
                        (reg ? (size_t)IL2_NW * IL2_QCAP * 8 : (size_t)IL2_NW * cap * 8);
    if (reg) {
        auto kern = scan_pq_il2_kernel<true>;
        DFX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        DFX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100));  // 3 CTAs/SM
        DFX_LAUNCH(kern, (unsigned)(qc * ngroups), IL2_THREADS, smem, st, idx->w_lut.as<float>(),
                   idx->w_dis0.as<float>(), keys, nprobe, G, ngroups, idx->blk_off.as<int64_t>(),
                   idx->il_codes.as<uint4>(), idx->il_tvals.as<float>(), idx->il_ids.as<int32_t>(), k, cap, part);
    } else {
        auto kern = scan_pq_il2_kernel<false>;
        DFX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        DFX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100));  // 3 CTAs/SM
        DFX_LAUNCH(kern, (unsigned)(qc * ngroups), IL2_THREADS, smem, st, idx->w_lut.as<float>(),
                   idx->w_dis0.as<float>(), keys, nprobe, G, ngroups, idx->blk_off.as<int64_t>(),
                   idx->il_codes.as<uint4>(), idx->il_tvals.as<float>(), idx->il_ids.as<int32_t>(), k, cap, part);
    }
}
