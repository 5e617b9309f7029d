// dfx_internal.h -- index object + launcher declarations shared by the .cu files.
#pragma once
#include "dfx_common.cuh"
#include <vector>
#include <mutex>
#include <memory>
#include <algorithm>

struct dfx_index {
    dfx_cfg cfg{};
    int M = 0, ksub = 0, dsub = 0;
    int64_t nprobe = 1;  // faiss default (index.py:47 copies it back into cfg for knnlm)
    bool trained = false;

    // list-sorted storage ("sorted") and arrival-order staging ("pending")
    int64_t n_sorted = 0, n_pending = 0;
    DevBuf centroids, cnorm, codebooks;
    DevBuf list_off;                    // int64[nlist+1]
    std::vector<int64_t> h_list_off;    // host mirror
    DevBuf ids;                         // int32[n_sorted] shard-local ids, list-sorted
    DevBuf payload;                     // kind-specific rows, list-sorted (FLAT: arrival order)
    DevBuf tvals;                       // IVF_PQ: f32[n_sorted]
    DevBuf xnorm;                       // FLAT + L2: f32[n]
    DevBuf p_list, p_payload, p_tvals;  // pending: int32 list per row, payload rows, tvals
    int64_t p_cap = 0;
    int64_t reserve_hint = 0;

    // search workspace (grow-only)
    DevBuf w_vals, w_keys, w_dis0, w_lut, w_part, w_q, w_D, w_I, w_misc, w_best;
    // reconstruct support: inverse of ids (shard-local id -> storage position)
    DevBuf inv;
    bool inv_valid = false;
    // training knobs (dfx_set_param)
    int kmeans_niter = 25;            // faiss Clustering default niter
    int max_points_per_centroid = 256;  // faiss Clustering default
    uint64_t train_seed = 1234;       // faiss Clustering default seed

    // IVF-PQ, M == 32: interleaved block layout for the lane-per-subquantizer scan
    // (dfx_scan_il.cu).  While `il` is set the row-major payload/tvals/ids are released.
    bool il = false, il_enabled = true;
    DevBuf il_codes, il_tvals, il_ids, blk_off;
    int64_t nblk = 0;
    // which block layout / scan kernel: 1 = dfx_il_byte + scan_pq_il_kernel (default),
    // 2 = dfx_il2_byte + scan_pq_il2_kernel (dfx_scan_il2.cu; dfx_set_param "scan_variant"),
    // 3 = dfx_il3_byte + scan_pq_il_split_kernel (the default kernel on coalesced halves)
    // K3 variant (dfx_set_param "prep_variant"): 2 = pq_prep2_kernel on the transposed codebook
    int prep_variant = 1;
    DevBuf codebooksT;  // [ksub][M][dsub], built on demand
    bool cbT_valid = false;
    int il_variant = 1;  // requested
    int il_layout = 0;   // layout the il_* arrays currently hold (valid while `il`)

    // tensor-core coarse quantizer (dfx_tc.cu): bf16 hi/lo planes and screening workspace
    DevBuf tc_cent, tc_cent_tmp, tc_q, tc_gmin, tc_gmin2, tc_gargc, tc_groups, tc_cand, tc_qn, tc_amb;
    float tc_cmax2 = 0.f;
    bool tc_ready = false;
    bool tc_enabled = true;
    int rerank_variant = 1;  // 2 = rerank2_kernel (warp per query; dfx_set_param "rerank_variant")
    bool il2_ring = false;      // scan 2: feed the code blocks through shared-memory rings (dfx_set_param "scan_ring")
    int rows_inflight = 4;      // vectors in flight per warp of scan_rows_kernel (8: experimental)
    bool flat_tc = false;       // FLAT: search through the tensor-core screening (dfx_tc_flat_candidates)
    int64_t tc_flat_rows = -1;  // rows covered by the bf16 planes of a FLAT index (-1: none)

    // scan-kernel profiling (dfx_profile_enable)
    bool prof_on = false;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> prof_events;
    size_t prof_used = 0;

    // last-search bookkeeping
    int64_t last_nq = 0, last_nprobe = 0;
    bool last_keys_valid = false;

    cudaStream_t stream = nullptr;  // used by the host-pointer entry points
    // stream of the most recent *_dev call; host-pointer entry points (which run on `stream`,
    // a non-blocking stream) wait for it first so that the two never race
    cudaStream_t last_dev_stream = nullptr;
    bool dev_work_pending = false;
    void note_dev(cudaStream_t st) {
        last_dev_stream = st;
        dev_work_pending = true;
    }
    // a *_dev call on a different stream than the previous one: order them
    void join_dev_if_other(cudaStream_t st) {
        if (dev_work_pending && st != last_dev_stream) join_dev();
    }
    void join_dev() {
        if (dev_work_pending) {
            cudaStreamSynchronize(last_dev_stream);
            dev_work_pending = false;
        }
    }
    std::mutex mu;

    size_t row_bytes() const {
        switch (cfg.kind) {
            case DFX_FLAT:
            case DFX_IVF_FLAT: return (size_t)cfg.d * 4;
            case DFX_IVF_PQ: return (size_t)M;
            case DFX_IVF_SQ16: return (size_t)cfg.d * 2;
        }
        return 0;
    }
    int64_t ntotal() const { return n_sorted + n_pending; }
    bool is_ivf() const { return cfg.kind != DFX_FLAT; }
};

// ---- dfx_search.cu
// values[q][j] = ranking value of row (col0+j) of X for query q (IP: -q.x ; L2: |x|^2 - 2 q.x)
void dfx_launch_gemm_values(const float* Q, int64_t nq, const float* X, const float* xnorm,
                            int64_t ncols, int d, int metric, float* out, int64_t ld_out,
                            cudaStream_t st);
void dfx_launch_row_norms(const float* X, int64_t n, int d, float* out, cudaStream_t st);
// top-k columns per row of a values matrix -> keys int32[nrows,k] (and optional values)
void dfx_launch_select_cols(const float* vals, int64_t nrows, int n, int64_t ld, int k,
                            uint32_t col_base, int32_t* keys, float* kvals, uint64_t* comp_out,
                            int64_t comp_ld, cudaStream_t st);
void dfx_search_impl(dfx_index* idx, int64_t nq, const float* d_x, int64_t k, float* d_D,
                     int64_t* d_I, cudaStream_t st);
void dfx_merge_impl(int64_t S, int64_t nq, int64_t k, const float* d_D, const int64_t* d_I,
                    int negate, float* d_outD, int64_t* d_outI, cudaStream_t st);
void dfx_merge_packed_impl(int64_t R, int64_t S_loc, int64_t nq, int64_t k, const void* d_packed,
                           int64_t rank_stride, int64_t off_I, int negate, float* d_outD,
                           int64_t* d_outI, cudaStream_t st);
void dfx_encode_ids_impl(int64_t n, const int64_t* d_ids, int64_t tag, const int32_t* d_col,
                         int32_t drop_code, int64_t* d_out, cudaStream_t st);
void dfx_filter_compact_impl(int64_t nq, int64_t kin, int64_t kout, const float* d_D, const int64_t* d_I,
                             float* d_outD, int64_t* d_outI, int32_t* d_count, cudaStream_t st);
void dfx_map_ids_impl(int64_t n, const int64_t* d_ids, const int64_t* d_table, int64_t* d_out,
                      cudaStream_t st);
void dfx_stats_impl(dfx_index* idx, int64_t* ndis, cudaStream_t st);

void dfx_launch_select_comp(const uint64_t* comp, int64_t nrows, int n, int64_t ld, int k, int32_t* keys,
                            cudaStream_t st);

// Interleaved IVF-PQ block (M == 32): 32 vectors x 32 codes = 1 KB.  Vector v = 8u + w of the
// block (u = group 0..3, w = 0..7) and subquantizer m = i + 8j (i = 0..7, j = 0..3) live at
//   byte  lane*32 + r*4 + t   with lane = 8u + i,  r = w ^ i,  t = (j - u) & 3.
// Lane (u,i) of the scanning warp owns subquantizers {i, i+8, i+16, i+24} of the 8 vectors of
// group u: row r (one 32-bit word) holds their 4 codes in the order the lane looks them up
// (j = (t + u) & 3, which makes the 32 simultaneous table reads hit 32 different banks).
__host__ __device__ __forceinline__ int dfx_il_byte(int v, int m) {
    const int u = v >> 3, w = v & 7, i = m & 7, j = m >> 3;
    return (8 * u + i) * 32 + (w ^ i) * 4 + ((j - u) & 3);
}

// Layout 2 (dfx_scan_il2.cu): one lane per vector.  Lane v of the scanning warp owns vector v of
// the block and walks its 32 subquantizers in the rotated order m = (t + v) & 31, t = 0..31, so
// that at every step the 32 lanes read 32 different table columns (bank == column).  Byte t of
// the lane's 32 code bytes therefore holds the code of subquantizer (t + v) & 31; the two 16-byte
// halves of all lanes are stored contiguously (bytes 0..511: t = 0..15 of lanes 0..31, bytes
// 512..1023: t = 16..31) so that each of the two 128-bit loads of a warp is one 512-byte run.
__host__ __device__ __forceinline__ int dfx_il2_byte(int v, int m) {
    const int t = (m - v) & 31;
    return (t >> 4) * 512 + v * 16 + (t & 15);
}
// Layout 3: the words of layout 1 with each lane's two 16-byte halves stored 512 bytes apart.
__host__ __device__ __forceinline__ int dfx_il3_byte(int v, int m) {
    const int b = dfx_il_byte(v, m), lane = b >> 5, r = (b >> 2) & 7, t = b & 3;
    return (r >> 2) * 512 + lane * 16 + (r & 3) * 4 + t;
}
__host__ __device__ __forceinline__ int dfx_il_byte_of(int layout, int v, int m) {
    return layout == 2 ? dfx_il2_byte(v, m) : layout == 3 ? dfx_il3_byte(v, m) : dfx_il_byte(v, m);
}

// ---- dfx_scan_il.cu
bool dfx_il_wanted(const dfx_index* idx);
void dfx_pq_rm_to_il(dfx_index* idx, cudaStream_t st);
void dfx_pq_il_to_rm(dfx_index* idx, cudaStream_t st);
void dfx_launch_scan_pq_il(dfx_index* idx, int64_t qc, const int32_t* keys, int nprobe, int G, int ngroups,
                           int k, int cap, uint64_t* part, cudaStream_t st);  // layout 1 or 3
// ---- dfx_scan_il2.cu  (lutW: [nq][256][64] wide table, see pq_prep_kernel mode 2)
void dfx_launch_scan_pq_il2(dfx_index* idx, int64_t qc, const int32_t* keys, int nprobe, int G, int ngroups,
                            int k, int cap, uint64_t* part, cudaStream_t st);

// ---- dfx_tc.cu
bool dfx_tc_supported(int d);
void dfx_tc_prepare_centroids(dfx_index* idx, cudaStream_t st);
void dfx_tc_coarse_search(dfx_index* idx, const float* d_x, int64_t nq, int nprobe, int32_t* keys,
                          cudaStream_t st);
int dfx_tc_flat_candidates(dfx_index* idx, const float* d_x, int64_t nq, int k, cudaStream_t st);
void dfx_tc_assign(dfx_index* idx, int d, const float* d_cent, const float* d_cnorm, int64_t nlist, int metric,
                   int64_t n, const float* d_x, int32_t* d_assign, cudaStream_t st);

// ---- dfx_build.cu
void dfx_train_impl(dfx_index* idx, int64_t n, const float* d_x, cudaStream_t st);
void dfx_add_impl(dfx_index* idx, int64_t n, const float* d_x, cudaStream_t st);
void dfx_finalize_impl(dfx_index* idx, cudaStream_t st);
// tag < 0: d_ids are shard-local ids (unknown / -1 rows become NaN); tag >= 0: d_ids are exchange
// ids and only the rows owned by shard `tag` are written
void dfx_reconstruct_impl(dfx_index* idx, int64_t n, const int64_t* d_ids, float* d_out,
                          cudaStream_t st, int64_t tag = -1);
// fused nearest-centroid (GEMM + argmin epilogue), no values matrix
void dfx_launch_assign_fused(const float* X, int64_t n, const float* cent, const float* cnorm,
                             int64_t nlist, int d, int metric, unsigned long long* best,
                             int32_t* out, cudaStream_t st);
void dfx_assign_impl(dfx_index* idx, const float* d_cent, const float* d_cnorm, int64_t nlist,
                     int metric, int d, int64_t n, const float* d_x, int32_t* d_assign,
                     cudaStream_t st);
void dfx_compute_tvals_sorted(dfx_index* idx, cudaStream_t st);
