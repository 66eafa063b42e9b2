// icpf_apply.cu -- the wrapper around the ICP loop (utils_icp.apply_icp, utils_icp.py:20-48) and the orchestration of
// the whole per-pair path (utils_match.hist_icp, utils_match.py:138-157) as stream-ordered launches:
//
//   hist_votes -> hist_peaks -> hist_score            (init pose, icpf_hist.cu; histogram chunked over pairs)
//   icp_pairs (init pose applied on load) -> resolve batch stop -> re-run pass        (icpf_icp.cu)
//   icp_finalize: compose ICP with the init pose, mean NN error before/after, roll back, undo the swap
#include "icpf_internal.h"
#include "icpf_pair.cuh"
#include "icpf_gridnn.cuh"

namespace icpf {

struct FinalizeArgs {
    const float* src;        // [P,N,4]
    const float* dst;        // [P,N,4]
    int N;
    const float* init_pose;  // [P,16]
    const float* icp_R;      // [P,9]  row-vector convention
    const float* icp_T;      // [P,3]
    int auto_swap;
    float tau;               // thres_dist: only sizes the cells of the NN grid
    float* out_pose;         // [P,16]
    float* out_err;          // [P,2] {error_init, error_icp} (may be NULL)
    int* out_flags;          // [P]   bit 0 rolled back, bit 1 swapped (may be NULL)
};

// mean over the valid rows of S of the unbounded NN distance of (pose * S_i) among D[0, n_d)   (utils_icp.py:28-33)
// Rows are accumulated per thread in increasing row order (q = tid, tid + kThreads, ...) whichever search is used; the
// grid search (icpf_gridnn.cuh) returns the same minimum as the full scan, so both variants produce the same bits.
template <bool GRIDNN>
__device__ __forceinline__ float mean_nn_error(const float (&m)[12], const float4* S, int n_s, const float4* D, int n_d,
                                               float* scratch, const GridInfo& g, const float4* sorted,
                                               const unsigned short* runs) {
    float sum[1] = {0.f};
    if (GRIDNN && n_d > 0) {
        for (int q0 = 0; q0 < n_s; q0 += kThreads) {          // (warp-uniform trip count: the far queries are scanned by the warp)
            const int q = q0 + threadIdx.x;
            const bool active = q < n_s;
            const float4 raw = active ? S[q] : make_float4(0.f, 0.f, 0.f, 0.f);
            const float4 s = transform_row(m, raw);
            const float best = nn_unbounded_grid_warp<false>(g, sorted, runs, n_d, active, s.x, s.y, s.z, s.x, s.y, s.z, 0.f,
                                                             0.f, 0.f);
            if (active && raw.w > 0.f) sum[0] += sqrtf(best);
        }
    } else {
        constexpr int QB = 4;
        for (int q0 = threadIdx.x; q0 < n_s; q0 += kThreads * QB) {
            float qx[QB], qy[QB], qz[QB], best[QB];
            int bidx[QB];
#pragma unroll
            for (int k = 0; k < QB; ++k) {
                const int q = q0 + k * kThreads;
                const float4 s = transform_row(m, q < n_s ? S[q] : make_float4(0.f, 0.f, 0.f, 0.f));
                qx[k] = s.x; qy[k] = s.y; qz[k] = s.z;
            }
            nn_brute<QB>(D, n_d, qx, qy, qz, best, bidx);
#pragma unroll
            for (int k = 0; k < QB; ++k) {
                const int q = q0 + k * kThreads;
                if (q < n_s && S[q].w > 0.f) sum[0] += sqrtf(best[k]);
            }
        }
    }
    block_allreduce_sum<1, kWarps>(sum, scratch);
    __syncthreads();
    return sum[0];
}

// GRIDNN: the fixed cloud is counting-sorted into a uniform grid in shared memory and the two error passes are grid
// searches (same minima as the full scans, same bits); the rows in storage order are streamed from global memory.
// !GRIDNN (the grid does not fit shared memory): full scans over the rows in global memory.
template <bool GRIDNN>
__global__ void __launch_bounds__(kThreads) icp_finalize_kernel(FinalizeArgs a) {
    const int p = blockIdx.x, tid = threadIdx.x;
    __shared__ float s_scratch[kWarps * 4];
    const float4* S = reinterpret_cast<const float4*>(a.src) + (size_t)p * a.N;
    const float4* D = reinterpret_cast<const float4*>(a.dst) + (size_t)p * a.N;
    float cnt[2] = {0.f, 0.f};
    for (int q = tid; q < a.N; q += kThreads) {
        cnt[0] += (S[q].w > 0.f) ? 1.f : 0.f;
        cnt[1] += (D[q].w > 0.f) ? 1.f : 0.f;
    }
    block_allreduce_sum<2, kWarps>(cnt, s_scratch);
    __syncthreads();
    int n_s = (int)cnt[0], n_d = (int)cnt[1];
    const bool swapped = a.auto_swap && n_s > n_d;
    if (swapped) {
        const float4* t = S; S = D; D = t;
        const int n = n_s; n_s = n_d; n_d = n;
    }
    // NN grid over the fixed cloud (after the role swap), shared by the two error passes
    GridInfo g;
    const float4* sorted = nullptr;
    const unsigned short* runs = nullptr;
    if constexpr (GRIDNN) {
        float4* extra = g_tile;
        GridTiles td{D, extra, reinterpret_cast<uint32_t*>(extra + a.N), reinterpret_cast<float*>(extra + gridnn_units(a.N))};
        if (n_d > 0) g = build_grid(td, n_d, a.tau);
        sorted = td.sorted_p;
        runs = reinterpret_cast<const unsigned short*>(td.cells_p);
    }
    // M0 = init pose, Micp = [[R^T, T],[0,1]] (utils_icp.py:60-65), M = Micp * M0 (utils_icp.py:24)
    float m0[16], mi[16], mm[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) m0[i] = a.init_pose[(size_t)p * 16 + i];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
#pragma unroll
        for (int c = 0; c < 3; ++c) mi[4 * r + c] = a.icp_R[(size_t)p * 9 + c * 3 + r];
        mi[4 * r + 3] = a.icp_T[(size_t)p * 3 + r];
    }
    mi[12] = mi[13] = mi[14] = 0.f;
    mi[15] = 1.f;
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c)
            mm[4 * r + c] = fmaf(mi[4 * r + 3], m0[12 + c],
                                 fmaf(mi[4 * r + 2], m0[8 + c], fmaf(mi[4 * r + 1], m0[4 + c], mi[4 * r] * m0[c])));
    float a0[12], a1[12];
#pragma unroll
    for (int i = 0; i < 12; ++i) { a0[i] = m0[i]; a1[i] = mm[i]; }
    const float e0 = __fdiv_rn(mean_nn_error<GRIDNN>(a0, S, n_s, D, n_d, s_scratch, g, sorted, runs), (float)n_s);
    const float e1 = __fdiv_rn(mean_nn_error<GRIDNN>(a1, S, n_s, D, n_d, s_scratch, g, sorted, runs), (float)n_s);
    if (tid == 0) {
        const bool rolled = e1 >= e0;          // utils_icp.py:34-35 (NaN compares false: keep the ICP result)
        float out[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) out[i] = rolled ? m0[i] : mm[i];
        if (swapped) {
            // torch.linalg.inv of the 4x4 (utils_match.py:152-154); the last row is (0,0,0,1) by construction, so
            // the inverse is [[A^-1, -A^-1 b],[0,1]] with A^-1 = adj(A) / det(A)
            const float A00 = out[0], A01 = out[1], A02 = out[2], A10 = out[4], A11 = out[5], A12 = out[6], A20 = out[8],
                        A21 = out[9], A22 = out[10], b0 = out[3], b1 = out[7], b2 = out[11];
            const float c00 = A11 * A22 - A12 * A21, c01 = A02 * A21 - A01 * A22, c02 = A01 * A12 - A02 * A11;
            const float c10 = A12 * A20 - A10 * A22, c11 = A00 * A22 - A02 * A20, c12 = A02 * A10 - A00 * A12;
            const float c20 = A10 * A21 - A11 * A20, c21 = A01 * A20 - A00 * A21, c22 = A00 * A11 - A01 * A10;
            const float det = A00 * c00 + A01 * c10 + A02 * c20;
            const float id = 1.0f / det;
            const float i00 = c00 * id, i01 = c01 * id, i02 = c02 * id, i10 = c10 * id, i11 = c11 * id, i12 = c12 * id,
                        i20 = c20 * id, i21 = c21 * id, i22 = c22 * id;
            out[0] = i00; out[1] = i01; out[2] = i02; out[3] = -(i00 * b0 + i01 * b1 + i02 * b2);
            out[4] = i10; out[5] = i11; out[6] = i12; out[7] = -(i10 * b0 + i11 * b1 + i12 * b2);
            out[8] = i20; out[9] = i21; out[10] = i22; out[11] = -(i20 * b0 + i21 * b1 + i22 * b2);
            out[12] = 0.f; out[13] = 0.f; out[14] = 0.f; out[15] = 1.f;
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) a.out_pose[(size_t)p * 16 + i] = out[i];
        if (a.out_err) {
            a.out_err[(size_t)p * 2] = e0;
            a.out_err[(size_t)p * 2 + 1] = e1;
        }
        if (a.out_flags) a.out_flags[p] = (rolled ? 1 : 0) | (swapped ? 2 : 0);
    }
}

int launch_icp_finalize(const float* src, const float* dst, int P, int N, const float* init_pose, const float* icp_R,
                        const float* icp_T, int auto_swap, float tau, float* out_pose, float* out_err, int* out_flags,
                        cudaStream_t stream) {
    if (P == 0) return ICPF_OK;
    const size_t with_grid = ((size_t)gridnn_units(N) + up16(kRedFloats * 4)) * 16;
    const bool gridnn = with_grid <= (size_t)227 * 1024 && tau > 0.f;
    const size_t smem = gridnn ? with_grid : 0;
    auto kernel = gridnn ? icp_finalize_kernel<true> : icp_finalize_kernel<false>;
    cudaError_t err = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess) return (int)err;
    FinalizeArgs a{src, dst, N, init_pose, icp_R, icp_T, auto_swap, tau, out_pose, out_err, out_flags};
    ICPF_LAUNCH(kernel, P, kThreads, smem, stream)(a);
    return (int)cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------ orchestration
// workspace: [icp workspace][R P*9][T P*3][init pose P*16][cand idx P*5][votes P*5][need P][scoring items][histogram chunk]
size_t path_workspace_bytes(int P, int N, int lx, int ly, int lz) {
    size_t b = icp_workspace_bytes(P) + icp_big_workspace_bytes(P, N);
    b += align_up((size_t)P * 9 * 4, 256) + align_up((size_t)P * 3 * 4, 256) + align_up((size_t)P * 16 * 4, 256);
    b += 2 * align_up((size_t)P * 5 * 4, 256) + align_up((size_t)P * 4, 256);
    b += align_up(score_defer_words(P) * 4, 256);
    // histogram chunk [chunk, lx, ly, lz] (+ [chunk, 2, lx, ly] max planes when they do not fit shared memory)
    if (lx > 0 && ly > 0 && lz > 0)
        b += align_up((size_t)hist_chunk_pairs(P, lx, ly, lz) * ((size_t)lx * ly * lz + hist_peaks_scratch_floats(lx, ly)) * 4, 256);
    return b;
}

int hist_chunk_pairs(int P, int lx, int ly, int lz) {
    const size_t per = ((size_t)lx * ly * lz + hist_peaks_scratch_floats(lx, ly)) * 4;
    const size_t budget = (size_t)64 << 20;          // keep the live histograms L2-resident (126 MB L2)
    size_t c = budget / (per ? per : 1);
    if (c < 1) c = 1;
    if (c > 32768) c = 32768;                        // gridDim.y of the vote kernel
    if (c > (size_t)P) c = (size_t)P;
    return (int)(c > 0 ? c : 1);
}

struct PathWs {
    unsigned char* icp;
    float* R;
    float* T;
    float* init;
    int* cand;
    float* votes;
    int* need;      // [P] pairs that take the global-memory histogram path
    int* defer;     // candidate-scoring items (icpf_hist.cu: score_defer_words)
    float* hist;
};

static PathWs carve_ws(void* workspace, int P, int N) {
    PathWs w;
    unsigned char* p = static_cast<unsigned char*>(workspace);
    w.icp = p; p += icp_workspace_bytes(P) + icp_big_workspace_bytes(P, N);
    w.R = reinterpret_cast<float*>(p); p += align_up((size_t)P * 9 * 4, 256);
    w.T = reinterpret_cast<float*>(p); p += align_up((size_t)P * 3 * 4, 256);
    w.init = reinterpret_cast<float*>(p); p += align_up((size_t)P * 16 * 4, 256);
    w.cand = reinterpret_cast<int*>(p); p += align_up((size_t)P * 5 * 4, 256);
    w.votes = reinterpret_cast<float*>(p); p += align_up((size_t)P * 5 * 4, 256);
    w.need = reinterpret_cast<int*>(p); p += align_up((size_t)P * 4, 256);
    w.defer = reinterpret_cast<int*>(p); p += align_up(score_defer_words(P) * 4, 256);
    w.hist = reinterpret_cast<float*>(p);
    return w;
}

int launch_hist_init(const float* src, const float* dst, int P, int N, const icpf_hist_bins& hb, int auto_swap,
                     float* out_pose, int* out_cand, float* out_votes, float* out_scores, int* out_which,
                     void* workspace, size_t workspace_bytes, cudaStream_t stream) {
    if (P == 0) return ICPF_OK;
    if (workspace == nullptr || workspace_bytes < path_workspace_bytes(P, N, hb.len[0], hb.len[1], hb.len[2]))
        return ICPF_E_WORKSPACE;
    PathWs w = carve_ws(workspace, P, N);
    int* cand = out_cand ? out_cand : w.cand;
    float* votes = out_votes ? out_votes : w.votes;
    // utils_hist.py:69: hist(dst, src, ...) -> votes dst_i - src_j.  Fused shared-memory kernel first; the pairs whose
    // difference box does not fit are flagged and redone through the chunked global-memory histogram.
    int rc = launch_hist_fused(dst, src, P, N, hb.min, hb.max, hb.len, auto_swap, cand, votes, w.need, stream);
    if (rc != ICPF_OK) return rc;
    const int chunk = hist_chunk_pairs(P, hb.len[0], hb.len[1], hb.len[2]);
    for (int lo = 0; lo < P; lo += chunk) {
        const int n = (P - lo < chunk) ? (P - lo) : chunk;
        rc = launch_hist_votes(dst + (size_t)lo * N * 4, src + (size_t)lo * N * 4, n, N, N, hb.min, hb.max, hb.len,
                               w.hist, auto_swap, w.need + lo, stream);
        if (rc != ICPF_OK) return rc;
        rc = launch_hist_peaks(w.hist, n, hb.len[0], hb.len[1], hb.len[2], cand + (size_t)lo * 5,
                               votes + (size_t)lo * 5, w.need + lo,
                               w.hist + (size_t)chunk * hb.len[0] * hb.len[1] * hb.len[2], stream);
        if (rc != ICPF_OK) return rc;
    }
    // bin width of the z axis = thres_dist (utils_hist.py:65: arange(-tau, 2 tau - eps, tau)); it only sizes NN-grid cells
    const float tau = hb.len[2] > 1 ? (hb.max[2] - hb.min[2]) / (float)(hb.len[2] - 1) : 0.1f;
    return launch_hist_score(src, dst, P, N, cand, hb.bins_x, hb.bins_y, hb.bins_z, hb.len[0], hb.len[1], hb.len[2],
                             hb.half_bin, tau, auto_swap, out_pose, out_scores, out_which, w.defer, stream);
}

int launch_apply_icp(const float* src, const float* dst, const float* init_pose, int P, int N, const icpf_params& prm,
                     int auto_swap, float* out_pose, float* out_err, int* out_flags, int* out_batch, void* workspace,
                     size_t workspace_bytes, cudaStream_t stream, const IcpPhase* phase) {
    if (P == 0) return ICPF_OK;
    if (workspace == nullptr || workspace_bytes < path_workspace_bytes(P, N, 0, 0, 0)) return ICPF_E_WORKSPACE;
    PathWs w = carve_ws(workspace, P, N);
    // (a phased call keeps iterations and masks in the workspace between its phases: out_iters / out_conv stay NULL)
    int rc = launch_icp(src, dst, nullptr, nullptr, init_pose, auto_swap, P, N, prm, w.R, w.T, nullptr, nullptr,
                        nullptr, nullptr, out_batch, w.icp, icp_workspace_bytes(P) + icp_big_workspace_bytes(P, N), stream,
                        phase);
    if (rc != ICPF_OK) return rc;
    if (phase != nullptr && phase->phase < 2) return ICPF_OK;
    return launch_icp_finalize(src, dst, P, N, init_pose, w.R, w.T, auto_swap, (float)prm.thres_dist, out_pose, out_err,
                               out_flags, stream);
}

int launch_hist_icp(const float* src, const float* dst, int P, int N, const icpf_hist_bins& hb, const icpf_params& prm,
                    float* out_pose, float* out_init, int* out_batch, void* workspace, size_t workspace_bytes,
                    cudaStream_t stream) {
    if (P == 0) return ICPF_OK;
    if (workspace == nullptr || workspace_bytes < path_workspace_bytes(P, N, hb.len[0], hb.len[1], hb.len[2]))
        return ICPF_E_WORKSPACE;
    PathWs w = carve_ws(workspace, P, N);
    float* init = out_init ? out_init : w.init;
    int rc = launch_hist_init(src, dst, P, N, hb, 1, init, nullptr, nullptr, nullptr, nullptr, workspace,
                              workspace_bytes, stream);
    if (rc != ICPF_OK) return rc;
    return launch_apply_icp(src, dst, init, P, N, prm, 1, out_pose, nullptr, nullptr, out_batch, workspace,
                            workspace_bytes, stream);
}

}  // namespace icpf
