// icpf_internal.h -- declarations shared between the translation units of libicpflow_b200.so (not installed).
#pragma once

#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include "../../include/icpflow_b200.h"

// Kernel launches and dynamic shared memory go through these two macros so that the same sources also compile for the
// SIMT-on-CPU emulator of the test-suite (tests/simt/, test infrastructure only: the product is the nvcc build).
#ifdef ICPF_SIMT_EMU
#define ICPF_LAUNCH(kernel, grid, block, smem, stream) simt::bind(kernel, dim3(grid), dim3(block), (size_t)(smem))
#else
#define ICPF_LAUNCH(kernel, grid, block, smem, stream) kernel<<<(grid), (block), (smem), (stream)>>>
#endif

namespace icpf {

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// iters [P] int32 | conv [P,4] uint32 | batch [2] int32 + flags (256 B slot) | stats [P,2] int32 {full searches, cache
// refreshes} | history [P, kIcpHistDepth, kIcpHistFloats] fp32: (R, T, rmse) after each of the first 128 iterations of a pair
// | state [P, kIcpStateWords]: the loop state of a pair paused at the cap of the first pass
constexpr int kIcpHistDepth = 128;     // = the iterations the 128-bit convergence masks cover: a batch stop lies inside
constexpr int kIcpHistFloats = 13;     // R[9] T[3] rmse
inline size_t icp_ws_off_conv(int P) { return align_up((size_t)P * 4, 256); }
inline size_t icp_ws_off_batch(int P) { return icp_ws_off_conv(P) + align_up((size_t)P * 16, 256); }
inline size_t icp_ws_off_stats(int P) { return icp_ws_off_batch(P) + 256; }
inline size_t icp_ws_off_hist(int P) { return icp_ws_off_stats(P) + align_up((size_t)P * 8, 256); }
constexpr int kIcpStateWords = 36;     // loop state a pair leaves behind when the capped first pass pauses it (icpf_icploop.cuh)
inline size_t icp_ws_off_state(int P) { return icp_ws_off_hist(P) + align_up((size_t)P * kIcpHistDepth * kIcpHistFloats * 4, 256); }
inline size_t icp_workspace_bytes(int P) { return icp_ws_off_state(P) + align_up((size_t)P * kIcpStateWords * 4, 256); }

// The ICP call split at the two points where the reference's batch stop couples the pairs (utils_icp_pytorch3d.py:209),
// for callers whose batch is spread over several devices (icp_flow_b200/shard.py): between the phases the caller ANDs
// the convergence masks of all shards and hands the stop it found to the last phase.
//   phase 0  first pass (capped at 32 iterations when early exit applies); and_out[4] = AND of this shard's masks
//   phase 1  full pass for the pairs still moving at the cap (only when phase 0 could not decide); and_out likewise
//   phase 2  batch = {batch_iters, converged} as given; pairs beyond it read their state back; outputs are final
struct IcpPhase {
    int phase;
    int batch_iters;
    int converged;
    uint32_t* and_out;     // device [4], phases 0 and 1
};

int launch_icp(const float* src, const float* dst, const float* init_R, const float* init_T, const float* init_pose,
               int auto_swap, int P, int N, const icpf_params& prm, float* out_R, float* out_T,
               float* out_rmse, float* out_pose, int* out_iters, uint32_t* out_conv, int* out_batch, void* workspace,
               size_t workspace_bytes, cudaStream_t stream, const IcpPhase* phase = nullptr,
               const icpf_icp_ext* ext = nullptr);
int launch_peer_push(const float* local_pose, float* const* peer_pose_dev, int world, int row0, int P, cudaStream_t stream);
int launch_expand_rows(const float* rows, const int32_t* offsets, int B, int N, float* out, cudaStream_t stream);

size_t icp_big_workspace_bytes(int P, int N);     // 0 unless the clusters need the global-memory variant
int hist_chunk_pairs(int P, int lx, int ly, int lz);
size_t path_workspace_bytes(int P, int N, int lx, int ly, int lz);

int launch_hist_votes(const float* X, const float* Y, int B, int NX, int NY, const float* mins, const float* maxs,
                      const int* lens, float* bins, int auto_swap, const int* need, cudaStream_t stream);
size_t hist_peaks_scratch_floats(int lx, int ly);      // per pair, 0 when the max planes fit shared memory
int launch_hist_peaks(const float* bins, int B, int lx, int ly, int lz, int* out_idx, float* out_votes,
                      const int* need, float* scratch, cudaStream_t stream);
int launch_hist_fused(const float* X, const float* Y, int P, int N, const float* mins, const float* maxs,
                      const int* lens, int auto_swap, int* out_idx, float* out_votes, int* need_global,
                      cudaStream_t stream);
int launch_hist_score(const float* src, const float* dst, int P, int N, const int* cand_idx, const float* bins_x,
                      const float* bins_y, const float* bins_z, int lx, int ly, int lz, float half_bin, float tau,
                      int auto_swap, float* out_pose, float* out_scores, int* out_which, int* defer, cudaStream_t stream);
size_t score_defer_words(int P);
int launch_icp_finalize(const float* src, const float* dst, int P, int N, const float* init_pose, const float* icp_R,
                        const float* icp_T, int auto_swap, float tau, float* out_pose, float* out_err, int* out_flags,
                        cudaStream_t stream);
int launch_hist_init(const float* src, const float* dst, int P, int N, const icpf_hist_bins& hb, int auto_swap,
                     float* out_pose, int* out_cand, float* out_votes, float* out_scores, int* out_which,
                     void* workspace, size_t workspace_bytes, cudaStream_t stream);
int launch_apply_icp(const float* src, const float* dst, const float* init_pose, int P, int N, const icpf_params& prm,
                     int auto_swap, float* out_pose, float* out_err, int* out_flags, int* out_batch, void* workspace,
                     size_t workspace_bytes, cudaStream_t stream, const IcpPhase* phase = nullptr);
int launch_hist_icp(const float* src, const float* dst, int P, int N, const icpf_hist_bins& hb, const icpf_params& prm,
                    float* out_pose, float* out_init, int* out_batch, void* workspace, size_t workspace_bytes,
                    cudaStream_t stream);

int launch_nn(const float* src, const float* dst, int B, int Ns, int Nd, int src_stride, int dst_stride,
              int64_t* out_idx, float* out_dist, cudaStream_t stream);

int launch_match_eval(const float* src, const float* dst, const float* pose, int P, int N, float thr, float* errors,
                      float* inliers, float* ratios, float* ious, float* translations, float* rotations,
                      const icpf_match_gates* gates, int* accept, cudaStream_t stream);
int launch_transform_points(const float* xyz, const float* pose, int B, int N, float* out, cudaStream_t stream);


// scan-level callers (icpf_scan.cu)
size_t cluster_index_workspace_bytes(int n, int n_labels);
int launch_cluster_index(const float* points, int stride, const float* labels, int n, int n_labels, int* order,
                         int* offsets, float* stats, void* ws, size_t ws_bytes, cudaStream_t stream);
int launch_sanity_check(const int* src_offsets, const float* src_stats, int n_src, const int* dst_offsets,
                        const float* dst_stats, int n_dst, const int64_t* pairs, int P, int min_cluster_size,
                        float translation_frame, float thres_box, int* out_keep, int64_t* out_pairs, int* out_count,
                        cudaStream_t stream, int cross_nd = 0);
size_t match_select_workspace_bytes(int ns, int nd);
int hdbscan_labels_host(const int* edge_a, const int* edge_b, const double* edge_w, int n, int min_cluster_size,
                        int presorted, int* labels);
size_t hdbscan_workspace_bytes(int n);
int launch_hdbscan_mst(const float* points, int stride, int n, int min_samples, int prim_order, double* out_core,
                       int* out_src, int* out_dst, double* out_w, void* workspace, cudaStream_t stream);
int launch_match_select(const int64_t* pairs, int P, const int64_t* src_unq, int ns, const int64_t* dst_unq, int nd,
                        const float* errors, const float* inliers, const float* ratios, const float* ious,
                        const int* accept, const float* transforms, float thres_error, float* out_rows,
                        float* out_transforms, int64_t* out_src_left, int64_t* out_dst_left, int* out_counts,
                        void* workspace, cudaStream_t stream);
int launch_gather_pairs(const float* src_points, int src_stride, const int* src_order, const int* src_offsets, int n_src,
                        const float* dst_points, int dst_stride, const int* dst_order, const int* dst_offsets, int n_dst,
                        const int64_t* pairs, int P, int max_points, const int* sample_rows,
                        const int64_t* sample_offsets, float* out_src, float* out_dst, cudaStream_t stream);
int launch_flow(const float* points, int stride, const float* labels, int n, const float* pair_labels, int pair_stride,
                const float* transforms, int K, const float* pose, float* flow, cudaStream_t stream);

// clustering (icpf_cluster.cu)
size_t dbscan_workspace_bytes(int n);
int launch_dbscan(const float* points, int stride, int n, double eps, int min_points, int* out_labels, int* out_num_clusters,
                  void* workspace, size_t workspace_bytes, cudaStream_t stream);

}  // namespace icpf
