/*
 * icpflow_b200.h -- C ABI of the B200-native batched ICP registration engine (libicpflow_b200.so).
 *
 * Drop-in boundary for ICP-Flow's per-cluster-pair alignment path.  Plain pointers and sizes only: no torch /
 * ATen / pybind11 types cross this boundary (the reference's native boundary is a pybind11+ATen extension,
 * /root/reference/hist_cuda/cpp/hist.h:4-10, hist.cpp:25-27; everything else on the path is Python calling
 * torch/pytorch3d ops).  All `*_f32` entry points
 *   - take DEVICE pointers (fp32, contiguous, the reference's padded layout: [P, N, 4] rows (x, y, z, flag),
 *     valid rows first with flag > 0, padded rows (1e8, 1e8, 1e8, 0) -- utils_helper.py:185-196),
 *   - enqueue their work on `stream` (a cudaStream_t passed as void*) and return without synchronising,
 *   - never allocate or free device memory (the caller provides outputs and, where stated, a workspace),
 *   - return 0 on success, a negative ICPF_E_* code for argument errors, or a positive cudaError_t.
 * Host buffers are staged by the caller (icp_flow_b200.ops.IcpHostPipeline: pinned memory, double-buffered H2D on a
 * copy stream overlapping the kernels of the previous batch).
 */
#ifndef ICPFLOW_B200_H
#define ICPFLOW_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ICPF_OK 0
#define ICPF_E_NULL (-1)        /* required pointer is NULL                         */
#define ICPF_E_SHAPE (-2)       /* P/N/len out of the supported range               */
#define ICPF_E_PARAM (-3)       /* invalid parameter value                          */
#define ICPF_E_ALIGN (-4)       /* pointer not 16-byte aligned                      */
#define ICPF_E_WORKSPACE (-5)   /* workspace too small                              */
#define ICPF_E_UNSUPPORTED (-6) /* configuration not implemented on this build      */

#define ICPF_MAX_ITERATIONS 128 /* convergence history is a 128-bit mask per pair   */

/* The values of `args` and the hard-coded constants the reference path reads
 * (main.py:45-132; utils_icp.py:54-55; utils_hist.py:21). */
typedef struct icpf_params {
    double thres_dist;        /* args.thres_dist (tau): ICP gate d^2 <= fp32(tau^2), histogram bin width       */
    int32_t max_iterations;   /* utils_icp.py:54 -> 100                                                        */
    float relative_rmse_thr;  /* utils_icp.py:55 -> 1e-6; negative disables the batch stop                     */
    int32_t early_exit;       /* 1: a pair stops at its bitwise fixed point (result-identical, SURVEY finding 1);
                                 0: every pair executes the full batch iteration count                         */
    int32_t batch_stop;       /* 1: reproduce the reference's batch-coupled stop (utils_icp_pytorch3d.py:209):
                                 results are the state at the first iteration where ALL pairs satisfy the
                                 relative-RMSE test; 0: each pair independent (max_iterations / fixed point)   */
    int32_t nn_mode;          /* 0 auto (= 3 when the tiles fit), 1 brute force, 2 uniform grid (radius-bounded),
                                 3 grid + correspondence cache; all modes are result-identical                 */
    int32_t reserved[2];      /* [0]: tuning knob, grid cell size in 1/1000 of the gate radius (0 = default)     */
} icpf_params;

/* Histogram geometry of the translation initialisation.  The reference builds the bin starts with torch.arange in
 * the default dtype and passes bins.min() / bins.max() -- the LAST BIN START, not the upper edge -- to the vote
 * kernel, then decodes the winning bins with bins[idx] (utils_hist.py:63-78); the shim therefore builds the three
 * arrays with the same torch call and hands them over instead of recomputing edges in C. */
typedef struct icpf_hist_bins {
    const float* bins_x;      /* DEVICE [len[0]]  torch.arange(-F, F + tau - 1e-8, tau)                          */
    const float* bins_y;      /* DEVICE [len[1]]                                                                 */
    const float* bins_z;      /* DEVICE [len[2]]  torch.arange(-tau, 2 tau - 1e-8, tau)                          */
    int32_t len[3];
    float min[3];             /* bins.min()                                                                      */
    float max[3];             /* bins.max()                                                                      */
    float half_bin;           /* args.thres_dist // 2 (python floor division: 0.0 for tau = 0.1)                 */
} icpf_hist_bins;

int icpf_version(void);
const char* icpf_error_string(int code);

/* Fill `p` with the reference defaults (tau 0.1, 100 iterations, thr 1e-6, early_exit 1, batch_stop 1). */
void icpf_default_params(icpf_params* p);

/* Bytes of device workspace needed by the calls below for P pairs of N padded rows and an lx*ly*lz histogram.
 * (About 6.7 KB per pair -- the (R, T, rmse) record of up to 128 ICP iterations the batch stop is read back from -- plus
 * a histogram chunk of at most 64 MB, plus per-pair scratch for clusters too large for shared memory.) */
size_t icpf_workspace_bytes(int32_t P, int32_t N, int32_t lx, int32_t ly, int32_t lz);

/*
 * Batched ICP loop -- replaces utils_icp_pytorch3d.iterative_closest_point (utils_icp_pytorch3d.py:37-225) with
 * init_transform=None, estimate_scale=False, allow_reflection=False, and the Kabsch solve it calls
 * (corresponding_points_alignment, :233-382).  Row-vector convention of the reference:  X R + T ~ Y[NN].
 *   src, dst   [P,N,4]            out_R [P,9] row-major   out_T [P,3]   out_rmse [P] (may be NULL)
 *   init_R [P,9], init_T [P,3]    the reference's init_transform (both NULL = identity): used for the FIRST
 *                                 correspondence search only, the transform is always re-estimated from `src`
 *   out_pose   [P,16] the same transform packed as the reference's column-convention 4x4 [[R^T, T],[0,1]]
 *              (utils_icp.py:60-65) (may be NULL)
 *   out_iters  [P] int32 iterations this pair executed (may be NULL)
 *   out_conv   [P,4] uint32 bit k set <=> relative rmse <= thr at iteration k (may be NULL)
 *   out_batch  [2] int32: {batch iterations the reference would have executed, converged flag} (may be NULL)
 *   workspace  icpf_workspace_bytes(P, N, 0, 0, 0)
 */
int icpf_icp_f32(const float* src, const float* dst, const float* init_R, const float* init_T, int32_t P, int32_t N,
                 const icpf_params* params, float* out_R, float* out_T, float* out_rmse, float* out_pose,
                 int32_t* out_iters, uint32_t* out_conv, int32_t* out_batch, void* workspace, size_t workspace_bytes,
                 void* stream);

/*
 * Unbounded K=1 nearest neighbour over ALL rows -- replaces utils_helper.nearest_neighbor_batch
 * (utils_helper.py:20-30; pytorch3d knn_points without lengths).  Rows are `stride` floats apart (3 or 4).
 *   src [B,Ns,stride], dst [B,Nd,stride] -> out_idx [B,Ns] int64, out_dist [B,Ns] (sqrt of the min squared L2).
 */
int icpf_nn_f32(const float* src, const float* dst, int32_t B, int32_t Ns, int32_t Nd, int32_t src_stride,
                int32_t dst_stride, int64_t* out_idx, float* out_dist, void* stream);

/*
 * Homogeneous transform that keeps the flag column -- replaces utils_helper.transform_points_batch
 * (utils_helper.py:76-87).  xyz [B,N,4], pose [B,16] row-major 4x4 (column-vector convention) -> out [B,N,4].
 */
int icpf_transform_points_f32(const float* xyz, const float* pose, int32_t B, int32_t N, float* out, void* stream);

/*
 * Registration quality metrics of a batch of pairs -- replaces utils_match.match_eval (utils_match.py:159-213), the
 * consumer of hist_icp's transforms in match_pairs (utils_match.py:93).  One fused launch instead of a transform, two
 * knn_points passes and the masked reductions.
 *   src, dst [P,N,4] (the clouds as given to hist_icp, NOT swapped), pose [P,16] row-major 4x4 (src -> dst)
 *   out_errors [P,2]        mean NN distance of the valid rows: moved src -> dst, dst -> moved src
 *   out_inliers [P,2]       number of valid rows with NN distance < (float)thres_dist (strict), as fp32 like the reference
 *   out_ratios [P,2]        inliers / valid rows
 *   out_ious [P,2]          inliers_a / (n_src + n_dst - inliers_b)
 *   out_translations [P,3]  mean of the moved valid src rows - mean of the valid src rows
 *   out_rotations [P,3]     ZYX Euler angles of pose[:3,:3] in degrees (pytorch3d matrix_to_euler_angles * 180 / pi)
 * A cloud without valid rows gives NaN where the reference divides 0 by 0.
 * gates / out_accept (both NULL or both set): out_accept [P] int32 = utils_check.check_transformation(args, translation,
 * rotation, min(iou)) (utils_check.py:51-66) -- 0 when |translation| > translation_frame, min(iou) < thres_iou, or
 * max(|pitch|,|roll|) > thres_rot * 90 degrees; the per-pair Python loop of match_pairs (utils_match.py:96-100) becomes
 * one flag per pair written by the same launch.
 */
typedef struct icpf_match_gates {
    double translation_frame;   /* args.translation_frame */
    double thres_iou;           /* args.thres_iou */
    double thres_rot;           /* args.thres_rot (fraction of 90 degrees) */
} icpf_match_gates;

int icpf_match_eval_f32(const float* src, const float* dst, const float* pose, int32_t P, int32_t N, double thres_dist,
                        float* out_errors, float* out_inliers, float* out_ratios, float* out_ious,
                        float* out_translations, float* out_rotations, const icpf_match_gates* gates,
                        int32_t* out_accept, void* stream);

/*
 * All-pairs difference histogram -- bit-compatible replacement of HIST.hist (hist_cuda/cpp/hist.cpp:25-27,
 * hist_cuda.cu:19-90, hist_cuda_core.cuh:23-64; python wrapper hist_cuda/hist.py:39-51).  Votes X_i - Y_j over rows
 * whose flags are both > 0.  X [B,NX,4], Y [B,NY,4] -> bins [B,len_x,len_y,len_z] fp32 counts (zero-filled here).
 * min/max/len are HOST arrays of 3.
 */
int icpf_hist_votes_f32(const float* X, const float* Y, int32_t B, int32_t NX, int32_t NY, const float* min_xyz,
                        const float* max_xyz, const int32_t* len_xyz, float* bins, void* stream);

/*
 * Histogram-vote translation initialisation -- replaces utils_hist.estimate_init_pose (utils_hist.py:33-124): votes,
 * 3-D NMS + top-5 (topk_nms, :21-29), candidate scoring by bidirectional mean NN distance, arg-min.
 *   src, dst [P,N,4] -> out_pose [P,16] (identity rotation + best translation)
 *   auto_swap 1: apply hist_icp's "smaller cloud is the moved one" rule per pair (utils_match.py:139-146)
 *   optional diagnostics: out_cand [P,5] int32 flat bin indices, out_votes [P,5], out_scores [P,6], out_which [P] int32
 *   workspace  icpf_workspace_bytes(P, N, len[0], len[1], len[2])
 */
int icpf_hist_init_f32(const float* src, const float* dst, int32_t P, int32_t N, const icpf_hist_bins* bins,
                       int32_t auto_swap, float* out_pose, int32_t* out_cand, float* out_votes, float* out_scores,
                       int32_t* out_which, void* workspace, size_t workspace_bytes, void* stream);

/*
 * ICP from an initial pose with roll-back -- replaces utils_icp.apply_icp / pytorch3d_icp (utils_icp.py:20-73):
 * src' = init * src, ICP(src', dst), T = ICP o init, mean NN error before / after, T = init where it did not drop.
 *   init_pose [P,16] -> out_pose [P,16]; out_err [P,2] {error_init, error_icp}, out_flags [P] int32 (bit 0 rolled
 *   back, bit 1 swapped), out_batch [2] int32 (all optional).  With auto_swap = 1 the result of a swapped pair is
 *   inverted like utils_match.py:152-154.   workspace  icpf_workspace_bytes(P, N, 0, 0, 0)
 */
int icpf_apply_icp_f32(const float* src, const float* dst, const float* init_pose, int32_t P, int32_t N,
                       const icpf_params* params, int32_t auto_swap, float* out_pose, float* out_err,
                       int32_t* out_flags, int32_t* out_batch, void* workspace, size_t workspace_bytes, void* stream);

/*
 * apply_icp in phases, for a batch whose pairs are spread over several devices (one process per GPU): the reference's
 * batch stop (utils_icp_pytorch3d.py:209, `.all()` over the pairs of the call) is the only coupling between pairs, so a
 * sharded caller that wants the transforms of the unsharded batch bit for bit exchanges 16 bytes between the phases:
 *   phase 0  first ICP pass; out_and[4] (device uint32) = AND of the shard's 128-bit convergence masks.  The caller ANDs
 *            the words of all shards: the lowest set bit k* < min(32, max_iterations) is the batch stop.
 *   phase 1  only when phase 0 found none and max_iterations > 32: full pass for the pairs still moving; out_and again.
 *   phase 2  batch_iterations / batch_converged = the stop found over all shards ((k*+1, 1) or (max_iterations, 0)):
 *            pairs that went beyond it read their state at the stop back, roll-back / un-swap as in icpf_apply_icp_f32;
 *            out_pose, out_err, out_flags, out_batch are written by this phase only.
 * Same arguments and the SAME workspace (untouched in between) in every phase; a shard of P = 0 pairs skips the calls
 * and contributes all-ones words.  With one shard the three phases give exactly icpf_apply_icp_f32.
 */
int icpf_apply_icp_phase_f32(const float* src, const float* dst, const float* init_pose, int32_t P, int32_t N,
                             const icpf_params* params, int32_t auto_swap, int32_t phase, int32_t batch_iterations,
                             int32_t batch_converged, uint32_t* out_and, float* out_pose, float* out_err,
                             int32_t* out_flags, int32_t* out_batch, void* workspace, size_t workspace_bytes,
                             void* stream);

/*
 * The whole per-cluster-pair path -- replaces utils_match.hist_icp (utils_match.py:138-157).
 *   src, dst [P,N,4] -> out_pose [P,16] column-convention 4x4 transforms  p' = T[:3,:3] p + T[:3,3]
 *   out_init [P,16] (may be NULL) the histogram initialisation in the swapped frame; out_batch [2] (may be NULL)
 *   workspace  icpf_workspace_bytes(P, N, len[0], len[1], len[2])
 */
int icpf_hist_icp_f32(const float* src, const float* dst, int32_t P, int32_t N, const icpf_hist_bins* bins,
                      const icpf_params* params, float* out_pose, float* out_init, int32_t* out_batch,
                      void* workspace, size_t workspace_bytes, void* stream);

/*
 * ---------------------------------------------------------------------------------------------------------------
 * Scan-level callers of the path.  A scan is `points` [n, stride >= 3] fp32 rows
 * (x, y, z, ...) plus one fp32 label per point, as the reference holds them (main.py:205-208): cluster labels are the
 * non-negative integers, ground / unclustered points carry -1e8 / -1.
 * ---------------------------------------------------------------------------------------------------------------
 */

/*
 * Cluster index of one scan -- replaces the per-use boolean masks `points[labels == l]` of sanity_check
 * (utils_check.py:24-25) and match_pairs (utils_match.py:85-86) by one stable counting sort.
 *   labels [n] fp32; a label l is indexed when it is an integer with 0 <= l < n_labels
 *   out_order   [n] int32    point rows grouped by label, scan order kept inside a label (first out_offsets[n_labels] valid)
 *   out_offsets [n_labels+1] int32   rows of label l = out_order[out_offsets[l] .. out_offsets[l+1])
 *   out_stats   [n_labels,8] fp32    {mean x, mean y, mean z (fp64 sums rounded once), the three |max - min| extents
 *                                     sorted ascending (get_bbox_tensor, utils_helper.py:166-170), 0, 0}
 *   workspace   icpf_cluster_index_workspace_bytes(n, n_labels)
 */
size_t icpf_cluster_index_workspace_bytes(int32_t n_points, int32_t n_labels);
int icpf_cluster_index_f32(const float* points, int32_t point_stride, const float* labels, int32_t n_points,
                           int32_t n_labels, int32_t* out_order, int32_t* out_offsets, float* out_stats,
                           void* workspace, size_t workspace_bytes, void* stream);

/*
 * Candidate-pair filter -- replaces utils_check.sanity_check (utils_check.py:21-49), a Python loop with two scan-wide
 * masks and several host syncs per candidate pair, by one launch on the statistics of the two cluster indices.
 *   pairs [P,2] int64 (src label, dst label) as the reference's `pairs` tensor (utils_match.py:32,50)
 *   a pair is kept when both labels are >= 0, min(len) >= min_cluster_size, |mean_dst - mean_src|_xy <= translation_frame
 *   and, per sorted bounding-box extent, min >= thres_box * max   (all comparisons in fp32 like the reference)
 *   out_keep [P] int32 0/1; out_pairs [P,2] int64 the kept pairs in input order; out_count [1] int32 their number
 */
int icpf_sanity_check_f32(const int32_t* src_offsets, const float* src_stats, int32_t n_src_labels,
                          const int32_t* dst_offsets, const float* dst_stats, int32_t n_dst_labels,
                          const int64_t* pairs, int32_t P, int32_t min_cluster_size, double translation_frame,
                          double thres_box, int32_t* out_keep, int64_t* out_pairs, int32_t* out_count, void* stream);

/*
 * The same filter over ALL pairs of two label lists -- the dynamic stage of match_pcds (utils_match.py:43-51: every
 * unmatched src cluster against every unmatched dst cluster; the reference materialises the cross product with
 * repeat_interleave / repeat / stack) without building the candidate tensor.
 *   lists [n_src_list + n_dst_list] int64: the src labels followed by the dst labels; candidate q is
 *   (lists[q / n_dst_list], lists[n_src_list + q % n_dst_list]), i.e. src-major like the reference
 *   out_pairs [n_src_list * n_dst_list, 2] int64 (kept pairs in that order), out_count [1] int32
 */
int icpf_sanity_check_cross_f32(const int32_t* src_offsets, const float* src_stats, int32_t n_src_labels,
                                const int32_t* dst_offsets, const float* dst_stats, int32_t n_dst_labels,
                                const int64_t* lists, int32_t n_src_list, int32_t n_dst_list, int32_t min_cluster_size,
                                double translation_frame, double thres_box, int64_t* out_pairs, int32_t* out_count,
                                void* stream);

/*
 * Selection -- replaces the rejection loop and the selection of match_pairs (utils_match.py:70-75, 94-135: a Python loop
 * over the pairs with two torch.nonzero calls and five scattered stores each, then match_segments_descend) by one launch.
 *   pairs [P,2] int64; src_labels [n_src] / dst_labels [n_dst] int64, SORTED (torch.unique of the scan labels)
 *   errors, inliers, ratios, ious [P,2] fp32 and accept [P] int32 as written by icpf_match_eval_f32; transforms [P,16]
 *   Per src label: among its accepted pairs the dst label of least min(error) (ties: lowest dst position, as argmin
 *   over the reference's [n_src, n_dst] matrix; a NaN error wins the arg-min and then fails the threshold, as in torch),
 *   kept if that error < thres_error.
 *   out_rows [n_src,10] fp32: (src label, dst label, error x2, inliers x2, ratios x2, ious x2) of the kept matches in
 *   src-label order; out_transforms [n_src,16]; out_src_left [n_src] / out_dst_left [n_dst] int64: the labels without a
 *   kept match, sorted (the candidates of the dynamic stage); out_counts [3] int32: rows, src labels left, dst labels left.
 *   workspace: icpf_match_select_workspace_bytes(n_src, n_dst) bytes, 8-byte aligned.
 */
size_t icpf_match_select_workspace_bytes(int32_t n_src, int32_t n_dst);
int icpf_match_select_f32(const int64_t* pairs, int32_t P, const int64_t* src_labels, int32_t n_src,
                          const int64_t* dst_labels, int32_t n_dst, const float* errors, const float* inliers,
                          const float* ratios, const float* ious, const int32_t* accept, const float* transforms,
                          double thres_error, float* out_rows, float* out_transforms, int64_t* out_src_left,
                          int64_t* out_dst_left, int32_t* out_counts, void* workspace, size_t workspace_bytes,
                          void* stream);

/*
 * Padded pair batch -- replaces the gather loop of match_pairs (utils_match.py:81-91) and pad_segment
 * (utils_helper.py:185-196): out_src / out_dst [P, max_points, 4] rows (x, y, z, 1) in scan order, then
 * (1e8, 1e8, 1e8, 0).  Clusters with more than max_points rows: the reference keeps torch.randperm(len)[:max_points];
 * the caller draws that permutation (same call, same order, so the same RNG stream) and passes it --
 *   sample_rows     int32 positions inside the cluster, max_points per sampled cluster (may be NULL)
 *   sample_offsets  [P,2] int64 start of the (pair, side) sample inside sample_rows, -1 = none (may be NULL)
 * -- a larger cluster without a sample keeps its first max_points rows.
 */
int icpf_gather_pairs_f32(const float* src_points, int32_t src_stride, const int32_t* src_order,
                          const int32_t* src_offsets, int32_t n_src_labels, const float* dst_points,
                          int32_t dst_stride, const int32_t* dst_order, const int32_t* dst_offsets,
                          int32_t n_dst_labels, const int64_t* pairs, int32_t P, int32_t max_points,
                          const int32_t* sample_rows, const int64_t* sample_offsets, float* out_src, float* out_dst,
                          void* stream);

/*
 * Scene flow from the matched transforms -- replaces utils_flow.flow_estimation_torch (utils_flow.py:57-69):
 *   flow_i = (T_k * pose) p_i - p_i  with k the (last) row whose pair_labels[k] equals labels[i], T = identity otherwise.
 *   pair_labels  fp32, row k at pair_labels[k * pair_stride]  (column 0 of the [K,10] `pairs` rows: pair_stride = 10)
 *   transforms [K,16] row-major 4x4; pose [16] DEVICE row-major 4x4 (NULL = identity); out_flow [n,3]
 */
int icpf_flow_f32(const float* points, int32_t point_stride, const float* labels, int32_t n_points,
                  const float* pair_labels, int32_t pair_stride, const float* transforms, int32_t K, const float* pose,
                  float* out_flow, void* stream);

/*
 * Per-call extensions of the ICP loop.  Everything a call needs is an explicit argument: the library keeps NO state
 * between calls (no globals, no thread-locals), so calls from any number of host threads / streams are independent.
 *   peer_pose_dev / peer_world / peer_row0   fused all-gather of the transforms (multi-GPU, one process per GPU): the
 *       kernel epilogue also stores every pair's 4x4 (64 B) into row (peer_row0 + p) of the `[world * P, 16]` buffers of
 *       ALL ranks through peer-mapped pointers (DEVICE array of `peer_world` base pointers, e.g. torch symmetric memory
 *       `buffer_ptrs_dev`) instead of a separate NCCL all-gather; the caller runs a cross-rank barrier on the stream
 *       afterwards (rows are visible once every rank's kernel has ended).  NULL / 0 = off.
 *   start_event / stop_event   measurement hook (bench.py): cudaEvent_t handles recorded on `stream` immediately around
 *       the launch of the dominant kernel (icp_pairs_kernel, first pass).  NULL = off.  No effect on results.
 */
typedef struct icpf_icp_ext {
    void* const* peer_pose_dev;
    int32_t peer_world;
    int32_t peer_row0;
    void* start_event;
    void* stop_event;
} icpf_icp_ext;

/* icpf_icp_f32 with the extensions above (`ext` may be NULL: then exactly icpf_icp_f32). */
int icpf_icp_ex_f32(const float* src, const float* dst, const float* init_R, const float* init_T, int32_t P, int32_t N,
                    const icpf_params* params, float* out_R, float* out_T, float* out_rmse, float* out_pose,
                    int32_t* out_iters, uint32_t* out_conv, int32_t* out_batch, void* workspace, size_t workspace_bytes,
                    void* stream, const icpf_icp_ext* ext);

/*
 * The all-gather of the transforms as ONE small launch (multi-GPU): copies this rank's contiguous block local_pose
 * [P,16] into rows [row0, row0 + P) of the `[world * P, 16]` buffer of every rank (peer-mapped pointers as above) with
 * 16-byte stores.  The alternative to the fused epilogue stores for callers whose transforms are final only after a later
 * kernel (apply_icp / hist_icp).  The caller runs the cross-rank barrier afterwards.
 */
int icpf_peer_push_f32(const float* local_pose, void* const* peer_pose_dev, int32_t world, int32_t row0, int32_t P,
                       void* stream);

/*
 * Compact input format -> the padded layout.  The reference's padded batch ships 16 bytes per row for 12 bytes of
 * information and pads every cluster to max_points (utils_helper.py:185-196); a host that feeds the engine over PCIe can
 * send the valid rows only:
 *   rows     [total, 3] fp32 xyz, the valid rows of cluster 0, then of cluster 1, ... (DEVICE, after the H2D copy)
 *   offsets  [B + 1] int32, rows of cluster b = rows[offsets[b] .. offsets[b+1])   (at most N rows each)
 *   out      [B, N, 4] fp32: (x, y, z, 1) for the valid rows, then (1e8, 1e8, 1e8, 0) -- bit for bit what pad_segment
 *            builds from the same rows.
 */
int icpf_expand_rows_f32(const float* rows, const int32_t* offsets, int32_t B, int32_t N, float* out, void* stream);

/*
 * Clustering of one scan -- replaces utils_cluster.cluster_dbscan's call into Open3D (utils_cluster.py:32-38:
 * `pcd.cluster_dbscan(eps=args.epsilon, min_points=args.min_cluster_size)`), SURVEY.md section 8 row f4.
 *   points [n, stride >= 3] fp32 (x, y, z, ...) DEVICE; rows with a non-finite coordinate are noise
 *   out_labels [n] int32: -1 noise, else the cluster number.  A core point has >= min_points points (itself included)
 *   within eps (fp64 distance of the fp32 coordinates, dx*dx + dy*dy + dz*dz <= eps*eps); clusters are the connected
 *   components of the core points, numbered by their lowest core-point index (the order the sequential algorithm meets
 *   them); a border point joins the lowest-numbered cluster that has a core point within eps.  The labels are therefore
 *   the sequential algorithm's (sklearn.cluster.DBSCAN / Open3D) on the same points.
 *   out_num_clusters [1] int32 (may be NULL); workspace icpf_dbscan_workspace_bytes(n) (a cell table of min(64 n, 8 M) int32 + 21 B per point)
 * The reference's top-`num_clusters` selection (utils_cluster.py:40-46) is host logic on the label counts
 * (icp_flow_b200.cluster.cluster_dbscan).
 */
size_t icpf_dbscan_workspace_bytes(int32_t n_points);
int icpf_dbscan_f32(const float* points, int32_t point_stride, int32_t n_points, double eps, int32_t min_points,
                    int32_t* out_labels, int32_t* out_num_clusters, void* workspace, size_t workspace_bytes, void* stream);

/*
 * HDBSCAN -- replaces utils_cluster.cluster_hdbscan (utils_cluster.py:10-29: hdbscan.HDBSCAN(min_cluster_size,
 * min_samples=None, alpha=1, metric='euclidean').fit(points[:, 0:3]).labels_), the clustering every script of the
 * reference selects (--if_hdbscan).  Two calls:
 *
 * icpf_hdbscan_mst_f32 (device, stream-ordered, n_points + 3 launches): core distances (distance to the min_samples-th
 *   nearest point, the point itself included; fp64 on the fp32 coordinates) and the minimum spanning tree of the mutual-
 *   reachability graph max(core_a, core_b, d(a,b)) by Prim's algorithm from point 0 -- the edges in the ORDER Prim adds
 *   them (strict `<` relaxation, first minimum in index order), which is what scikit-learn's port of the package builds
 *   (sklearn/cluster/_hdbscan/_linkage.pyx) and what fixes the dendrogram among edges of equal weight.
 *   prim_order = 0: the tree that is unique under the strict edge order (weight, min(a,b), max(a,b)), built by Boruvka
 *   rounds (<= log2 n rounds of n^2 candidate edges instead of n - 1 dependent steps: ~10x faster); its edges come in no
 *   particular order, to be sorted by that order (icpf_hdbscan_labels_host with presorted = 0).  Same weights as the
 *   oracle's tree, not its order among equal weights: a few labels per scan may differ where the oracle's own result
 *   depends on that order.  This mode synchronises the stream once per round (the round count depends on the data).
 *   points [n, point_stride >= 3] fp32, all finite; out_core [n] f64; out_edge_src / out_edge_dst [n-1] int32,
 *   out_edge_w [n-1] f64; min_samples <= 64; workspace icpf_hdbscan_workspace_bytes(n) bytes, 256-byte aligned.
 *
 * icpf_hdbscan_labels_host: HOST arrays.  The n - 1 edges (a, b, weight) of that tree -> single-linkage dendrogram ->
 *   condensed tree (min_cluster_size) -> excess-of-mass selection (the root is never a cluster) -> out_labels [n] int32,
 *   clusters numbered by their lowest point, -1 = noise.  presorted != 0: the edges are already in ascending weight
 *   order and are merged in exactly that order (the host wrapper sorts them with numpy's argsort, as the oracle does);
 *   otherwise they are sorted here by (weight, min(a,b), max(a,b)).
 */
size_t icpf_hdbscan_workspace_bytes(int32_t n_points);
int icpf_hdbscan_mst_f32(const float* points, int32_t point_stride, int32_t n_points, int32_t min_samples,
                         int32_t prim_order, double* out_core, int32_t* out_edge_src, int32_t* out_edge_dst,
                         double* out_edge_w, void* workspace, size_t workspace_bytes, void* stream);
int icpf_hdbscan_labels_host(const int32_t* edge_a, const int32_t* edge_b, const double* edge_w, int32_t n_points,
                             int32_t min_cluster_size, int32_t presorted, int32_t* out_labels);

/* CPU-callable test hook: the closed-form 3x3 Kabsch rotation used inside the kernels, evaluated on the host
 * for `n` row-major cross-covariance matrices H (n*9 floats) -> R (n*9 floats).  Not part of the data path. */
void icpf_host_kabsch(const float* H, int32_t n, float* R);
/* Same, but the n matrices are solved as one ICP run would: each solve warm-starts from the previous one. */
void icpf_host_kabsch_sequence(const float* H, int32_t n, float* R);

#ifdef __cplusplus
}
#endif
#endif /* ICPFLOW_B200_H */
