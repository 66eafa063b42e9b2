// icpf_api.cu -- the extern "C" boundary (include/icpflow_b200.h): argument checks, then stream-ordered launches.
#include "icpf_internal.h"
#include "icpf_kabsch.h"

using namespace icpf;

namespace {
inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
}  // namespace

extern "C" {

int icpf_version(void) { return 105; }   // 102: icpf_icp_ex_f32, icpf_peer_push_f32, icpf_expand_rows_f32; 103: icpf_dbscan_f32

const char* icpf_error_string(int code) {
    switch (code) {
        case ICPF_OK: return "ok";
        case ICPF_E_NULL: return "required pointer is NULL";
        case ICPF_E_SHAPE: return "shape out of the supported range";
        case ICPF_E_PARAM: return "invalid parameter";
        case ICPF_E_ALIGN: return "pointer is not 16-byte aligned";
        case ICPF_E_WORKSPACE: return "workspace missing or too small";
        case ICPF_E_UNSUPPORTED: return "configuration not supported by this build";
        default: break;
    }
    if (code > 0) return cudaGetErrorString(static_cast<cudaError_t>(code));
    return "unknown error";
}

void icpf_default_params(icpf_params* p) {
    if (!p) return;
    p->thres_dist = 0.1;
    p->max_iterations = 100;
    p->relative_rmse_thr = 1e-6f;
    p->early_exit = 1;
    p->batch_stop = 1;
    p->nn_mode = 0;
    p->reserved[0] = p->reserved[1] = 0;
}

size_t icpf_workspace_bytes(int32_t P, int32_t N, int32_t lx, int32_t ly, int32_t lz) {
    if (P < 0 || lx < 0 || ly < 0 || lz < 0) return 0;
    return path_workspace_bytes(P, N, lx, ly, lz);
}

int icpf_icp_f32(const float* src, const float* dst, const float* init_R, const float* init_T, int32_t P, int32_t N,
                 const icpf_params* params, float* out_R,
                 float* out_T, float* out_rmse, float* out_pose, int32_t* out_iters, uint32_t* out_conv,
                 int32_t* out_batch, void* workspace, size_t workspace_bytes, void* stream) {
    return icpf_icp_ex_f32(src, dst, init_R, init_T, P, N, params, out_R, out_T, out_rmse, out_pose, out_iters, out_conv,
                           out_batch, workspace, workspace_bytes, stream, nullptr);
}

int icpf_icp_ex_f32(const float* src, const float* dst, const float* init_R, const float* init_T, int32_t P, int32_t N,
                    const icpf_params* params, float* out_R, float* out_T, float* out_rmse, float* out_pose,
                    int32_t* out_iters, uint32_t* out_conv, int32_t* out_batch, void* workspace, size_t workspace_bytes,
                    void* stream, const icpf_icp_ext* ext) {
    if (!params) return ICPF_E_NULL;
    if (ext != nullptr) {
        if (ext->peer_world < 0 || ext->peer_row0 < 0 || ext->peer_world > 64) return ICPF_E_PARAM;
        if (ext->peer_world > 0 && ext->peer_pose_dev == nullptr) return ICPF_E_NULL;
        if ((ext->start_event == nullptr) != (ext->stop_event == nullptr)) return ICPF_E_NULL;
    }
    if (P < 0 || N <= 0) return ICPF_E_SHAPE;
    if (P == 0) return ICPF_OK;
    if (!src || !dst || !out_R || !out_T) return ICPF_E_NULL;
    if (!aligned16(src) || !aligned16(dst)) return ICPF_E_ALIGN;
    if (params->max_iterations < 1 || params->max_iterations > ICPF_MAX_ITERATIONS) return ICPF_E_PARAM;
    if (!(params->thres_dist > 0.0)) return ICPF_E_PARAM;
    if ((init_R == nullptr) != (init_T == nullptr)) return ICPF_E_NULL;
    return launch_icp(src, dst, init_R, init_T, nullptr, 0, P, N, *params, out_R, out_T, out_rmse, out_pose, out_iters, out_conv, out_batch, workspace,
                      workspace_bytes, static_cast<cudaStream_t>(stream), nullptr, ext);
}

int icpf_nn_f32(const float* src, const float* dst, int32_t B, int32_t Ns, int32_t Nd, int32_t src_stride,
                int32_t dst_stride, int64_t* out_idx, float* out_dist, void* stream) {
    if (B < 0 || Ns < 0 || Nd <= 0) return ICPF_E_SHAPE;
    if (src_stride < 3 || dst_stride < 3) return ICPF_E_PARAM;
    if (B == 0 || Ns == 0) return ICPF_OK;
    if (!src || !dst || !out_idx || !out_dist) return ICPF_E_NULL;
    return launch_nn(src, dst, B, Ns, Nd, src_stride, dst_stride, out_idx, out_dist, static_cast<cudaStream_t>(stream));
}

int icpf_transform_points_f32(const float* xyz, const float* pose, int32_t B, int32_t N, float* out, void* stream) {
    if (B < 0 || N < 0) return ICPF_E_SHAPE;
    if (B == 0 || N == 0) return ICPF_OK;
    if (!xyz || !pose || !out) return ICPF_E_NULL;
    if (!aligned16(xyz) || !aligned16(out)) return ICPF_E_ALIGN;
    return launch_transform_points(xyz, pose, B, N, out, static_cast<cudaStream_t>(stream));
}

int icpf_match_eval_f32(const float* src, const float* dst, const float* pose, int32_t P, int32_t N, double thres_dist,
                        float* out_errors, float* out_inliers, float* out_ratios, float* out_ious,
                        float* out_translations, float* out_rotations, const icpf_match_gates* gates,
                        int32_t* out_accept, void* stream) {
    if (P < 0 || N < 1) return ICPF_E_SHAPE;
    if (P == 0) return ICPF_OK;
    if (!src || !dst || !pose || !out_errors || !out_inliers || !out_ratios || !out_ious || !out_translations ||
        !out_rotations)
        return ICPF_E_NULL;
    if (!aligned16(src) || !aligned16(dst)) return ICPF_E_ALIGN;
    if ((gates == nullptr) != (out_accept == nullptr)) return ICPF_E_NULL;
    if (!(thres_dist > 0.0)) return ICPF_E_PARAM;
    return launch_match_eval(src, dst, pose, P, N, (float)thres_dist, out_errors, out_inliers, out_ratios, out_ious,
                             out_translations, out_rotations, gates, out_accept, static_cast<cudaStream_t>(stream));
}

static int check_params(const icpf_params* params) {
    if (!params) return ICPF_E_NULL;
    if (params->max_iterations < 1 || params->max_iterations > ICPF_MAX_ITERATIONS) return ICPF_E_PARAM;
    if (!(params->thres_dist > 0.0)) return ICPF_E_PARAM;
    return ICPF_OK;
}

static int check_bins(const icpf_hist_bins* b) {
    if (!b) return ICPF_E_NULL;
    if (!b->bins_x || !b->bins_y || !b->bins_z) return ICPF_E_NULL;
    for (int k = 0; k < 3; ++k) {
        if (b->len[k] < 1 || b->len[k] > 4096) return ICPF_E_SHAPE;
        if (!(b->max[k] > b->min[k])) return ICPF_E_PARAM;
    }
    return ICPF_OK;
}

int icpf_hist_votes_f32(const float* X, const float* Y, int32_t B, int32_t NX, int32_t NY, const float* min_xyz,
                        const float* max_xyz, const int32_t* len_xyz, float* bins, void* stream) {
    if (B < 0 || NX < 0 || NY < 0 || B > 65535) return ICPF_E_SHAPE;
    if (!min_xyz || !max_xyz || !len_xyz) return ICPF_E_NULL;
    for (int k = 0; k < 3; ++k) {
        if (len_xyz[k] < 1 || len_xyz[k] > 4096) return ICPF_E_SHAPE;
        if (!(max_xyz[k] > min_xyz[k])) return ICPF_E_PARAM;
    }
    if (B == 0) return ICPF_OK;
    if (!X || !Y || !bins) return ICPF_E_NULL;
    if (!aligned16(X) || !aligned16(Y)) return ICPF_E_ALIGN;
    return launch_hist_votes(X, Y, B, NX, NY, min_xyz, max_xyz, len_xyz, bins, 0, nullptr,
                             static_cast<cudaStream_t>(stream));
}

int icpf_hist_init_f32(const float* src, const float* dst, int32_t P, int32_t N, const icpf_hist_bins* bins,
                       int32_t auto_swap, float* out_pose, int32_t* out_cand, float* out_votes, float* out_scores,
                       int32_t* out_which, void* workspace, size_t workspace_bytes, void* stream) {
    const int rc = check_bins(bins);
    if (rc != ICPF_OK) return rc;
    if (P < 0 || N <= 0 || P > 65535 * 64) return ICPF_E_SHAPE;
    if (P == 0) return ICPF_OK;
    if (!src || !dst || !out_pose) return ICPF_E_NULL;
    if (!aligned16(src) || !aligned16(dst)) return ICPF_E_ALIGN;
    return launch_hist_init(src, dst, P, N, *bins, auto_swap, out_pose, out_cand, out_votes, out_scores, out_which,
                            workspace, workspace_bytes, static_cast<cudaStream_t>(stream));
}

int icpf_apply_icp_f32(const float* src, const float* dst, const float* init_pose, int32_t P, int32_t N,
                       const icpf_params* params, int32_t auto_swap, float* out_pose, float* out_err,
                       int32_t* out_flags, int32_t* out_batch, void* workspace, size_t workspace_bytes, void* stream) {
    const int rc = check_params(params);
    if (rc != ICPF_OK) return rc;
    if (P < 0 || N <= 0) return ICPF_E_SHAPE;
    if (P == 0) return ICPF_OK;
    if (!src || !dst || !init_pose || !out_pose) return ICPF_E_NULL;
    if (!aligned16(src) || !aligned16(dst)) return ICPF_E_ALIGN;
    return launch_apply_icp(src, dst, init_pose, P, N, *params, auto_swap, out_pose, out_err, out_flags, out_batch,
                            workspace, workspace_bytes, static_cast<cudaStream_t>(stream));
}

int icpf_apply_icp_phase_f32(const float* src, const float* dst, const float* init_pose, int32_t P, int32_t N,
                             const icpf_params* params, int32_t auto_swap, int32_t phase, int32_t batch_iterations,
                             int32_t batch_converged, uint32_t* out_and, float* out_pose, float* out_err,
                             int32_t* out_flags, int32_t* out_batch, void* workspace, size_t workspace_bytes,
                             void* stream) {
    const int rc = check_params(params);
    if (rc != ICPF_OK) return rc;
    if (P < 0 || N <= 0) return ICPF_E_SHAPE;
    if (phase < 0 || phase > 2 || !params->batch_stop) return ICPF_E_PARAM;
    if (phase == 2 && (batch_iterations < 1 || batch_iterations > params->max_iterations)) return ICPF_E_PARAM;
    if (P == 0) return ICPF_OK;
    if (!src || !dst || !init_pose) return ICPF_E_NULL;
    if (phase < 2 ? !out_and : !out_pose) return ICPF_E_NULL;
    if (!aligned16(src) || !aligned16(dst)) return ICPF_E_ALIGN;
    IcpPhase ph{phase, batch_iterations, batch_converged != 0 ? 1 : 0, out_and};
    return launch_apply_icp(src, dst, init_pose, P, N, *params, auto_swap, out_pose, out_err, out_flags, out_batch,
                            workspace, workspace_bytes, static_cast<cudaStream_t>(stream), &ph);
}

int icpf_hist_icp_f32(const float* src, const float* dst, int32_t P, int32_t N, const icpf_hist_bins* bins,
                      const icpf_params* params, float* out_pose, float* out_init, int32_t* out_batch,
                      void* workspace, size_t workspace_bytes, void* stream) {
    int rc = check_params(params);
    if (rc != ICPF_OK) return rc;
    rc = check_bins(bins);
    if (rc != ICPF_OK) return rc;
    if (P < 0 || N <= 0) return ICPF_E_SHAPE;
    if (P == 0) return ICPF_OK;
    if (!src || !dst || !out_pose) return ICPF_E_NULL;
    if (!aligned16(src) || !aligned16(dst)) return ICPF_E_ALIGN;
    return launch_hist_icp(src, dst, P, N, *bins, *params, out_pose, out_init, out_batch, workspace, workspace_bytes,
                           static_cast<cudaStream_t>(stream));
}

size_t icpf_cluster_index_workspace_bytes(int32_t n_points, int32_t n_labels) {
    if (n_points < 0 || n_labels < 1 || n_labels > (1 << 20)) return 0;
    return cluster_index_workspace_bytes(n_points, n_labels);
}

int icpf_cluster_index_f32(const float* points, int32_t point_stride, const float* labels, int32_t n_points,
                           int32_t n_labels, int32_t* out_order, int32_t* out_offsets, float* out_stats,
                           void* workspace, size_t workspace_bytes, void* stream) {
    if (n_points < 0 || n_labels < 1 || n_labels > (1 << 20)) return ICPF_E_SHAPE;
    if (point_stride < 3) return ICPF_E_PARAM;
    if (!out_offsets || !out_stats) return ICPF_E_NULL;
    if (n_points > 0 && (!points || !labels || !out_order)) return ICPF_E_NULL;
    return launch_cluster_index(points, point_stride, labels, n_points, n_labels, out_order, out_offsets, out_stats,
                                workspace, workspace_bytes, static_cast<cudaStream_t>(stream));
}

int icpf_sanity_check_f32(const int32_t* src_offsets, const float* src_stats, int32_t n_src_labels,
                          const int32_t* dst_offsets, const float* dst_stats, int32_t n_dst_labels,
                          const int64_t* pairs, int32_t P, int32_t min_cluster_size, double translation_frame,
                          double thres_box, int32_t* out_keep, int64_t* out_pairs, int32_t* out_count, void* stream) {
    if (P < 0 || n_src_labels < 1 || n_dst_labels < 1) return ICPF_E_SHAPE;
    if (!src_offsets || !src_stats || !dst_offsets || !dst_stats || !out_count) return ICPF_E_NULL;
    if (P > 0 && (!pairs || !out_keep || !out_pairs)) return ICPF_E_NULL;
    return launch_sanity_check(src_offsets, src_stats, n_src_labels, dst_offsets, dst_stats, n_dst_labels, pairs, P,
                               min_cluster_size, (float)translation_frame, (float)thres_box, out_keep, out_pairs,
                               out_count, static_cast<cudaStream_t>(stream));
}

int icpf_sanity_check_cross_f32(const int32_t* src_offsets, const float* src_stats, int32_t n_src_labels,
                                const int32_t* dst_offsets, const float* dst_stats, int32_t n_dst_labels,
                                const int64_t* lists, int32_t n_src_list, int32_t n_dst_list, int32_t min_cluster_size,
                                double translation_frame, double thres_box, int64_t* out_pairs, int32_t* out_count,
                                void* stream) {
    if (n_src_list < 0 || n_dst_list < 0 || n_src_labels < 1 || n_dst_labels < 1) return ICPF_E_SHAPE;
    if ((long long)n_src_list * n_dst_list > 0x7fffffffLL) return ICPF_E_SHAPE;
    if (!src_offsets || !src_stats || !dst_offsets || !dst_stats || !out_count) return ICPF_E_NULL;
    const int P = n_src_list * n_dst_list;
    if (P > 0 && (!lists || !out_pairs)) return ICPF_E_NULL;
    return launch_sanity_check(src_offsets, src_stats, n_src_labels, dst_offsets, dst_stats, n_dst_labels, lists, P,
                               min_cluster_size, (float)translation_frame, (float)thres_box, nullptr, out_pairs,
                               out_count, static_cast<cudaStream_t>(stream), n_dst_list > 0 ? n_dst_list : 1);
}

int icpf_hdbscan_labels_host(const int32_t* edge_a, const int32_t* edge_b, const double* edge_w, int32_t n_points,
                             int32_t min_cluster_size, int32_t presorted, int32_t* out_labels) {
    if (n_points < 0 || min_cluster_size < 2) return ICPF_E_PARAM;
    if (n_points == 0) return ICPF_OK;
    if (!out_labels || (n_points > 1 && (!edge_a || !edge_b || !edge_w))) return ICPF_E_NULL;
    return hdbscan_labels_host(edge_a, edge_b, edge_w, n_points, min_cluster_size, presorted, out_labels);
}

size_t icpf_hdbscan_workspace_bytes(int32_t n_points) { return n_points > 0 ? hdbscan_workspace_bytes(n_points) : 0; }

int icpf_hdbscan_mst_f32(const float* points, int32_t point_stride, int32_t n_points, int32_t min_samples,
                         int32_t prim_order, double* out_core, int32_t* out_edge_src, int32_t* out_edge_dst,
                         double* out_edge_w, void* workspace, size_t workspace_bytes, void* stream) {
    if (n_points < 0 || point_stride < 3) return ICPF_E_SHAPE;
    if (min_samples < 1) return ICPF_E_PARAM;
    if (n_points == 0) return ICPF_OK;
    if (!points || !out_core) return ICPF_E_NULL;
    if (n_points > 1 && (!out_edge_src || !out_edge_dst || !out_edge_w)) return ICPF_E_NULL;
    if (workspace == nullptr || workspace_bytes < hdbscan_workspace_bytes(n_points)) return ICPF_E_WORKSPACE;
    if ((reinterpret_cast<uintptr_t>(workspace) & 255u) != 0) return ICPF_E_ALIGN;
    return launch_hdbscan_mst(points, point_stride, n_points, min_samples, prim_order, out_core, out_edge_src, out_edge_dst,
                              out_edge_w, workspace, static_cast<cudaStream_t>(stream));
}

size_t icpf_match_select_workspace_bytes(int32_t n_src, int32_t n_dst) {
    if (n_src < 0 || n_dst < 0) return 0;
    return match_select_workspace_bytes(n_src, n_dst);
}

int icpf_match_select_f32(const int64_t* pairs, int32_t P, const int64_t* src_labels, int32_t n_src,
                          const int64_t* dst_labels, int32_t n_dst, const float* errors, const float* inliers,
                          const float* ratios, const float* ious, const int32_t* accept, const float* transforms,
                          double thres_error, float* out_rows, float* out_transforms, int64_t* out_src_left,
                          int64_t* out_dst_left, int32_t* out_counts, void* workspace, size_t workspace_bytes,
                          void* stream) {
    if (P < 0 || n_src < 0 || n_dst < 0) return ICPF_E_SHAPE;
    if (!out_counts) return ICPF_E_NULL;
    if (P > 0 && (!pairs || !errors || !inliers || !ratios || !ious || !accept || !transforms)) return ICPF_E_NULL;
    if (n_src > 0 && (!src_labels || !out_rows || !out_transforms || !out_src_left)) return ICPF_E_NULL;
    if (n_dst > 0 && (!dst_labels || !out_dst_left)) return ICPF_E_NULL;
    if (workspace == nullptr || workspace_bytes < match_select_workspace_bytes(n_src, n_dst)) return ICPF_E_WORKSPACE;
    if ((reinterpret_cast<uintptr_t>(workspace) & 7u) != 0) return ICPF_E_ALIGN;
    return launch_match_select(pairs, P, src_labels, n_src, dst_labels, n_dst, errors, inliers, ratios, ious, accept,
                               transforms, (float)thres_error, out_rows, out_transforms, out_src_left, out_dst_left,
                               out_counts, workspace, static_cast<cudaStream_t>(stream));
}

int icpf_gather_pairs_f32(const float* src_points, int32_t src_stride, const int32_t* src_order,
                          const int32_t* src_offsets, int32_t n_src_labels, const float* dst_points,
                          int32_t dst_stride, const int32_t* dst_order, const int32_t* dst_offsets,
                          int32_t n_dst_labels, const int64_t* pairs, int32_t P, int32_t max_points,
                          const int32_t* sample_rows, const int64_t* sample_offsets, float* out_src, float* out_dst,
                          void* stream) {
    if (P < 0 || max_points < 1 || n_src_labels < 1 || n_dst_labels < 1) return ICPF_E_SHAPE;
    if (src_stride < 3 || dst_stride < 3) return ICPF_E_PARAM;
    if (P == 0) return ICPF_OK;
    if (!src_points || !src_order || !src_offsets || !dst_points || !dst_order || !dst_offsets || !pairs || !out_src ||
        !out_dst)
        return ICPF_E_NULL;
    if ((sample_rows == nullptr) != (sample_offsets == nullptr)) return ICPF_E_NULL;
    if (!aligned16(out_src) || !aligned16(out_dst)) return ICPF_E_ALIGN;
    return launch_gather_pairs(src_points, src_stride, src_order, src_offsets, n_src_labels, dst_points, dst_stride,
                               dst_order, dst_offsets, n_dst_labels, pairs, P, max_points, sample_rows, sample_offsets,
                               out_src, out_dst, static_cast<cudaStream_t>(stream));
}

int icpf_flow_f32(const float* points, int32_t point_stride, const float* labels, int32_t n_points,
                  const float* pair_labels, int32_t pair_stride, const float* transforms, int32_t K, const float* pose,
                  float* out_flow, void* stream) {
    if (n_points < 0 || K < 0 || K > 65534) return ICPF_E_SHAPE;
    if (point_stride < 3 || (K > 0 && pair_stride < 1)) return ICPF_E_PARAM;
    if (n_points == 0) return ICPF_OK;
    if (!points || !labels || !out_flow) return ICPF_E_NULL;
    if (K > 0 && (!pair_labels || !transforms)) return ICPF_E_NULL;
    return launch_flow(points, point_stride, labels, n_points, pair_labels, pair_stride, transforms, K, pose, out_flow,
                       static_cast<cudaStream_t>(stream));
}

int icpf_peer_push_f32(const float* local_pose, void* const* peer_pose_dev, int32_t world, int32_t row0, int32_t P,
                       void* stream) {
    if (world < 0 || row0 < 0 || world > 64 || P < 0) return ICPF_E_PARAM;
    if (P == 0 || world == 0) return ICPF_OK;
    if (!local_pose || !peer_pose_dev) return ICPF_E_NULL;
    if (!aligned16(local_pose)) return ICPF_E_ALIGN;
    return launch_peer_push(local_pose, reinterpret_cast<float* const*>(peer_pose_dev), world, row0, P,
                            static_cast<cudaStream_t>(stream));
}

int icpf_expand_rows_f32(const float* rows, const int32_t* offsets, int32_t B, int32_t N, float* out, void* stream) {
    if (B < 0 || N <= 0) return ICPF_E_SHAPE;
    if (B == 0) return ICPF_OK;
    if (!rows || !offsets || !out) return ICPF_E_NULL;
    if (!aligned16(out)) return ICPF_E_ALIGN;
    return launch_expand_rows(rows, offsets, B, N, out, static_cast<cudaStream_t>(stream));
}

size_t icpf_dbscan_workspace_bytes(int32_t n_points) { return n_points < 0 ? 0 : dbscan_workspace_bytes(n_points); }

int icpf_dbscan_f32(const float* points, int32_t point_stride, int32_t n_points, double eps, int32_t min_points,
                    int32_t* out_labels, int32_t* out_num_clusters, void* workspace, size_t workspace_bytes, void* stream) {
    if (n_points < 0) return ICPF_E_SHAPE;
    if (point_stride < 3 || !(eps > 0.0) || min_points < 1) return ICPF_E_PARAM;
    if (n_points > 0 && (!points || !out_labels)) return ICPF_E_NULL;
    return launch_dbscan(points, point_stride, n_points, eps, min_points, out_labels, out_num_clusters, workspace,
                         workspace_bytes, static_cast<cudaStream_t>(stream));
}

void icpf_host_kabsch_sequence(const float* H, int32_t n, float* R) {
    KabschState st;
    st.warm = false;
    for (int32_t i = 0; i < n; ++i) {
        float h[9];
        for (int k = 0; k < 9; ++k) h[k] = H[9 * i + k];
        const Rot3 r = kabsch_rotation(h, &st);
        for (int k = 0; k < 9; ++k) R[9 * i + k] = r.r[k];
    }
}

void icpf_host_kabsch(const float* H, int32_t n, float* R) {
    for (int32_t i = 0; i < n; ++i) {
        float h[9];
        for (int k = 0; k < 9; ++k) h[k] = H[9 * i + k];
        const Rot3 r = kabsch_rotation(h);
        for (int k = 0; k < 9; ++k) R[9 * i + k] = r.r[k];
    }
}

}  // extern "C"
