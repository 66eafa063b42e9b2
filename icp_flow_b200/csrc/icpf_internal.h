// icpf_internal.h -- declarations shared between the translation units of libicpflow_b200.so (not installed).
#pragma once

#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include "../../include/icpflow_b200.h"

namespace icpf {

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// iters [P] int32 | conv [P,4] uint32 | batch [2] int32 (256 B slot) | stats [P,2] int32 {full searches, cache refreshes}
inline size_t icp_ws_off_conv(int P) { return align_up((size_t)P * 4, 256); }
inline size_t icp_ws_off_batch(int P) { return icp_ws_off_conv(P) + align_up((size_t)P * 16, 256); }
inline size_t icp_ws_off_stats(int P) { return icp_ws_off_batch(P) + 256; }
inline size_t icp_workspace_bytes(int P) { return icp_ws_off_stats(P) + align_up((size_t)P * 8, 256); }

int launch_icp(const float* src, const float* dst, const float* init_R, const float* init_T, int P, int N,
               const icpf_params& prm, float* out_R, float* out_T,
               float* out_rmse, float* out_pose, int* out_iters, uint32_t* out_conv, int* out_batch, void* workspace,
               size_t workspace_bytes, cudaStream_t stream);

void set_profile_events(cudaEvent_t start, cudaEvent_t stop);

int launch_nn(const float* src, const float* dst, int B, int Ns, int Nd, int src_stride, int dst_stride,
              int64_t* out_idx, float* out_dist, cudaStream_t stream);

int launch_transform_points(const float* xyz, const float* pose, int B, int N, float* out, cudaStream_t stream);

}  // namespace icpf
