// icpf_api.cu -- the extern "C" boundary (include/icpflow_b200.h): argument checks, then stream-ordered launches.
#include "icpf_internal.h"
#include "icpf_kabsch.h"

using namespace icpf;

namespace {
inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
}  // namespace

extern "C" {

int icpf_version(void) { return 100; }

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
    (void)N; (void)lx; (void)ly; (void)lz;
    if (P < 0) return 0;
    return icp_workspace_bytes(P);
}

int icpf_icp_f32(const float* src, const float* dst, const float* init_R, const float* init_T, int32_t P, int32_t N,
                 const icpf_params* params, float* out_R,
                 float* out_T, float* out_rmse, float* out_pose, int32_t* out_iters, uint32_t* out_conv,
                 int32_t* out_batch, void* workspace, size_t workspace_bytes, void* stream) {
    if (!params) return ICPF_E_NULL;
    if (P < 0 || N <= 0) return ICPF_E_SHAPE;
    if (P == 0) return ICPF_OK;
    if (!src || !dst || !out_R || !out_T) return ICPF_E_NULL;
    if (!aligned16(src) || !aligned16(dst)) return ICPF_E_ALIGN;
    if (params->max_iterations < 1 || params->max_iterations > ICPF_MAX_ITERATIONS) return ICPF_E_PARAM;
    if (!(params->thres_dist > 0.0)) return ICPF_E_PARAM;
    if ((init_R == nullptr) != (init_T == nullptr)) return ICPF_E_NULL;
    return launch_icp(src, dst, init_R, init_T, P, N, *params, out_R, out_T, out_rmse, out_pose, out_iters, out_conv, out_batch, workspace,
                      workspace_bytes, static_cast<cudaStream_t>(stream));
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

void icpf_profile_next_icp(void* start_event, void* stop_event) {
    set_profile_events(static_cast<cudaEvent_t>(start_event), static_cast<cudaEvent_t>(stop_event));
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
