// icpf_nn.cu -- stand-alone seams of the path that are not fused into the pair kernels:
//   * unbounded K=1 nearest neighbour over all rows   (utils_helper.nearest_neighbor_batch, utils_helper.py:20-30)
//   * homogeneous point transform keeping the flag     (utils_helper.transform_points_batch, utils_helper.py:76-87)
#include "icpf_internal.h"
#include "icpf_pair.cuh"

namespace icpf {

constexpr int kNnTile = 1024;   // candidate rows staged per shared-memory tile
constexpr int kNnQB = 4;        // query rows per thread

__global__ void __launch_bounds__(kThreads) nn_all_rows_kernel(const float* __restrict__ src,
                                                               const float* __restrict__ dst, int Ns, int Nd,
                                                               int src_stride, int dst_stride,
                                                               int64_t* __restrict__ out_idx,
                                                               float* __restrict__ out_dist) {
    __shared__ float4 tile[kNnTile];
    const int b = blockIdx.y;
    const float* s = src + (size_t)b * Ns * src_stride;
    const float* d = dst + (size_t)b * Nd * dst_stride;
    const int q0 = blockIdx.x * (kThreads * kNnQB) + threadIdx.x;
    float qx[kNnQB], qy[kNnQB], qz[kNnQB], best[kNnQB];
    int bidx[kNnQB];
#pragma unroll
    for (int k = 0; k < kNnQB; ++k) {
        const int q = q0 + k * kThreads;
        const bool in = q < Ns;
        qx[k] = in ? s[(size_t)q * src_stride + 0] : 0.f;
        qy[k] = in ? s[(size_t)q * src_stride + 1] : 0.f;
        qz[k] = in ? s[(size_t)q * src_stride + 2] : 0.f;
        best[k] = __int_as_float(0x7f800000);
        bidx[k] = 0;
    }
    for (int base = 0; base < Nd; base += kNnTile) {
        const int n = min(kNnTile, Nd - base);
        __syncthreads();
        for (int j = threadIdx.x; j < n; j += kThreads) {
            const float* r = d + (size_t)(base + j) * dst_stride;
            tile[j] = make_float4(r[0], r[1], r[2], 0.f);
        }
        __syncthreads();
#pragma unroll 4
        for (int j = 0; j < n; ++j) {
            const float4 c = tile[j];
#pragma unroll
            for (int k = 0; k < kNnQB; ++k) {
                const float dd = sqdist(qx[k], qy[k], qz[k], c.x, c.y, c.z);
                if (dd < best[k]) {   // strict: ties keep the lowest index
                    best[k] = dd;
                    bidx[k] = base + j;
                }
            }
        }
    }
#pragma unroll
    for (int k = 0; k < kNnQB; ++k) {
        const int q = q0 + k * kThreads;
        if (q < Ns) {
            out_idx[(size_t)b * Ns + q] = bidx[k];
            out_dist[(size_t)b * Ns + q] = sqrtf(best[k]);
        }
    }
}

int launch_nn(const float* src, const float* dst, int B, int Ns, int Nd, int src_stride, int dst_stride,
              int64_t* out_idx, float* out_dist, cudaStream_t stream) {
    if (B == 0 || Ns == 0) return ICPF_OK;
    // the batch runs on grid.y (<= 65535): larger batches go in slices
    for (int b0 = 0; b0 < B; b0 += 65535) {
        const int nb = (B - b0 < 65535) ? (B - b0) : 65535;
        dim3 grid((Ns + kThreads * kNnQB - 1) / (kThreads * kNnQB), nb);
        ICPF_LAUNCH(nn_all_rows_kernel, grid, kThreads, 0, stream)(src + (size_t)b0 * Ns * src_stride, dst + (size_t)b0 * Nd * dst_stride,
                                                          Ns, Nd, src_stride, dst_stride, out_idx + (size_t)b0 * Ns,
                                                          out_dist + (size_t)b0 * Ns);
    }
    return (int)cudaGetLastError();
}

// out[b,i,:3] = pose[b,:3,:3] xyz[b,i,:3] + pose[b,:3,3] evaluated as the 4-term dot product of the homogeneous
// row with pose^T (row 3 of pose is not assumed to be (0,0,0,1): only the first three output columns are kept).
__global__ void __launch_bounds__(256) transform_points_kernel(const float4* __restrict__ xyz,
                                                               const float* __restrict__ pose, int N,
                                                               float4* __restrict__ out) {
    const int b = blockIdx.y;
    __shared__ float m[12];
    if (threadIdx.x < 12) m[threadIdx.x] = pose[(size_t)b * 16 + threadIdx.x];
    __syncthreads();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const float4 p = xyz[(size_t)b * N + i];
    float4 o;
    o.x = fmaf(1.0f, m[3], fmaf(p.z, m[2], fmaf(p.y, m[1], p.x * m[0])));
    o.y = fmaf(1.0f, m[7], fmaf(p.z, m[6], fmaf(p.y, m[5], p.x * m[4])));
    o.z = fmaf(1.0f, m[11], fmaf(p.z, m[10], fmaf(p.y, m[9], p.x * m[8])));
    o.w = p.w;
    out[(size_t)b * N + i] = o;
}

int launch_transform_points(const float* xyz, const float* pose, int B, int N, float* out, cudaStream_t stream) {
    if (B == 0 || N == 0) return ICPF_OK;
    for (int b0 = 0; b0 < B; b0 += 65535) {        // the batch runs on grid.y (<= 65535)
        const int nb = (B - b0 < 65535) ? (B - b0) : 65535;
        dim3 grid((N + 255) / 256, nb);
        ICPF_LAUNCH(transform_points_kernel, grid, 256, 0, stream)(reinterpret_cast<const float4*>(xyz) + (size_t)b0 * N,
                                                         pose + (size_t)b0 * 16, N, reinterpret_cast<float4*>(out) + (size_t)b0 * N);
    }
    return (int)cudaGetLastError();
}

}  // namespace icpf
