// icpf_eval.cu -- registration quality metrics of a batch of cluster pairs.
//
// Replaces utils_match.match_eval (utils_match.py:159-213), the caller-side consumer of the transforms hist_icp returns
// (match_pairs, utils_match.py:93): the reference runs transform_points_batch, two knn_points passes and ~20 elementwise /
// reduction launches over [P,N] tensors; here one CTA per pair does everything in one launch and writes 14 floats.
//
// Row contract as everywhere in the engine: rows are (x,y,z,flag), the valid rows (flag > 0) form a prefix.  The
// reference searches all N rows of the other cloud; the padded rows sit at 1e8 and can only win when the other cloud has
// no valid row at all, so the candidates are the valid prefix -- or all rows when it is empty.
#include "icpf_internal.h"
#include "icpf_pair.cuh"
#include "icpf_gridnn.cuh"

namespace icpf {

constexpr int kEvTile = 1024;   // candidate rows staged per shared-memory tile
constexpr int kEvQB = 4;        // query rows per thread

struct EvalArgs {
    const float* src;
    const float* dst;
    const float* pose;
    int N;
    float thr;
    float* errors;
    float* inliers;
    float* ratios;
    float* ious;
    float* translations;
    float* rotations;
    float tau;              // thres_dist: only sizes the cells of the NN grids
    int* accept;            // optional: check_transformation verdict (utils_check.py:51-66)
    float gate_translation, gate_iou, gate_rot;
};

// Unbounded NN of the valid queries Q[0,n_q) among C[0,n_c): adds sqrt(d2) to err and (sqrt(d2) < thr) to inl.
// The moved cloud (pcd1 under the pose, utils_match.py:160) is the query side when MOVE_Q and the candidate side
// otherwise; rows are moved with the arithmetic of transform_points_batch while they are staged.  d2 is evaluated as
// (a-b)^2 summed x,y,z like knn_points, which is symmetric in its arguments, so both directions see the oracle's bits.
template <bool MOVE_Q>
__device__ __forceinline__ void eval_direction(const float (&m)[12], const float4* __restrict__ Q, int n_q,
                                               const float4* __restrict__ C, int n_c, float thr, float4* tile,
                                               float& err, float& inl) {
    for (int q0 = threadIdx.x; q0 - (int)threadIdx.x < n_q; q0 += kThreads * kEvQB) {
        float qx[kEvQB], qy[kEvQB], qz[kEvQB], best[kEvQB];
#pragma unroll
        for (int k = 0; k < kEvQB; ++k) {
            const int q = q0 + k * kThreads;
            float4 r = q < n_q ? Q[q] : make_float4(0.f, 0.f, 0.f, 0.f);
            if (MOVE_Q) r = transform_row(m, r);
            qx[k] = r.x; qy[k] = r.y; qz[k] = r.z;
            best[k] = __int_as_float(0x7f800000);
        }
        for (int base = 0; base < n_c; base += kEvTile) {
            const int n = min(kEvTile, n_c - base);
            __syncthreads();
            for (int j = threadIdx.x; j < n; j += kThreads) {
                float4 r = C[base + j];
                if (!MOVE_Q) r = transform_row(m, r);
                tile[j] = r;
            }
            __syncthreads();
#pragma unroll 4
            for (int j = 0; j < n; ++j) {
                const float4 c = tile[j];
#pragma unroll
                for (int k = 0; k < kEvQB; ++k) best[k] = fminf(best[k], sqdist(qx[k], qy[k], qz[k], c.x, c.y, c.z));
            }
        }
#pragma unroll
        for (int k = 0; k < kEvQB; ++k) {
            if (q0 + k * kThreads < n_q) {
                const float e = sqrtf(best[k]);
                err += e;
                inl += (e < thr) ? 1.f : 0.f;
            }
        }
    }
}

// Same sums through the NN grids (icpf_gridnn.cuh): the minimum over the inspected rows is the minimum of the full scan,
// and every thread accumulates its rows in the same increasing order, so both variants write the same bits.
template <bool MOVE_Q>
__device__ __forceinline__ void eval_direction_grid(const float (&m)[12], const float4* __restrict__ Q, int n_q,
                                                    const GridInfo& g, const float4* __restrict__ sorted,
                                                    const unsigned short* __restrict__ runs, int n_c, float thr,
                                                    float& err, float& inl) {
    for (int q0 = 0; q0 < n_q; q0 += kThreads) {              // (warp-uniform trip count: the far queries are scanned by the warp)
        const int q = q0 + threadIdx.x;
        const bool active = q < n_q;
        float4 r = active ? Q[q] : make_float4(0.f, 0.f, 0.f, 0.f);
        if (MOVE_Q) r = transform_row(m, r);
        const float e = sqrtf(nn_unbounded_grid_warp<false>(g, sorted, runs, n_c, active, r.x, r.y, r.z, r.x, r.y, r.z, 0.f,
                                                            0.f, 0.f));
        if (active) {
            err += e;
            inl += (e < thr) ? 1.f : 0.f;
        }
    }
}

// GRIDNN: dst and the MOVED src are counting-sorted into uniform grids in shared memory (2 (N + 257) * 16 B per pair);
// !GRIDNN (grids do not fit): full scans through a staged tile.
template <bool GRIDNN>
__global__ void __launch_bounds__(kThreads) match_eval_kernel(EvalArgs a) {
    const int p = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    __shared__ float4 tile[GRIDNN ? 1 : kEvTile];
    __shared__ float s_scratch[kWarps * 8];
    __shared__ double s_mean[kWarps][6];
    const float4* S = reinterpret_cast<const float4*>(a.src) + (size_t)p * a.N;
    const float4* D = reinterpret_cast<const float4*>(a.dst) + (size_t)p * a.N;
    float m[12];
#pragma unroll
    for (int i = 0; i < 12; ++i) m[i] = a.pose[(size_t)p * 16 + i];

    // valid counts and the two coordinate sums of pcd1 (as stored / moved); the sums are carried in fp64 and rounded once
    float cnt[2] = {0.f, 0.f};
    double sum[6] = {0., 0., 0., 0., 0., 0.};
    for (int q = tid; q < a.N; q += kThreads) {
        const float4 s = S[q];
        cnt[1] += (D[q].w > 0.f) ? 1.f : 0.f;
        if (s.w > 0.f) {
            cnt[0] += 1.f;
            const float4 t = transform_row(m, s);
            sum[0] += t.x; sum[1] += t.y; sum[2] += t.z;
            sum[3] += s.x; sum[4] += s.y; sum[5] += s.z;
        }
    }
    block_allreduce_sum<2, kWarps>(cnt, s_scratch);
#pragma unroll
    for (int i = 0; i < 6; ++i) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sum[i] += __shfl_down_sync(0xffffffffu, sum[i], o);
        if (lane == 0) s_mean[warp][i] = sum[i];
    }
    __syncthreads();
    const int n_s = (int)cnt[0], n_d = (int)cnt[1];

    float acc[4] = {0.f, 0.f, 0.f, 0.f};   // err(src->dst), inliers(src), err(dst->src), inliers(dst)
    if (GRIDNN && n_s > 0 && n_d > 0) {
        float4* extra = g_tile;
        float* scratch = reinterpret_cast<float*>(extra + 2 * gridnn_units(a.N));
        GridTiles td{D, extra, reinterpret_cast<uint32_t*>(extra + a.N), scratch};
        GridTiles ts{S, extra + gridnn_units(a.N), reinterpret_cast<uint32_t*>(extra + gridnn_units(a.N) + a.N), scratch};
        const GridInfo gD = build_grid(td, n_d, a.tau);
        // the candidates of the second direction are the rows of pcd1 under the pose: sorted as they are moved
        const GridInfo gS = build_grid_rows(ts, n_s, a.tau, kCellFactor, [&](int j) { return transform_row(m, S[j]); });
        eval_direction_grid<true>(m, S, n_s, gD, td.sorted_p, reinterpret_cast<const unsigned short*>(td.cells_p), n_d,
                                  a.thr, acc[0], acc[1]);
        eval_direction_grid<false>(m, D, n_d, gS, ts.sorted_p, reinterpret_cast<const unsigned short*>(ts.cells_p), n_s,
                                   a.thr, acc[2], acc[3]);
    } else {
        // (an empty cloud in the grid variant lands here too: its one-row `tile` is a placeholder, the staging area is
        //  the dynamic shared memory the grids would have used -- 2 (N + 257) rows, a tile holds at most min(kEvTile, N))
        float4* stage = GRIDNN ? g_tile : tile;
        eval_direction<true>(m, S, n_s, D, n_d > 0 ? n_d : a.N, a.thr, stage, acc[0], acc[1]);
        eval_direction<false>(m, D, n_d, S, n_s > 0 ? n_s : a.N, a.thr, stage, acc[2], acc[3]);
    }
    __syncthreads();
    block_allreduce_sum<4, kWarps>(acc, s_scratch);

    if (tid == 0) {
        const float fs = (float)n_s, fd = (float)n_d, both = (float)(n_s + n_d);
        a.errors[(size_t)p * 2 + 0] = __fdiv_rn(acc[0], fs);
        a.errors[(size_t)p * 2 + 1] = __fdiv_rn(acc[2], fd);
        a.inliers[(size_t)p * 2 + 0] = acc[1];
        a.inliers[(size_t)p * 2 + 1] = acc[3];
        a.ratios[(size_t)p * 2 + 0] = __fdiv_rn(acc[1], fs);
        a.ratios[(size_t)p * 2 + 1] = __fdiv_rn(acc[3], fd);
        const float iou0 = __fdiv_rn(acc[1], __fsub_rn(both, acc[3])), iou1 = __fdiv_rn(acc[3], __fsub_rn(both, acc[1]));
        a.ious[(size_t)p * 2 + 0] = iou0;
        a.ious[(size_t)p * 2 + 1] = iou1;
        double tot[6];
#pragma unroll
        for (int i = 0; i < 6; ++i) {
            tot[i] = 0.;
            for (int w = 0; w < kWarps; ++w) tot[i] += s_mean[w][i];
        }
        float tr[3];
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            tr[i] = __fsub_rn(__fdiv_rn((float)tot[i], fs), __fdiv_rn((float)tot[3 + i], fs));
            a.translations[(size_t)p * 3 + i] = tr[i];
        }
        // pytorch3d matrix_to_euler_angles(R, "ZYX") * 180. / pi (utils_match.py:184)
        const float pi = 3.14159274101257324f;
        const float rz = __fdiv_rn(__fmul_rn(atan2f(m[4], m[0]), 180.f), pi);
        const float ry = __fdiv_rn(__fmul_rn(asinf(-m[8]), 180.f), pi);
        const float rx = __fdiv_rn(__fmul_rn(atan2f(m[9], m[10]), 180.f), pi);
        a.rotations[(size_t)p * 3 + 0] = rz;
        a.rotations[(size_t)p * 3 + 1] = ry;
        a.rotations[(size_t)p * 3 + 2] = rx;
        if (a.accept) {
            // check_transformation(args, translation, rotation, min(iou)): every test is "reject if x > / < gate", so a
            // NaN operand never rejects (utils_check.py:54-64); python min(iou) keeps iou[0] unless iou[1] < iou[0]
            const float iou = iou1 < iou0 ? iou1 : iou0;
            const float norm = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(tr[0], tr[0]), __fmul_rn(tr[1], tr[1])), __fmul_rn(tr[2], tr[2])));
            const float ay = fabsf(ry), ax = fabsf(rx);
            const float rmax = (ay != ay || ax != ax) ? __int_as_float(0x7fc00000) : fmaxf(ay, ax);   // torch.max keeps NaN
            a.accept[p] = !(norm > a.gate_translation) && !(iou < a.gate_iou) && !(rmax > a.gate_rot);
        }
    }
}

int launch_match_eval(const float* src, const float* dst, const float* pose, int P, int N, float thr, float* errors,
                      float* inliers, float* ratios, float* ious, float* translations, float* rotations,
                      const icpf_match_gates* gates, int* accept, cudaStream_t stream) {
    if (P == 0) return ICPF_OK;
    EvalArgs a{src, dst, pose, N, thr, errors, inliers, ratios, ious, translations, rotations, thr, nullptr, 0.f, 0.f, 0.f};
    if (gates) {
        a.accept = accept;
        a.gate_translation = (float)gates->translation_frame;
        a.gate_iou = (float)gates->thres_iou;
        a.gate_rot = (float)(gates->thres_rot * 90.0);
    }
    const size_t with_grids = ((size_t)2 * gridnn_units(N) + up16(kRedFloats * 4)) * 16;
    const bool gridnn = with_grids <= (size_t)227 * 1024;
    auto kernel = gridnn ? match_eval_kernel<true> : match_eval_kernel<false>;
    const size_t smem = gridnn ? with_grids : 0;
    cudaError_t err = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess) return (int)err;
    ICPF_LAUNCH(kernel, P, kThreads, smem, stream)(a);
    return (int)cudaGetLastError();
}

}  // namespace icpf
