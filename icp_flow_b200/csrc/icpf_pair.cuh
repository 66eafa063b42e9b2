// icpf_pair.cuh -- per-pair device routines: one CTA owns one (src, dst) cluster pair whose padded row blocks live
// in shared memory for the whole registration (HBM is read once per pair).
//
// Reference semantics restated here (file:line under /root/reference):
//   NN with lengths, ties -> lowest index, squared L2 in coordinate order ..... utils_icp_pytorch3d.py:154-156 (+pytorch3d knn_points)
//   gate d^2 <= tau^2, mask = flag && gate ..................................... utils_icp_pytorch3d.py:160-164
//   weighted centroids / centred cross-covariance / rotation / translation ..... utils_icp_pytorch3d.py:314-377
//   Xt = X0 R + T from the INITIAL cloud, rmse of new Xt against old NN ......... utils_icp_pytorch3d.py:177,191-192
//   relative rmse, convergence flag ............................................. utils_icp_pytorch3d.py:195-209
#pragma once

#include "icpf_common.cuh"

namespace icpf {

constexpr int kThreads = 128;          // threads per pair CTA
constexpr int kWarps = kThreads / 32;

// Shared-memory carve-up for one pair (all offsets 16-byte aligned).
struct PairTiles {
    float4* src;    // [N]  (x,y,z,flag) -- the cloud being moved (initial coordinates X0)
    float4* dst;    // [N]  (x,y,z,flag) -- the fixed cloud
    int* nn;        // [N]  NN index into dst of each src row, -1 when masked out
    float* red;     // [kRedFloats] reduction scratch (three disjoint regions)
    float* bcast;   // [16] R (9), T (3), flags
    uint64_t* bar;  // TMA mbarrier
};

constexpr int kRedA = 0;                       // 8 sums
constexpr int kRedB = kRedA + kWarps * 8;      // 9 sums
constexpr int kRedC = kRedB + kWarps * 9;      // 2 sums
constexpr int kRedFloats = kRedC + kWarps * 2;

__host__ __device__ inline size_t pair_smem_bytes(int N) {
    return (size_t)N * 16 * 2 + (size_t)N * 4 + (size_t)(kRedFloats + 16) * 4 + 16;
}

__device__ __forceinline__ PairTiles carve_pair_tiles(unsigned char* base, int N) {
    PairTiles t;
    t.src = reinterpret_cast<float4*>(base);
    t.dst = t.src + N;
    t.nn = reinterpret_cast<int*>(t.dst + N);
    t.red = reinterpret_cast<float*>(t.nn + N);
    t.bcast = t.red + kRedFloats;
    t.bar = reinterpret_cast<uint64_t*>(t.bcast + 16);   // N*36 + (kRedFloats+16)*4 is a multiple of 8
    return t;
}

// Stage both row blocks of pair p with two TMA bulk copies; every thread returns once the bytes have landed.
// `phase` is the parity of the mbarrier phase to wait for (0 for the first use after init).
__device__ __forceinline__ void load_pair_tiles(const PairTiles& t, const float* src_rows, const float* dst_rows, int N,
                                                uint32_t phase) {
    if (threadIdx.x == 0) {
        const uint32_t bytes = (uint32_t)N * 16u;
        mbar_arrive_expect_tx(t.bar, 2u * bytes);
        tma_load_1d(t.src, src_rows, bytes, t.bar);
        tma_load_1d(t.dst, dst_rows, bytes, t.bar);
    }
    mbar_wait(t.bar, phase);
}

// squared L2 exactly as the pinned oracle computes it: d = dx*dx; d += dy*dy; d += dz*dz  (no FMA contraction)
__device__ __forceinline__ float sqdist(float ax, float ay, float az, float bx, float by, float bz) {
    const float dx = __fsub_rn(ax, bx), dy = __fsub_rn(ay, by), dz = __fsub_rn(az, bz);
    return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}

// x R + T, row-vector convention (R row-major r[3*i+j])
__device__ __forceinline__ void apply_rt(const float (&r)[9], const float (&t)[3], float x, float y, float z, float& ox,
                                         float& oy, float& oz) {
    ox = __fadd_rn(fmaf(z, r[6], fmaf(y, r[3], __fmul_rn(x, r[0]))), t[0]);
    oy = __fadd_rn(fmaf(z, r[7], fmaf(y, r[4], __fmul_rn(x, r[1]))), t[1]);
    oz = __fadd_rn(fmaf(z, r[8], fmaf(y, r[5], __fmul_rn(x, r[2]))), t[2]);
}

// Brute-force NN of QB query points held in registers against dst[0, n_d): ties -> lowest index.
template <int QB>
__device__ __forceinline__ void nn_brute(const float4* __restrict__ dst, int n_d, const float (&qx)[QB],
                                         const float (&qy)[QB], const float (&qz)[QB], float (&best)[QB],
                                         int (&bidx)[QB]) {
#pragma unroll
    for (int k = 0; k < QB; ++k) {
        best[k] = __int_as_float(0x7f800000);  // +inf
        bidx[k] = 0;
    }
#pragma unroll 4
    for (int j = 0; j < n_d; ++j) {
        const float4 c = dst[j];  // broadcast LDS.128
#pragma unroll
        for (int k = 0; k < QB; ++k) {
            const float d = sqdist(qx[k], qy[k], qz[k], c.x, c.y, c.z);
            if (d < best[k]) {
                best[k] = d;
                bidx[k] = j;
            }
        }
    }
}

struct IcpResult {
    float r[9];
    float t[3];
    float rmse;
    int iters;          // iterations executed by this pair
    uint32_t conv[4];   // bit k: relative rmse <= thr at iteration k (bits after a fixed-point exit are extrapolated)
};

// The ICP loop for the pair held in `tl` (src = X0 already initialised, dst = Y).  All threads return the same result.
// n_s / n_d are the valid-row counts (knn `lengths`), tau2 = fp32(thres^2).
// init_R / init_T (may be NULL) = init_transform of the reference: used for the first correspondence search only.
__device__ inline IcpResult icp_iterations(const PairTiles& tl, int N, int n_s, int n_d, float tau2, int max_it,
                                           float rel_thr, bool early_exit, const float* init_R = nullptr,
                                           const float* init_T = nullptr) {
    IcpResult res;
#pragma unroll
    for (int i = 0; i < 9; ++i) res.r[i] = init_R ? init_R[i] : ((i % 4 == 0) ? 1.f : 0.f);
#pragma unroll
    for (int i = 0; i < 3; ++i) res.t[i] = init_T ? init_T[i] : 0.f;
    res.rmse = 0.f;
    res.iters = 0;
    res.conv[0] = res.conv[1] = res.conv[2] = res.conv[3] = 0u;
    if (n_s <= 0 || n_d <= 0 || max_it <= 0) return res;   // engine-defined: nothing to align -> identity

    const int tid = threadIdx.x;
    float prev_rmse = 0.f;
    bool have_prev = false;
    constexpr int QB = 4;

    for (int it = 0; it < max_it; ++it) {
        // ---------------- correspondence search on the current cloud + first-pass sums
        float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        for (int q0 = tid; q0 < N; q0 += kThreads * QB) {
            float qx[QB], qy[QB], qz[QB], best[QB];
            int bidx[QB];
            float4 x0[QB];
#pragma unroll
            for (int k = 0; k < QB; ++k) {
                const int q = q0 + k * kThreads;
                x0[k] = (q < N) ? tl.src[q] : make_float4(0.f, 0.f, 0.f, 0.f);
                apply_rt(res.r, res.t, x0[k].x, x0[k].y, x0[k].z, qx[k], qy[k], qz[k]);
            }
            nn_brute<QB>(tl.dst, n_d, qx, qy, qz, best, bidx);
#pragma unroll
            for (int k = 0; k < QB; ++k) {
                const int q = q0 + k * kThreads;
                if (q >= N) continue;
                // rows >= len_s keep dist 0 / idx 0 in knn_points; the flag decides whether they count
                const bool in_len = q < n_s;
                const int j = in_len ? bidx[k] : 0;
                const bool gate = in_len ? (best[k] <= tau2) : true;
                const bool m = (x0[k].w > 0.f) && gate;
                tl.nn[q] = m ? j : -1;
                if (m) {
                    const float4 y = tl.dst[j];
                    acc[0] += 1.f;
                    acc[1] += x0[k].x; acc[2] += x0[k].y; acc[3] += x0[k].z;
                    acc[4] += y.x; acc[5] += y.y; acc[6] += y.z;
                }
            }
        }
        block_allreduce_sum<8, kWarps>(acc, tl.red + kRedA);
        const float W = fmaxf(acc[0], 1e-9f);
        const float mux = __fdiv_rn(acc[1], W), muy = __fdiv_rn(acc[2], W), muz = __fdiv_rn(acc[3], W);
        const float mvx = __fdiv_rn(acc[4], W), mvy = __fdiv_rn(acc[5], W), mvz = __fdiv_rn(acc[6], W);

        // ---------------- second pass: centred cross-covariance H = Xc^T Yc / W
        float h[9] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        for (int q = tid; q < N; q += kThreads) {
            const int j = tl.nn[q];
            if (j < 0) continue;
            const float4 x = tl.src[q];
            const float4 y = tl.dst[j];
            const float ax = x.x - mux, ay = x.y - muy, az = x.z - muz;
            const float bx = y.x - mvx, by = y.y - mvy, bz = y.z - mvz;
            h[0] = fmaf(ax, bx, h[0]); h[1] = fmaf(ax, by, h[1]); h[2] = fmaf(ax, bz, h[2]);
            h[3] = fmaf(ay, bx, h[3]); h[4] = fmaf(ay, by, h[4]); h[5] = fmaf(ay, bz, h[5]);
            h[6] = fmaf(az, bx, h[6]); h[7] = fmaf(az, by, h[7]); h[8] = fmaf(az, bz, h[8]);
        }
        block_allreduce_sum<9, kWarps>(h, tl.red + kRedB);

        // ---------------- rotation / translation by one thread, broadcast through shared memory
        if (tid == 0) {
#pragma unroll
            for (int i = 0; i < 9; ++i) h[i] = __fdiv_rn(h[i], W);
            const Rot3 rot = kabsch_rotation(h);
            float t[3];
            t[0] = mvx - fmaf(muz, rot.r[6], fmaf(muy, rot.r[3], mux * rot.r[0]));
            t[1] = mvy - fmaf(muz, rot.r[7], fmaf(muy, rot.r[4], mux * rot.r[1]));
            t[2] = mvz - fmaf(muz, rot.r[8], fmaf(muy, rot.r[5], mux * rot.r[2]));
            bool same = true;
#pragma unroll
            for (int i = 0; i < 9; ++i) same = same && (__float_as_uint(rot.r[i]) == __float_as_uint(res.r[i]));
#pragma unroll
            for (int i = 0; i < 3; ++i) same = same && (__float_as_uint(t[i]) == __float_as_uint(res.t[i]));
#pragma unroll
            for (int i = 0; i < 9; ++i) tl.bcast[i] = rot.r[i];
#pragma unroll
            for (int i = 0; i < 3; ++i) tl.bcast[9 + i] = t[i];
            tl.bcast[12] = (same && it > 0) ? 1.f : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int i = 0; i < 9; ++i) res.r[i] = tl.bcast[i];
#pragma unroll
        for (int i = 0; i < 3; ++i) res.t[i] = tl.bcast[9 + i];
        const bool fixed = tl.bcast[12] != 0.f;

        // ---------------- rmse of the re-transformed cloud against the correspondences just used
        float sq[2] = {0.f, 0.f};
        for (int q = tid; q < N; q += kThreads) {
            const int j = tl.nn[q];
            if (j < 0) continue;
            const float4 x = tl.src[q];
            const float4 y = tl.dst[j];
            float tx, ty, tz;
            apply_rt(res.r, res.t, x.x, x.y, x.z, tx, ty, tz);
            sq[0] += sqdist(tx, ty, tz, y.x, y.y, y.z);
        }
        block_allreduce_sum<2, kWarps>(sq, tl.red + kRedC);
        const float rmse = sqrtf(__fdiv_rn(sq[0], W));
        const float rel = have_prev ? __fdiv_rn(prev_rmse - rmse, prev_rmse) : 1.0f;
        const bool ok = rel <= rel_thr;
        if (ok && it < 128) res.conv[it >> 5] |= 1u << (it & 31);
        res.rmse = rmse;
        res.iters = it + 1;
        prev_rmse = rmse;
        have_prev = true;
        if (early_exit && fixed) {
            // from here on the state repeats bit for bit: rel = (rmse - rmse) / rmse = 0 (NaN when rmse == 0)
            const bool tail_ok = (rmse > 0.f) && (0.0f <= rel_thr) && (rmse < __int_as_float(0x7f800000));
            if (tail_ok) {
                for (int k = it + 1; k < max_it && k < 128; ++k) res.conv[k >> 5] |= 1u << (k & 31);
            }
            break;
        }
        // bcast[] is rewritten by thread 0 only after two more block-wide barriers (reductions A and B) -> no hazard
    }
    return res;
}

}  // namespace icpf
