// icpf_pair.cuh -- per-pair device routines: one CTA owns one (src, dst) cluster pair whose padded row blocks live
// in shared memory for the whole registration (HBM is read once per pair).
//
// Reference semantics restated here (file:line under /root/reference):
//   NN with lengths, ties -> lowest index, squared L2 in coordinate order ..... utils_icp_pytorch3d.py:154-156 (+pytorch3d knn_points)
//   gate d^2 <= tau^2, mask = flag && gate ..................................... utils_icp_pytorch3d.py:160-164
//   weighted centroids / centred cross-covariance / rotation / translation ..... utils_icp_pytorch3d.py:314-377
//   Xt = X0 R + T from the INITIAL cloud, rmse of new Xt against old NN ......... utils_icp_pytorch3d.py:177,191-192
//   relative rmse, convergence flag ............................................. utils_icp_pytorch3d.py:195-209
//
// Input contract (pad_segment, utils_helper.py:185-196): the n valid rows (flag > 0) of a cloud are rows [0, n).
//
// Correspondence search.  Only neighbours within thres_dist survive the gate, so inside the ICP loop a radius-bounded
// search is result-identical to the reference's brute force (SURVEY.md finding 2): the dst cloud is counting-sorted
// once per pair into a uniform grid in shared memory (cell >= 2*tau, <= kGridMaxCells cells, z fastest so that the
// cells a query needs along z are one contiguous run), and each query inspects the <= 2x2 runs overlapping its
// tau-box.  Candidates are ranked by (squared distance, original row index), i.e. exactly the reference's
// "sequential scan, strict <" tie rule, so brute force (nn_mode 1) and grid (nn_mode 2) agree bit for bit.
#pragma once

#include "icpf_common.cuh"

namespace icpf {

constexpr int kThreads = 128;          // threads per pair CTA
constexpr int kWarps = kThreads / 32;
constexpr int kGridMaxCells = 2048;    // uniform-grid cells per pair (u16 offsets: 4 KB of shared memory)

constexpr int kRedA = 0;                       // 8 sums
constexpr int kRedB = kRedA + kWarps * 8;      // 9 sums
constexpr int kRedC = kRedB + kWarps * 2 * 9;  // 2 sums (region B doubles as the 6+6 min/max scratch of the grid build)
constexpr int kRedFloats = kRedC + kWarps * 2;
constexpr int kCellWords = (kGridMaxCells + 2 + 1) / 2 + 2;   // packed u16 entries 0..G (+pad), as u32 words

// Shared-memory carve-up for one pair (all offsets 16-byte aligned).
struct PairTiles {
    float4* src;      // [N]  (x,y,z,flag) -- the cloud being moved (initial coordinates X0)
    float4* dst;      // [N]  (x,y,z,flag) -- the fixed cloud as stored (TMA landing zone)
    float4* sorted;   // [N]  grid mode: dst rows in cell order, .w = original row index (int bits)
    uint32_t* cells;  // [kCellWords] grid mode: packed u16 run boundaries; run of cell i = [a[i], a[i+1])
    int* nn;          // [N]  correspondence of each src row (index into the candidate array), -1 when masked out
    float* red;       // [kRedFloats] reduction scratch (disjoint regions)
    float* bcast;     // [16] R (9), T (3), flags
    uint64_t* bar;    // TMA mbarrier
};

__host__ __device__ inline size_t pair_smem_bytes(int N, bool grid) {
    size_t b = (size_t)N * 16 * 2;                                        // src + dst
    b += grid ? (size_t)N * 16 + (size_t)kCellWords * 4 : (size_t)N * 4;   // sorted + cells | nn (grid: nn aliases dst)
    b = (b + 15) / 16 * 16;
    return b + (size_t)(kRedFloats + 16) * 4 + 16;
}

template <bool GRID>
__device__ __forceinline__ PairTiles carve_pair_tiles(unsigned char* base, int N) {
    PairTiles t;
    t.src = reinterpret_cast<float4*>(base);
    t.dst = t.src + N;
    unsigned char* p = reinterpret_cast<unsigned char*>(t.dst + N);
    if (GRID) {
        t.sorted = reinterpret_cast<float4*>(p);
        p += (size_t)N * 16;
        t.cells = reinterpret_cast<uint32_t*>(p);
        p += (size_t)kCellWords * 4;
        t.nn = reinterpret_cast<int*>(t.dst);      // the raw dst rows are dead once the grid is built
    } else {
        t.sorted = t.dst;
        t.cells = nullptr;
        t.nn = reinterpret_cast<int*>(p);
        p += (size_t)N * 4;
    }
    p = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(p) + 15) & ~uintptr_t(15));
    t.red = reinterpret_cast<float*>(p);
    t.bcast = t.red + kRedFloats;
    t.bar = reinterpret_cast<uint64_t*>(t.bcast + 16);
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

// ------------------------------------------------------------------------------------------------ uniform grid
struct GridInfo {
    float ox, oy, oz;   // grid origin = bbox min of the valid dst rows
    float inv_c;        // 1 / cell size
    float r;            // padded gate radius in cells (<= 0.5 + eps)
    int gx, gy, gz;
};

__device__ __forceinline__ int grid_cell(const GridInfo& g, float x, float y, float z) {
    const int ix = min(g.gx - 1, max(0, (int)((x - g.ox) * g.inv_c)));
    const int iy = min(g.gy - 1, max(0, (int)((y - g.oy) * g.inv_c)));
    const int iz = min(g.gz - 1, max(0, (int)((z - g.oz) * g.inv_c)));
    return (ix * g.gy + iy) * g.gz + iz;
}

// Counting sort of dst[0, n_d) into tl.sorted by cell; fills the run boundaries in tl.cells.  All threads return the
// same GridInfo.  Uses red region B as scratch; ends with a block barrier.
__device__ inline GridInfo build_grid(const PairTiles& tl, int n_d, float tau) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float INF = __int_as_float(0x7f800000);
    float lo[3] = {INF, INF, INF}, hi[3] = {-INF, -INF, -INF};
    for (int j = tid; j < n_d; j += kThreads) {
        const float4 p = tl.dst[j];
        lo[0] = fminf(lo[0], p.x); lo[1] = fminf(lo[1], p.y); lo[2] = fminf(lo[2], p.z);
        hi[0] = fmaxf(hi[0], p.x); hi[1] = fmaxf(hi[1], p.y); hi[2] = fmaxf(hi[2], p.z);
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo[k] = fminf(lo[k], __shfl_xor_sync(FULL_MASK, lo[k], o));
            hi[k] = fmaxf(hi[k], __shfl_xor_sync(FULL_MASK, hi[k], o));
        }
    }
    float* scr = tl.red + kRedB;
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            scr[warp * 6 + k] = lo[k];
            scr[warp * 6 + 3 + k] = hi[k];
        }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        lo[k] = scr[k];
        hi[k] = scr[3 + k];
#pragma unroll
        for (int w = 1; w < kWarps; ++w) {
            lo[k] = fminf(lo[k], scr[w * 6 + k]);
            hi[k] = fmaxf(hi[k], scr[w * 6 + 3 + k]);
        }
    }
    GridInfo g;
    g.ox = lo[0]; g.oy = lo[1]; g.oz = lo[2];
    float maxabs = 0.f;
#pragma unroll
    for (int k = 0; k < 3; ++k) maxabs = fmaxf(maxabs, fmaxf(fabsf(lo[k]), fabsf(hi[k])));
    // the pad absorbs the fp32 rounding of the cell arithmetic (1 ulp at 50 m is 4e-6 m)
    const float tau_pad = tau + fmaxf(1e-4f, 1e-6f * maxabs);
    const float ex = fminf(fmaxf(hi[0] - lo[0], 0.f), 1e6f), ey = fminf(fmaxf(hi[1] - lo[1], 0.f), 1e6f),
                ez = fminf(fmaxf(hi[2] - lo[2], 0.f), 1e6f);
    float c = 2.0f * tau_pad;
    g.gx = g.gy = g.gz = 1;
    bool fits = false;
    for (int k = 0; k < 40 && !fits; ++k) {
        g.gx = (int)(ex / c) + 1; g.gy = (int)(ey / c) + 1; g.gz = (int)(ez / c) + 1;
        const float cells = (float)g.gx * (float)g.gy * (float)g.gz;
        fits = cells <= (float)kGridMaxCells;
        if (!fits) c *= fmaxf(1.05f, cbrtf(cells / (float)kGridMaxCells));
    }
    if (!fits) { g.gx = g.gy = g.gz = 1; c = 4e6f; }
    g.inv_c = 1.0f / c;
    g.r = tau_pad * g.inv_c;
    const int G = g.gx * g.gy * g.gz;

    uint32_t* w = tl.cells;
    for (int i = tid; i < kCellWords; i += kThreads) w[i] = 0u;
    __syncthreads();
    // counts: entry e = cell + 1 (u16 halves of u32 words; a count never exceeds n_d < 65536 so halves do not carry)
    for (int j = tid; j < n_d; j += kThreads) {
        const float4 p = tl.dst[j];
        const int e = grid_cell(g, p.x, p.y, p.z) + 1;
        atomicAdd(&w[e >> 1], 1u << ((e & 1) * 16));
    }
    __syncthreads();
    // exclusive scan of the counts: afterwards a[i+1] = first sorted position of cell i
    unsigned short* a = reinterpret_cast<unsigned short*>(w);
    const int chunk = (G + kThreads - 1) / kThreads;
    const int c0 = min(G, tid * chunk), c1 = min(G, c0 + chunk);
    unsigned int local = 0;
    for (int i = c0; i < c1; ++i) local += a[i + 1];
    unsigned int incl = local;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned int v = __shfl_up_sync(FULL_MASK, incl, o);
        if (lane >= o) incl += v;
    }
    unsigned int* wtot = reinterpret_cast<unsigned int*>(tl.red + kRedB) + 32;   // beyond the min/max scratch
    if (lane == 31) wtot[warp] = incl;
    __syncthreads();
    unsigned int run = incl - local;
    for (int k = 0; k < warp; ++k) run += wtot[k];
    for (int i = c0; i < c1; ++i) {
        const unsigned int cnt = a[i + 1];
        a[i + 1] = (unsigned short)run;
        run += cnt;
    }
    __syncthreads();
    // scatter: the atomic turns "first position" into "one past the last", i.e. the first position of the next cell
    for (int j = tid; j < n_d; j += kThreads) {
        const float4 p = tl.dst[j];
        const int e = grid_cell(g, p.x, p.y, p.z) + 1;
        const int sh = (e & 1) * 16;
        const unsigned int old = atomicAdd(&w[e >> 1], 1u << sh);
        tl.sorted[(old >> sh) & 0xffffu] = make_float4(p.x, p.y, p.z, __int_as_float(j));
    }
    __syncthreads();
    return g;
}

// Radius-bounded NN: best candidate among the cells overlapping the padded tau-box of q, ranked by (d^2, original row).
// bj = -1 when no candidate was inspected (the true NN is then farther than tau, so the gate fails either way).
__device__ __forceinline__ void nn_grid(const GridInfo& g, const float4* __restrict__ sorted,
                                        const unsigned short* __restrict__ a, float qx, float qy, float qz, float& best,
                                        int& bj) {
    best = __int_as_float(0x7f800000);
    bj = -1;
    int borig = 0x7fffffff;
    const float fx = (qx - g.ox) * g.inv_c, fy = (qy - g.oy) * g.inv_c, fz = (qz - g.oz) * g.inv_c;
    const float x0 = floorf(fx - g.r), x1 = floorf(fx + g.r);
    const float y0 = floorf(fy - g.r), y1 = floorf(fy + g.r);
    const float z0 = floorf(fz - g.r), z1 = floorf(fz + g.r);
    // (NaN coordinates fail every comparison below and fall through to "no candidate")
    if (!(x1 >= 0.f && y1 >= 0.f && z1 >= 0.f && x0 <= (float)(g.gx - 1) && y0 <= (float)(g.gy - 1) &&
          z0 <= (float)(g.gz - 1)))
        return;
    const int ix0 = max(0, (int)x0), ix1 = min(g.gx - 1, (int)x1);
    const int iy0 = max(0, (int)y0), iy1 = min(g.gy - 1, (int)y1);
    const int iz0 = max(0, (int)z0), iz1 = min(g.gz - 1, (int)z1);
    for (int ix = ix0; ix <= ix1; ++ix) {
        for (int iy = iy0; iy <= iy1; ++iy) {
            const int base = (ix * g.gy + iy) * g.gz;
            const int s = a[base + iz0], e = a[base + iz1 + 1];
            for (int j = s; j < e; ++j) {
                const float4 c = sorted[j];
                const float d = sqdist(qx, qy, qz, c.x, c.y, c.z);
                const int oi = __float_as_int(c.w);
                if (d < best || (d == best && oi < borig)) {
                    best = d;
                    bj = j;
                    borig = oi;
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------ ICP loop
struct IcpResult {
    float r[9];
    float t[3];
    float rmse;
    int iters;                    // iterations executed by this pair
    unsigned long long conv_lo;   // bit k: relative rmse <= thr at iteration k      (k < 64)
    unsigned long long conv_hi;   //                                                 (64 <= k < 128)
};                                // bits after a fixed-point exit are extrapolated (the state repeats)

__device__ __forceinline__ void set_conv_bit(IcpResult& r, int k) {
    if (k < 64) r.conv_lo |= 1ull << k;
    else if (k < 128) r.conv_hi |= 1ull << (k - 64);
}

// The ICP loop for the pair held in `tl`.  All threads return the same result.
//   GRID  : candidates = tl.sorted + grid `g` (build_grid must have run); otherwise candidates = tl.dst, brute force
//   n_s / n_d : valid-row counts (knn `lengths`); tau2 = fp32(thres^2)
//   init_R / init_T (may be NULL) = init_transform of the reference: used for the first correspondence search only.
template <bool GRID>
__device__ inline IcpResult icp_iterations(const PairTiles& tl, const GridInfo& g, int n_s, int n_d, float tau2,
                                           int max_it, float rel_thr, bool early_exit, const float* init_R = nullptr,
                                           const float* init_T = nullptr) {
    IcpResult res;
#pragma unroll
    for (int i = 0; i < 9; ++i) res.r[i] = init_R ? init_R[i] : ((i % 4 == 0) ? 1.f : 0.f);
#pragma unroll
    for (int i = 0; i < 3; ++i) res.t[i] = init_T ? init_T[i] : 0.f;
    res.rmse = 0.f;
    res.iters = 0;
    res.conv_lo = res.conv_hi = 0ull;
    if (n_s <= 0 || n_d <= 0 || max_it <= 0) return res;   // engine-defined: nothing to align -> identity

    const int tid = threadIdx.x;
    const float4* __restrict__ cand = GRID ? tl.sorted : tl.dst;
    const unsigned short* cell_runs = reinterpret_cast<const unsigned short*>(tl.cells);
    float prev_rmse = 0.f;
    bool have_prev = false;
    constexpr int QB = 4;

    for (int it = 0; it < max_it; ++it) {
        // ---------------- correspondence search on the current cloud + first-pass sums
        float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        if (GRID) {
            for (int q = tid; q < n_s; q += kThreads) {
                const float4 x0 = tl.src[q];
                float qx, qy, qz, best;
                int j;
                apply_rt(res.r, res.t, x0.x, x0.y, x0.z, qx, qy, qz);
                nn_grid(g, cand, cell_runs, qx, qy, qz, best, j);
                const bool m = (j >= 0) && (best <= tau2) && (x0.w > 0.f);
                tl.nn[q] = m ? j : -1;
                if (m) {
                    const float4 y = cand[j];
                    acc[0] += 1.f;
                    acc[1] += x0.x; acc[2] += x0.y; acc[3] += x0.z;
                    acc[4] += y.x; acc[5] += y.y; acc[6] += y.z;
                }
            }
        } else {
            for (int q0 = tid; q0 < n_s; q0 += kThreads * QB) {
                float qx[QB], qy[QB], qz[QB], best[QB];
                int bidx[QB];
                float4 x0[QB];
#pragma unroll
                for (int k = 0; k < QB; ++k) {
                    const int q = q0 + k * kThreads;
                    x0[k] = (q < n_s) ? tl.src[q] : make_float4(0.f, 0.f, 0.f, 0.f);
                    apply_rt(res.r, res.t, x0[k].x, x0[k].y, x0[k].z, qx[k], qy[k], qz[k]);
                }
                nn_brute<QB>(cand, n_d, qx, qy, qz, best, bidx);
#pragma unroll
                for (int k = 0; k < QB; ++k) {
                    const int q = q0 + k * kThreads;
                    if (q >= n_s) continue;
                    const bool m = (best[k] <= tau2) && (x0[k].w > 0.f);
                    tl.nn[q] = m ? bidx[k] : -1;
                    if (m) {
                        const float4 y = cand[bidx[k]];
                        acc[0] += 1.f;
                        acc[1] += x0[k].x; acc[2] += x0[k].y; acc[3] += x0[k].z;
                        acc[4] += y.x; acc[5] += y.y; acc[6] += y.z;
                    }
                }
            }
        }
        block_allreduce_sum<8, kWarps>(acc, tl.red + kRedA);
        const float W = fmaxf(acc[0], 1e-9f);
        const float mux = __fdiv_rn(acc[1], W), muy = __fdiv_rn(acc[2], W), muz = __fdiv_rn(acc[3], W);
        const float mvx = __fdiv_rn(acc[4], W), mvy = __fdiv_rn(acc[5], W), mvz = __fdiv_rn(acc[6], W);

        // ---------------- second pass: centred cross-covariance H = Xc^T Yc / W
        float h[9] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        for (int q = tid; q < n_s; q += kThreads) {
            const int j = tl.nn[q];
            if (j < 0) continue;
            const float4 x = tl.src[q];
            const float4 y = cand[j];
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
        for (int q = tid; q < n_s; q += kThreads) {
            const int j = tl.nn[q];
            if (j < 0) continue;
            const float4 x = tl.src[q];
            const float4 y = cand[j];
            float tx, ty, tz;
            apply_rt(res.r, res.t, x.x, x.y, x.z, tx, ty, tz);
            sq[0] += sqdist(tx, ty, tz, y.x, y.y, y.z);
        }
        block_allreduce_sum<2, kWarps>(sq, tl.red + kRedC);
        const float rmse = sqrtf(__fdiv_rn(sq[0], W));
        const float rel = have_prev ? __fdiv_rn(prev_rmse - rmse, prev_rmse) : 1.0f;
        if (rel <= rel_thr) set_conv_bit(res, it);
        res.rmse = rmse;
        res.iters = it + 1;
        prev_rmse = rmse;
        have_prev = true;
        if (early_exit && fixed) {
            // from here on the state repeats bit for bit: rel = (rmse - rmse) / rmse = 0 (NaN when rmse == 0)
            const bool tail_ok = (rmse > 0.f) && (0.0f <= rel_thr) && (rmse < __int_as_float(0x7f800000));
            if (tail_ok) {
                for (int k = it + 1; k < max_it && k < 128; ++k) set_conv_bit(res, k);
            }
            break;
        }
        // bcast[] is rewritten by thread 0 only after two more block-wide barriers (reductions A and B) -> no hazard
    }
    return res;
}

}  // namespace icpf
