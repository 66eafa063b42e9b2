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
// once per pair into a uniform grid in shared memory (cell = kCellFactor * tau, <= kGridMaxCells cells, z fastest so
// that the cells a query needs along z are one contiguous run), and each query inspects the <= 2x2 runs overlapping
// its tau-box.  Candidates are ranked by (squared distance, original row index), i.e. exactly the reference's
// "sequential scan, strict <" tie rule, so brute force (nn_mode 1), grid (2) and grid + cache (3) agree bit for bit.
#pragma once

#include <cuda_fp16.h>

#include "icpf_common.cuh"

namespace icpf {

#ifndef ICPF_PAIR_THREADS
#define ICPF_PAIR_THREADS 128
#endif
constexpr int kThreads = ICPF_PAIR_THREADS;          // threads per pair CTA
constexpr int kWarps = kThreads / 32;
constexpr float kCellFactor = 2.2f;  // grid cell size in units of the padded gate radius (>= 2.002)
constexpr int kGridMaxCells = 2048;    // uniform-grid cells per pair (u16 offsets: 4 KB of shared memory)

// reduction scratch (floats): per-warp partials of the 18 per-iteration sums, their totals; the grid build reuses it
constexpr int kSums = 18;                      // 16 moments + rmse numerator + rows searched
constexpr int kScrPart = 0;                    // [kWarps][kSums]
constexpr int kScrTotal = kScrPart + kWarps * kSums;
constexpr int kRedFloats = kScrTotal + 24;
constexpr int kBcastFloats = 64;
constexpr int kCellWords = (kGridMaxCells + 2 + 1) / 2 + 2;   // packed u16 entries 0..G (+pad), as u32 words

// broadcast block written by thread 0 once per iteration
enum : int { B_R = 0, B_T = 9, B_RC = 12 /* last step R_k - R_{k-1} */, B_TC = 21 /* last step T_k - T_{k-1} */, B_PX = 24, B_PY = 27, B_EXIT = 30,
             B_SEARCH = 31 /* int: rows queued for a search */, B_KABSCH = 32 /* KabschState: 9 floats + 2 flags */,
             B_HPREV = 44 /* cross-covariance of the previous solve */,
             B_DIRTY = 53 /* int[kWarps]: a search changed a correspondence of this warp's rows */, B_PIVMOVED = 57 /* the pivots moved */,
             B_ZEROSTEP = 58 /* the last step was exactly zero */ };
static_assert(B_ZEROSTEP < kBcastFloats && B_HPREV + 9 <= B_DIRTY, "broadcast block");

// Dynamic shared memory of every pair kernel.  Tiles are addressed as OFFSETS into this one array so that the
// compiler always knows the address space (a run-time swap of two pointers degrades every access to a generic LD/ST).
ICPF_DYN_SHARED __align__(128) float4 g_tile[];

// Shared-memory carve-up for one pair (float4 / 16-byte units unless noted).
struct PairTiles {
    int src_off;      // [N] float4 (x,y,z,flag) -- the cloud being moved (initial coordinates X0)
    int dst_off;      // [N] float4 (x,y,z,flag) -- the fixed cloud as stored (TMA landing zone)
    int sorted_off;   // [N] float4 grid mode: dst rows in cell order, .w = (original row << 16 | sorted position)
    int cells_off;    // [kCellWords] u32 grid mode: packed u16 run boundaries; run of cell i = [a[i], a[i+1])
    int nn_off;       // [N] u32 correspondence word of each src row (NnWord); grid mode: inside the dead raw dst rows
    int defer_off;    // [kWarps * defer_cap >= N] u16 grid mode: rows whose cached neighbour could not be proven
    int defer_cap;
    int red_off;      // [kRedFloats] float reduction scratch
    static constexpr int kPosBits = 13;
    __device__ __forceinline__ float4* src() const { return g_tile + src_off; }
    __device__ __forceinline__ float4* dst() const { return g_tile + dst_off; }
    __device__ __forceinline__ float4* sorted() const { return g_tile + sorted_off; }
    __device__ __forceinline__ uint32_t* cells() const { return reinterpret_cast<uint32_t*>(g_tile + cells_off); }
    __device__ __forceinline__ unsigned int* nn() const { return reinterpret_cast<unsigned int*>(g_tile + nn_off); }
    __device__ __forceinline__ unsigned short* defer() const { return reinterpret_cast<unsigned short*>(g_tile + defer_off); }
    __device__ __forceinline__ float* red() const { return reinterpret_cast<float*>(g_tile + red_off); }
    __device__ __forceinline__ float* bcast() const { return red() + kRedFloats; }
    __device__ __forceinline__ uint64_t* bar() const { return reinterpret_cast<uint64_t*>(bcast() + kBcastFloats); }
    // the smaller cloud is the moved one (utils_match.py:139-146): exchange the roles of the two staged tiles
    template <bool GRID>
    __device__ __forceinline__ void swap_clouds() {
        const int t = src_off; src_off = dst_off; dst_off = t;
        if (GRID) nn_off = dst_off;      // the correspondence words alias the raw rows of the (new) fixed cloud
        else sorted_off = dst_off;
    }
};

__host__ __device__ inline int pair_defer_cap(int N) { return (N + kThreads - 1) / kThreads * 32; }
__host__ __device__ inline int up16(int bytes) { return (bytes + 15) / 16; }

__host__ __device__ inline size_t pair_smem_bytes(int N, bool grid) {
    int u = 2 * N;                                                        // src + dst
    if (grid) u += N + up16(kCellWords * 4) + up16(pair_defer_cap(N) * kWarps * 2);
    else u += up16(N * 4);                                                // nn (grid mode: nn aliases the raw dst rows)
    u += up16((kRedFloats + kBcastFloats) * 4 + 16);
    return (size_t)u * 16;
}

// Large clusters (row blocks that do not fit shared memory, max_points up to 10 000): the same interface, but the rows
// stay in global memory (L2-resident: one pair is 2 x 16 N bytes) and the sorted copy, the correspondence words and the
// search list live in a caller-provided workspace; only the grid runs, the reduction scratch and the broadcast block
// are in shared memory.
struct PairTilesG {
    float4* src_p;
    float4* dst_p;
    float4* sorted_p;
    unsigned int* nn_p;
    unsigned short* defer_p;
    int defer_cap;
    static constexpr int kPosBits = 14;
    __device__ __forceinline__ float4* src() const { return src_p; }
    __device__ __forceinline__ float4* dst() const { return dst_p; }
    __device__ __forceinline__ float4* sorted() const { return sorted_p; }
    __device__ __forceinline__ uint32_t* cells() const { return reinterpret_cast<uint32_t*>(g_tile); }
    __device__ __forceinline__ unsigned int* nn() const { return nn_p; }
    __device__ __forceinline__ unsigned short* defer() const { return defer_p; }
    __device__ __forceinline__ float* red() const { return reinterpret_cast<float*>(g_tile + up16(kCellWords * 4)); }
    __device__ __forceinline__ float* bcast() const { return red() + kRedFloats; }
    template <bool GRID>
    __device__ __forceinline__ void swap_clouds() {
        float4* t = src_p; src_p = dst_p; dst_p = t;
        if (!GRID) sorted_p = dst_p;
    }
};

__host__ __device__ inline size_t pair_global_smem_bytes() {
    return (size_t)(up16(kCellWords * 4) + up16((kRedFloats + kBcastFloats) * 4 + 16)) * 16;
}
// per-pair workspace of the global-memory variant: sorted rows | transformed src rows | nn words | search list
__host__ __device__ inline size_t pair_global_ws_bytes(int N) {
    const size_t n = (size_t)(N + kThreads - 1) / kThreads * kThreads;
    return n * 16 + n * 16 + n * 4 + ((n * 2 + 15) / 16 * 16);
}
__host__ inline bool pair_needs_global(int N) { return pair_smem_bytes(N, true) > (size_t)227 * 1024; }

template <bool GRID>
__device__ __forceinline__ PairTiles carve_pair_tiles(int N) {
    PairTiles t;
    t.src_off = 0;
    t.dst_off = N;
    int u = 2 * N;
    t.defer_cap = pair_defer_cap(N);
    if (GRID) {
        t.sorted_off = u; u += N;
        t.cells_off = u; u += up16(kCellWords * 4);
        t.defer_off = u; u += up16(t.defer_cap * kWarps * 2);
        t.nn_off = t.dst_off;              // the raw dst rows are dead once the grid is built
    } else {
        t.sorted_off = t.dst_off;
        t.cells_off = 0;
        t.defer_off = 0;
        t.nn_off = u; u += up16(N * 4);
    }
    t.red_off = u;
    return t;
}

// Stage both row blocks of pair p with two TMA bulk copies; every thread returns once the bytes have landed.
// `phase` is the parity of the mbarrier phase to wait for (0 for the first use after init).
__device__ __forceinline__ void load_pair_tiles(const PairTiles& t, const float* src_rows, const float* dst_rows, int N,
                                                uint32_t phase) {
    if (threadIdx.x == 0) {
        const uint32_t bytes = (uint32_t)N * 16u;
        mbar_arrive_expect_tx(t.bar(), 2u * bytes);
        tma_load_1d(t.src(), src_rows, bytes, t.bar());
        tma_load_1d(t.dst(), dst_rows, bytes, t.bar());
    }
    mbar_wait(t.bar(), phase);
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

// [x y z 1] pose^T for the first three rows of a row-major 4x4 (m[0..11]); the flag column is carried through
// (utils_helper.transform_points_batch, utils_helper.py:76-87)
__device__ __forceinline__ float4 transform_row(const float (&m)[12], const float4& p) {
    float4 o;
    o.x = fmaf(1.0f, m[3], fmaf(p.z, m[2], fmaf(p.y, m[1], p.x * m[0])));
    o.y = fmaf(1.0f, m[7], fmaf(p.z, m[6], fmaf(p.y, m[5], p.x * m[4])));
    o.z = fmaf(1.0f, m[11], fmaf(p.z, m[10], fmaf(p.y, m[9], p.x * m[8])));
    o.w = p.w;
    return o;
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
    float r;            // padded gate radius in cells (< 0.5, so a query overlaps at most 2 cells per axis)
    float c;            // cell size (m)
    float pad;          // slack (m) that absorbs the fp32 rounding of the cell arithmetic
    int gx, gy, gz;
};

__device__ __forceinline__ int grid_cell(const GridInfo& g, float x, float y, float z) {
    const int ix = min(g.gx - 1, max(0, (int)((x - g.ox) * g.inv_c)));
    const int iy = min(g.gy - 1, max(0, (int)((y - g.oy) * g.inv_c)));
    const int iz = min(g.gz - 1, max(0, (int)((z - g.oz) * g.inv_c)));
    return (ix * g.gy + iy) * g.gz + iz;
}

// Counting sort of dst[0, n_d) into tl.sorted() by cell; fills the run boundaries in tl.cells().  All threads return the
// same GridInfo.  Uses the reduction scratch; ends with a block barrier.
// `row(j)` yields row j of the cloud being sorted (the stored row, or a row moved on the fly).
template <class Tiles, class RowFn>
__device__ __forceinline__ GridInfo build_grid_rows(const Tiles& tl, int n_d, float tau, float cell_factor, RowFn row) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float INF = __int_as_float(0x7f800000);
    float lo[3] = {INF, INF, INF}, hi[3] = {-INF, -INF, -INF};
    for (int j = tid; j < n_d; j += kThreads) {
        const float4 p = row(j);
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
    float* scr = tl.red() + kScrPart;
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
    g.pad = fmaxf(1e-4f, 1e-6f * maxabs);
    const float tau_pad = tau + g.pad;
    const float ex = fminf(fmaxf(hi[0] - lo[0], 0.f), 1e6f), ey = fminf(fmaxf(hi[1] - lo[1], 0.f), 1e6f),
                ez = fminf(fmaxf(hi[2] - lo[2], 0.f), 1e6f);
    // cells of kCellFactor*tau: a query then overlaps <= 2 cells per axis while everything it does NOT inspect is at
    // least 0.4995 cells (~1.25 tau) away -- the head-room the correspondence cache needs to prove "still masked"
    float c = fmaxf(cell_factor, 2.002f) * tau_pad;
    g.gx = g.gy = g.gz = 1;
    bool fits = false;
    for (int k = 0; k < 40 && !fits; ++k) {
        g.gx = (int)(ex / c) + 1; g.gy = (int)(ey / c) + 1; g.gz = (int)(ez / c) + 1;
        const float cells = (float)g.gx * (float)g.gy * (float)g.gz;
        fits = cells <= (float)kGridMaxCells;
        if (!fits) c *= fmaxf(1.05f, cbrtf(cells / (float)kGridMaxCells));
    }
    if (!fits) { g.gx = g.gy = g.gz = 1; c = 4e6f; }
    g.c = c;
    g.inv_c = 1.0f / c;
    g.r = 0.4995f;   // >= tau_pad / c because c >= 2.002 * tau_pad
    const int G = g.gx * g.gy * g.gz;

    uint32_t* w = tl.cells();
    for (int i = tid; i < kCellWords; i += kThreads) w[i] = 0u;
    __syncthreads();
    // counts: entry e = cell + 1 (u16 halves of u32 words; a count never exceeds n_d < 65536 so halves do not carry)
    for (int j = tid; j < n_d; j += kThreads) {
        const float4 p = row(j);
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
    unsigned int* wtot = reinterpret_cast<unsigned int*>(tl.red() + kScrPart) + 32;   // beyond the min/max scratch
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
        const float4 p = row(j);
        const int e = grid_cell(g, p.x, p.y, p.z) + 1;
        const int sh = (e & 1) * 16;
        const unsigned int old = atomicAdd(&w[e >> 1], 1u << sh);
        const unsigned int pos = (old >> sh) & 0xffffu;
        // .w = (original row << 16) | sorted position: the low word of the 64-bit ranking key (d^2 bits, row, position)
        tl.sorted()[pos] = make_float4(p.x, p.y, p.z, __uint_as_float(((unsigned int)j << 16) | pos));   // pos < 2^13
    }
    __syncthreads();
    return g;
}

template <class Tiles>
__device__ __forceinline__ GridInfo build_grid(const Tiles& tl, int n_d, float tau, float cell_factor = kCellFactor) {
    const float4* rows = tl.dst();
    return build_grid_rows(tl, n_d, tau, cell_factor, [rows](int j) { return rows[j]; });
}

// Radius-bounded NN: the best candidate among the cells overlapping the box q +- g.r cells (the padded gate
// radius), ranked by the 64-bit key (d^2 bits, original row, sorted position) -- i.e. (squared distance, row
// index), the reference's tie rule.
//   pos1/d1   sorted position / squared distance of the winner (-1 / +inf when no candidate was inspected: the true NN
//             is then farther than tau)
//   d2        second smallest squared distance among the inspected candidates (+inf when fewer than two)
//   box       distance (m) from q to the nearest face of the inspected block that has un-inspected space behind it:
//             every point that was NOT inspected is farther than `box` from q
struct NnTop2 {
    float d1, d2, box;
    int pos1;
};

__device__ __forceinline__ NnTop2 grid_search(const GridInfo& g, const float4* __restrict__ sorted,
                                              const unsigned short* __restrict__ a, float qx, float qy, float qz) {
    const float INF = __int_as_float(0x7f800000);
    const unsigned long long kNone = (0x7f800000ull << 32) | 0xffffffffull;
    unsigned long long key1 = kNone;
    NnTop2 o;
    o.d2 = INF;
    const float r = g.r;
    const float fx = (qx - g.ox) * g.inv_c, fy = (qy - g.oy) * g.inv_c, fz = (qz - g.oz) * g.inv_c;
    const float x0 = floorf(fx - r), x1 = floorf(fx + r);
    const float y0 = floorf(fy - r), y1 = floorf(fy + r);
    const float z0 = floorf(fz - r), z1 = floorf(fz + r);
    const float hx = (float)(g.gx - 1), hy = (float)(g.gy - 1), hz = (float)(g.gz - 1);
    // (NaN coordinates fail every comparison below and fall through to "no candidate")
    if (!(x1 >= 0.f && y1 >= 0.f && z1 >= 0.f && x0 <= hx && y0 <= hy && z0 <= hz)) {
        // the box misses the grid: all points lie inside [0, g]^3 cell coordinates
        const float gapx = fmaxf(-fx, fx - (hx + 1.f)), gapy = fmaxf(-fy, fy - (hy + 1.f)),
                    gapz = fmaxf(-fz, fz - (hz + 1.f));
        o.box = fmaxf(fmaxf(gapx, fmaxf(gapy, gapz)) * g.c - g.pad, 0.f);
        o.d1 = INF;
        o.pos1 = -1;
        return o;
    }
    // faces clamped by the grid boundary have no points behind them
    const float bx = fminf(x0 < 0.f ? INF : fx - x0, x1 > hx ? INF : x1 + 1.f - fx);
    const float by = fminf(y0 < 0.f ? INF : fy - y0, y1 > hy ? INF : y1 + 1.f - fy);
    const float bz = fminf(z0 < 0.f ? INF : fz - z0, z1 > hz ? INF : z1 + 1.f - fz);
    o.box = fmaxf(fminf(bx, fminf(by, bz)) * g.c - g.pad, 0.f);
    const int ix0 = max(0, (int)x0), ix1 = min(g.gx - 1, (int)x1);
    const int iy0 = max(0, (int)y0), iy1 = min(g.gy - 1, (int)y1);
    const int iz0 = max(0, (int)z0), iz1 = min(g.gz - 1, (int)z1);
    // r < 0.5 cells: the block is at most 2 x 2 columns, and the cells a column contributes are one contiguous run of the
    // sorted rows (z runs fastest).  The <= 4 runs are walked as ONE flat candidate sequence t = 0 .. total-1, so that the
    // lanes of a warp never wait for each other's runs: a lane's trip count is its own number of candidates.
    const bool two_x = ix1 > ix0, two_y = iy1 > iy0;
    const int b00 = (ix0 * g.gy + iy0) * g.gz, b01 = b00 + g.gz, b10 = b00 + g.gy * g.gz, b11 = b10 + g.gz;
    const int s0 = a[b00 + iz0], e0 = a[b00 + iz1 + 1];
    int s1 = 0, e1 = 0, s2 = 0, e2 = 0, s3 = 0, e3 = 0;
    if (two_y) { s1 = a[b01 + iz0]; e1 = a[b01 + iz1 + 1]; }
    if (two_x) { s2 = a[b10 + iz0]; e2 = a[b10 + iz1 + 1]; }
    if (two_x && two_y) { s3 = a[b11 + iz0]; e3 = a[b11 + iz1 + 1]; }
    const int c1 = e0 - s0, c2 = c1 + (e1 - s1), c3 = c2 + (e2 - s2), total = c3 + (e3 - s3);
    const int o0 = s0, o1 = s1 - c1, o2 = s2 - c2, o3 = s3 - c3;       // sorted position = t + offset of t's run
    for (int t = 0; t < total; ++t) {
        const int off = t < c2 ? (t < c1 ? o0 : o1) : (t < c3 ? o2 : o3);
        const float4 c = sorted[t + off];
        const float d = sqdist(qx, qy, qz, c.x, c.y, c.z);
        const unsigned long long k = ((unsigned long long)__float_as_uint(d) << 32) | __float_as_uint(c.w);
        const bool better = k < key1;
        o.d2 = fminf(o.d2, better ? __uint_as_float((unsigned int)(key1 >> 32)) : d);
        key1 = better ? k : key1;
    }
    o.d1 = __uint_as_float((unsigned int)(key1 >> 32));
    o.pos1 = (key1 == kNone) ? -1 : (int)((unsigned int)key1 & 0xffffu);
    return o;
}

// correspondence word kept per src row (PB = position bits: 13 for the shared-memory tiles, 14 for the large-cluster
// variant): bits [0,PB) sorted position of the best candidate (all ones = none), bit PB "masked out", the remaining
// high bits a lower bound on the distance of every OTHER dst point (the top bits of the fp32 pattern without its sign:
// truncation rounds a positive value down; 10 / 9 mantissa bits)
template <int PB> struct NnWord {
    static constexpr unsigned int kPosMask = (1u << PB) - 1u;
    static constexpr unsigned int kNone = kPosMask;
    static constexpr unsigned int kMasked = 1u << PB;
    static constexpr unsigned int kBoundMask = ~((1u << (PB + 1)) - 1u);
    static constexpr int kMaxRows = (1 << PB) - 2;
    __device__ __forceinline__ static unsigned int pack(int pos, float bound, bool used) {
        const unsigned int b = (__float_as_uint(fmaxf(bound, 0.f)) << 1) & kBoundMask;
        return (pos < 0 ? kNone : (unsigned int)pos) | b | (used ? 0u : kMasked);
    }
    __device__ __forceinline__ static float bound(unsigned int w) { return __uint_as_float((w & kBoundMask) >> 1); }
};
constexpr int kMaxRows = NnWord<14>::kMaxRows;      // rows per cloud the engine accepts

}  // namespace icpf
