// icpf_histfused.cu -- votes + non-maximum suppression + top-5 of ONE pair inside ONE CTA, histogram in shared memory.
//
// The difference histogram of a pair only has support where differences X_i - Y_j exist: the box
// [min X - max Y, max X - min Y], usually a few metres wide, while the reference allocates (and zero-fills, pools and
// sorts) the full [lx, ly, lz] volume for every pair (135 x 135 x 3 fp32 = 219 KB at the default translation_frame;
// utils_hist.py:69-77, hist_cuda.cu:59).  Here the CTA derives the bin range that box can reach, keeps exactly that
// sub-histogram as u32 counters in shared memory (shared-memory atomics instead of L2 atomics), and runs the 11^3
// max-pool / top-5 on it in place -- bins outside the box are zero by construction, so the result is identical to
// pooling the full volume.  Pairs whose box needs more columns than fit are flagged and take the global-memory path
// (icpf_hist.cu).  Vote arithmetic is the bit-compatible restatement of hist_cuda_core.cuh:48-60.
#include "icpf_internal.h"
#include "icpf_common.cuh"

#include <math.h>
#include <string.h>
#include <limits>
#include <type_traits>

namespace icpf {

constexpr int kFusedTile = 1024;       // Y rows staged per tile
constexpr int kFusedTopK = 5;
constexpr int kFusedNmsHalf = 5;

struct FusedHistArgs {
    const float4* X;         // the caller's dst  [P,N,4]   (votes X_i - Y_j; roles swap with auto_swap)
    const float4* Y;         // the caller's src  [P,N,4]
    int N;
    float min_x, min_y, min_z, max_x, max_y, max_z;
    int len_x, len_y, len_z;
    int auto_swap;
    int cap_cols;            // sub-histogram columns (x,y) that fit the dynamic shared memory
    int* out_idx;            // [P,5]
    float* out_votes;        // [P,5]
    int* need_global;        // [P] 1 = this pair did not fit and must take the next (wider / global) path
    int only_flagged;        // 1: second tier -- handle only the pairs the first tier flagged
    float zt1, zt2;          // MODE 2 (three z bins): smallest differences in [min_z, max_z) that fall into bin 1 / bin 2
};

__device__ __forceinline__ int vote_bin(float v, float mn, float range, float flen, int len) {
    const int p = __float2int_rd(__fmul_rn(__fdiv_rn(__fsub_rn(v, mn), range), flen));
    return min(p, len - 1);
}

// The same bin with the IEEE division (~25 instructions, three per vote: 40 % of the kernel) replaced by Markstein's
// correction step: with y = RN(1/b), q0 = RN(x y) and the exact remainder r = x - q0 b (one FMA), RN(q0 + r y) is the
// correctly rounded quotient x / b for every x whenever the significand of b is not all ones (Markstein 1990; the
// fast path of every IEEE-compliant software division).  launch_hist_fused() checks that condition and the exponent
// range on the host and selects the plain division otherwise; tools/check_fastdiv.cu compares the two bit for bit for
// ALL dividends in [0, b) for the divisors the reference's settings produce.
__device__ __forceinline__ int vote_bin_fast(float v, float mn, float range, float inv_range, float flen, int len) {
    const float x = __fsub_rn(v, mn);
    const float q0 = __fmul_rn(x, inv_range);
    const float q = __fmaf_rn(__fmaf_rn(-q0, range, x), inv_range, q0);
    const int p = __float2int_rd(__fmul_rn(q, flen));
    return min(p, len - 1);
}

__device__ __forceinline__ unsigned long long fused_peak_key(float v, int idx) {
    return ((unsigned long long)__float_as_uint(v) << 32) | (unsigned int)(0x7fffffff - idx);
}

// kFusedThreads: the sub-histogram of the widest pair fixes the shared memory of the launch (~200 KB at the default
// 135 x 135 x 3 bins), i.e. ONE CTA per SM whatever its size -- so large clusters run 1024 threads (32 warps per SM
// instead of 8) and small ones 256.
// COUNT16: second tier for pairs whose sub-histogram does not fit as u32 counters -- u16 counters and max planes (10 B
// per column at three z bins: the whole 135 x 135 window of the default translation_frame fits).  A counter that is
// about to wrap is caught on the increment that wraps it (the atomic returns the old value) and sends the pair on to
// the exact global-memory path.
// MODE: 0 the plain IEEE division, 1 Markstein's division (FASTDIV), 2 FASTDIV and the z bin from two comparisons -- with
// the three z bins of utils_hist.py:65 the bin is a monotone step function of the difference, so it is fixed by the two
// differences at which the reference's formula steps (found on the host by bisection over the fp32 values, with the
// formula itself); the run of a row only holds partners that pass the z range test, so that test goes as well.
template <int kFusedThreads, int MODE, bool COUNT16>
__global__ void __launch_bounds__(kFusedThreads) hist_fused_kernel(FusedHistArgs a) {
    constexpr bool FASTDIV = MODE >= 1;
    constexpr bool ZSTEP = MODE == 2;
    ICPF_DYN_SHARED __align__(16) float4 fsm[];
    __shared__ float s_red[kFusedThreads / 32][12];
    __shared__ int s_cnt[2];
    __shared__ int s_bad;
    __shared__ unsigned long long s_best[kFusedThreads / 32];
    __shared__ unsigned long long s_pick[kFusedTopK];
    const int p = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float INF = __int_as_float(0x7f800000);
    const float4* xb = a.X + (size_t)p * a.N;
    const float4* yb = a.Y + (size_t)p * a.N;
    if (a.only_flagged && a.need_global[p] == 0) return;
    if (tid < 2) s_cnt[tid] = 0;
    if (tid == 0) s_bad = 0;
    __syncthreads();

    // ---- valid counts (for the swap rule) and bounding boxes of the flagged rows
    float lo[6] = {INF, INF, INF, INF, INF, INF}, hi[6] = {-INF, -INF, -INF, -INF, -INF, -INF};
    int cx = 0, cy = 0;
    for (int i = tid; i < a.N; i += kFusedThreads) {
        const float4 u = xb[i], v = yb[i];
        if (u.w > 0.f) {
            ++cx;
            lo[0] = fminf(lo[0], u.x); lo[1] = fminf(lo[1], u.y); lo[2] = fminf(lo[2], u.z);
            hi[0] = fmaxf(hi[0], u.x); hi[1] = fmaxf(hi[1], u.y); hi[2] = fmaxf(hi[2], u.z);
        }
        if (v.w > 0.f) {
            ++cy;
            lo[3] = fminf(lo[3], v.x); lo[4] = fminf(lo[4], v.y); lo[5] = fminf(lo[5], v.z);
            hi[3] = fmaxf(hi[3], v.x); hi[4] = fmaxf(hi[4], v.y); hi[5] = fmaxf(hi[5], v.z);
        }
    }
#pragma unroll
    for (int k = 0; k < 6; ++k) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo[k] = fminf(lo[k], __shfl_xor_sync(FULL_MASK, lo[k], o));
            hi[k] = fmaxf(hi[k], __shfl_xor_sync(FULL_MASK, hi[k], o));
        }
    }
    cx = __reduce_add_sync(FULL_MASK, cx);
    cy = __reduce_add_sync(FULL_MASK, cy);
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < 6; ++k) {
            s_red[warp][k] = lo[k];
            s_red[warp][6 + k] = hi[k];
        }
        atomicAdd(&s_cnt[0], cx);
        atomicAdd(&s_cnt[1], cy);
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 6; ++k) {
        lo[k] = s_red[0][k];
        hi[k] = s_red[0][6 + k];
        for (int w = 1; w < kFusedThreads / 32; ++w) {
            lo[k] = fminf(lo[k], s_red[w][k]);
            hi[k] = fmaxf(hi[k], s_red[w][6 + k]);
        }
    }
    // here X = dst and Y = src of the caller: swap when n_valid(src) > n_valid(dst)  (utils_match.py:139-146)
    int kx = 0, ky = 3;      // offsets of the X / Y boxes in lo[] / hi[]
    if (a.auto_swap && s_cnt[1] > s_cnt[0]) {
        const float4* t = xb; xb = yb; yb = t;
        kx = 3; ky = 0;
    }
    const float rx = __fsub_rn(a.max_x, a.min_x), ry = __fsub_rn(a.max_y, a.min_y), rz = __fsub_rn(a.max_z, a.min_z);
    const float flx = (float)a.len_x, fly = (float)a.len_y, flz = (float)a.len_z;
    const float irx = __frcp_rn(rx), iry = __frcp_rn(ry), irz = __frcp_rn(rz);
    // ---- bin range the differences can reach (x, y); +-1 bin of slack, clamped
    int bx0 = 0, bx1 = -1, by0 = 0, by1 = -1;
    {
        const float dx0 = lo[kx] - hi[ky], dx1 = hi[kx] - lo[ky];
        const float dy0 = lo[kx + 1] - hi[ky + 1], dy1 = hi[kx + 1] - lo[ky + 1];
        const float dz0 = lo[kx + 2] - hi[ky + 2], dz1 = hi[kx + 2] - lo[ky + 2];
        const bool any = (s_cnt[0] > 0) && (s_cnt[1] > 0) && (dx1 >= a.min_x) && (dx0 < a.max_x) && (dy1 >= a.min_y) &&
                         (dy0 < a.max_y) && (dz1 >= a.min_z) && (dz0 < a.max_z);
        if (any) {
            bx0 = max(0, vote_bin(fmaxf(dx0, a.min_x), a.min_x, rx, flx, a.len_x) - 1);
            bx1 = (dx1 >= a.max_x) ? a.len_x - 1 : min(a.len_x - 1, vote_bin(dx1, a.min_x, rx, flx, a.len_x) + 1);
            by0 = max(0, vote_bin(fmaxf(dy0, a.min_y), a.min_y, ry, fly, a.len_y) - 1);
            by1 = (dy1 >= a.max_y) ? a.len_y - 1 : min(a.len_y - 1, vote_bin(dy1, a.min_y, ry, fly, a.len_y) + 1);
        }
    }
    const int wx = bx1 - bx0 + 1, wy = by1 - by0 + 1, lz = a.len_z;
    const int ncol = (wx > 0 && wy > 0) ? wx * wy : 0;
    if (ncol > a.cap_cols) {
        if (tid == 0) a.need_global[p] = 1;
        return;
    }
    if (tid == 0) a.need_global[p] = 0;
    using cnt_t = typename std::conditional<COUNT16, unsigned short, unsigned int>::type;
    float4* tile = fsm;                                                          // [kFusedTile]
    unsigned int* histw = reinterpret_cast<unsigned int*>(fsm + kFusedTile);       // counters as 32-bit words
    const cnt_t* hist = reinterpret_cast<const cnt_t*>(histw);                     // [ncol * lz]
    const int hist_words = COUNT16 ? (a.cap_cols * lz + 1) / 2 : a.cap_cols * lz;
    cnt_t* colmax = reinterpret_cast<cnt_t*>(histw + hist_words);                  // [ncol] max over z
    cnt_t* rowmax = colmax + a.cap_cols + (a.cap_cols & 1);                        // [ncol] max over the y window
#if !defined(ICPF_SIMT_EMU)
    unsigned int hist_s;        // (an opaque copy: a known constant base is rematerialised inside the vote loop)
    asm volatile("mov.u32 %0, %1;" : "=r"(hist_s) : "r"((unsigned int)__cvta_generic_to_shared(histw)));
#endif
    for (int i = tid; i < (COUNT16 ? (ncol * lz + 1) / 2 : ncol * lz); i += kFusedThreads) histw[i] = 0u;
    __syncthreads();

    // ---- votes (hist_cuda_core.cuh:48-60 restated; shared-memory atomics).
    // Only |dz| < tau survives the z test -- about one pair in eight -- and fl(z_i - z_j) is monotone in z_j, so with
    // the Y tile sorted by z the partners of an X row are ONE contiguous run found by two binary searches on the exact
    // fp32 predicates; the inner loop then only touches pairs that pass the z test and every lane that iterates votes.
    if (ncol > 0) {
        for (int base = 0; base < a.N; base += kFusedTile) {
            const int n = min(kFusedTile, a.N - base);
            int npow = 1;
            while (npow < n) npow <<= 1;
            __syncthreads();
            for (int j = tid; j < npow; j += kFusedThreads) {
                float4 y = (j < n) ? yb[base + j] : make_float4(0.f, 0.f, INF, 0.f);
                // unflagged rows sort to the end and never match; neither does a NaN height (it fails the range test of
                // the reference, and the sort needs a total order)
                if (!(y.w > 0.f) || !(y.z == y.z)) y.z = INF;
                tile[j] = y;
            }
            __syncthreads();
            for (int k = 2; k <= npow; k <<= 1) {     // bitonic sort by z
                for (int st = k >> 1; st > 0; st >>= 1) {
                    for (int t = tid; t < npow; t += kFusedThreads) {
                        const int u = t ^ st;
                        if (u > t) {
                            const float4 p0 = tile[t], p1 = tile[u];
                            const bool up = (t & k) == 0;
                            if ((p0.z > p1.z) == up) { tile[t] = p1; tile[u] = p0; }
                        }
                    }
                    __syncthreads();
                }
            }
            // Each warp takes 32 consecutive X rows: every lane finds the run of ITS row (two binary searches, the same
            // trip count in every lane), then the rows are served one after the other with the run of the row dealt to
            // the 32 lanes -- runs differ a lot in length, and a lane that iterated over its own run alone left the warp
            // at 19 of 32 active lanes (profiles/hist_fused_r1_by_line.txt); consecutive lanes now also read consecutive
            // tile rows (no bank conflicts).  Votes are integer atomics: the counts do not depend on the order.
            for (int i0 = warp * 32; i0 < a.N; i0 += kFusedThreads) {
                const int i = i0 + lane;
                const float4 xi = (i < a.N) ? xb[i] : make_float4(0.f, 0.f, 0.f, 0.f);
                int j0 = 0, j1 = 0;
                if (xi.w > 0.f) {
                    // first j with  fl(z_i - z_j) <  max_z   (predicate false..false true..true as z_j grows)
                    int lo_j = 0, hi_j = n;
                    while (lo_j < hi_j) {
                        const int mid = (lo_j + hi_j) >> 1;
                        if (__fsub_rn(xi.z, tile[mid].z) < a.max_z) hi_j = mid; else lo_j = mid + 1;
                    }
                    j0 = lo_j;
                    // first j with  fl(z_i - z_j) >= min_z  false  (true..true false..false)
                    hi_j = n;
                    while (lo_j < hi_j) {
                        const int mid = (lo_j + hi_j) >> 1;
                        if (__fsub_rn(xi.z, tile[mid].z) >= a.min_z) lo_j = mid + 1; else hi_j = mid;
                    }
                    j1 = lo_j;
                }
                unsigned int live = __ballot_sync(FULL_MASK, j1 > j0);
                while (live != 0u) {
                    const int r = __ffs((int)live) - 1;
                    live &= live - 1u;
                    const float xix = __shfl_sync(FULL_MASK, xi.x, r), xiy = __shfl_sync(FULL_MASK, xi.y, r),
                                xiz = __shfl_sync(FULL_MASK, xi.z, r);
                    const int rs = __shfl_sync(FULL_MASK, j0, r), re = __shfl_sync(FULL_MASK, j1, r);
                    for (int j = rs + lane; j < re; j += 32) {
                        const float4 yj = tile[j];
                        const float vz = __fsub_rn(xiz, yj.z);
                        const float vx = __fsub_rn(xix, yj.x), vy = __fsub_rn(xiy, yj.y);
                        // (the run [rs, re) was cut with the exact z predicates: min_z <= vz < max_z holds inside it)
                        if (vx >= a.min_x && vx < a.max_x && vy >= a.min_y && vy < a.max_y &&
                            (ZSTEP || (vz >= a.min_z && vz < a.max_z))) {
                            const int px = (FASTDIV ? vote_bin_fast(vx, a.min_x, rx, irx, flx, a.len_x)
                                                    : vote_bin(vx, a.min_x, rx, flx, a.len_x)) - bx0;
                            const int py = (FASTDIV ? vote_bin_fast(vy, a.min_y, ry, iry, fly, a.len_y)
                                                    : vote_bin(vy, a.min_y, ry, fly, a.len_y)) - by0;
                            const int pz = ZSTEP ? (vz >= a.zt1 ? 1 : 0) + (vz >= a.zt2 ? 1 : 0)
                                                 : (FASTDIV ? vote_bin_fast(vz, a.min_z, rz, irz, flz, a.len_z)
                                                            : vote_bin(vz, a.min_z, rz, flz, a.len_z));
                            if ((unsigned int)px >= (unsigned int)wx || (unsigned int)py >= (unsigned int)wy) {
                                s_bad = 1;      // cannot happen (the range is conservative); fall back if it ever does
                            } else {
                                const int bin = (px * wy + py) * lz + pz;
                                if (COUNT16) {
                                    const int sh = (bin & 1) * 16;
                                    const unsigned int was = atomicAdd(&histw[bin >> 1], 1u << sh);
                                    if (((was >> sh) & 0xffffu) == 0xffffu) s_bad = 1;      // this increment wrapped it
                                } else {
#if defined(ICPF_SIMT_EMU) || defined(ICPF_NO_RED_PTX)
                                    atomicAdd(&histw[bin], 1u);
#else
                                    // (a 32-bit shared address kept in a register: the generic pointer made the compiler
                                    //  rebuild the shared window base -- five instructions -- in every iteration)
                                    asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(hist_s + 4u * (unsigned int)bin) : "memory");
#endif
                                }
                            }
                        }
                    }
                }
            }
        }
    }
    __syncthreads();
    if (s_bad) {
        if (tid == 0) a.need_global[p] = 1;
        return;
    }

    // ---- 11^3 max-pool (separable; the z extent is always inside the window) + survivors == window max
    for (int c = tid; c < ncol; c += kFusedThreads) {
        unsigned int m = hist[c * lz];
        for (int z = 1; z < lz; ++z) m = max(m, (unsigned int)hist[c * lz + z]);
        colmax[c] = (cnt_t)m;
    }
    __syncthreads();
    for (int c = tid; c < ncol; c += kFusedThreads) {
        const int x = c / wy, y = c - x * wy;
        unsigned int m = colmax[c];
        for (int d = max(0, y - kFusedNmsHalf); d <= min(wy - 1, y + kFusedNmsHalf); ++d)
            m = max(m, (unsigned int)colmax[x * wy + d]);
        rowmax[c] = (cnt_t)m;
    }
    __syncthreads();
    unsigned long long top[kFusedTopK];
#pragma unroll
    for (int k = 0; k < kFusedTopK; ++k) top[k] = 0ull;
    for (int c = tid; c < ncol; c += kFusedThreads) {
        const int x = c / wy, y = c - x * wy;
        unsigned int mi = rowmax[c];
        for (int d = max(0, x - kFusedNmsHalf); d <= min(wx - 1, x + kFusedNmsHalf); ++d)
            mi = max(mi, (unsigned int)rowmax[d * wy + y]);
        if (mi == 0u) continue;
        for (int z = 0; z < lz; ++z) {
            const float v = (float)hist[c * lz + z];
            if ((unsigned int)hist[c * lz + z] == mi) {
                // flat index in the FULL volume: ties rank by it, exactly as on the global path
                unsigned long long key = fused_peak_key(v, ((x + bx0) * a.len_y + (y + by0)) * lz + z);
#pragma unroll
                for (int k = 0; k < kFusedTopK; ++k) {
                    if (key > top[k]) { const unsigned long long t = top[k]; top[k] = key; key = t; }
                }
            }
        }
    }
    for (int round = 0; round < kFusedTopK; ++round) {
        unsigned long long best = top[0];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const unsigned long long other = __shfl_xor_sync(FULL_MASK, best, o);
            best = other > best ? other : best;
        }
        if (lane == 0) s_best[warp] = best;
        __syncthreads();
        if (tid == 0) {
            unsigned long long m = s_best[0];
            for (int w = 1; w < kFusedThreads / 32; ++w) m = s_best[w] > m ? s_best[w] : m;
            s_pick[round] = m;
        }
        __syncthreads();
        const unsigned long long win = s_pick[round];
        if (win != 0ull && top[0] == win) {
#pragma unroll
            for (int k = 0; k + 1 < kFusedTopK; ++k) top[k] = top[k + 1];
            top[kFusedTopK - 1] = 0ull;
        }
        __syncthreads();
    }
    if (tid == 0) {
        int picked[kFusedTopK];
        int np = 0;
        for (int k = 0; k < kFusedTopK; ++k) {
            if (s_pick[k] != 0ull) {
                picked[np] = 0x7fffffff - (int)(unsigned int)(s_pick[k] & 0xffffffffu);
                a.out_votes[(size_t)p * kFusedTopK + np] = __uint_as_float((unsigned int)(s_pick[k] >> 32));
                ++np;
            }
        }
        int fill = 0;
        const int npos = np;
        while (np < kFusedTopK) {       // zero-vote fillers: lowest flat indices (see icpf_hist.cu)
            bool used = false;
            for (int k = 0; k < npos; ++k) used = used || (picked[k] == fill);
            if (!used) {
                picked[np] = fill;
                a.out_votes[(size_t)p * kFusedTopK + np] = 0.f;
                ++np;
            }
            ++fill;
        }
        for (int k = 0; k < kFusedTopK; ++k) a.out_idx[(size_t)p * kFusedTopK + k] = picked[k];
    }
}

// host restatement of vote_bin() (every step rounded to fp32) and the order-preserving map float <-> u32
static int host_vote_bin(float v, float mn, float range, float flen, int len) {
    volatile float x = v - mn;
    volatile float q = x / range;
    volatile float m = q * flen;
    const int p = (int)floorf(m);
    return p < len - 1 ? p : len - 1;
}
static uint32_t ordered_key(float f) {
    uint32_t u;
    memcpy(&u, &f, 4);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
static float from_ordered_key(uint32_t k) {
    const uint32_t u = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
    float f;
    memcpy(&f, &u, 4);
    return f;
}

int launch_hist_fused(const float* X, const float* Y, int P, int N, const float* mins, const float* maxs,
                      const int* lens, int auto_swap, int* out_idx, float* out_votes, int* need_global,
                      cudaStream_t stream) {
    if (P == 0) return ICPF_OK;
    const size_t budget = 200 * 1024;
    const size_t fixed = (size_t)kFusedTile * 16;
    const bool wide = N >= 512;
    // Markstein's division needs a divisor whose significand is not all ones and quotients / remainders far from the
    // under- and overflow thresholds (dividends are differences of coordinates, |x| < range)
    bool fastdiv = true;
    for (int k = 0; k < 3; ++k) {
        const float r = maxs[k] - mins[k];
        uint32_t u;
        memcpy(&u, &r, 4);
        const int e = (int)((u >> 23) & 0xffu);
        fastdiv = fastdiv && (u & 0x7fffffu) != 0x7fffffu && e > 127 - 40 && e < 127 + 40;
    }
    // three z bins: the differences at which the reference's bin formula steps from bin k - 1 to bin k (MODE 2)
    float zt[2] = {0.f, 0.f};
    const bool zstep = fastdiv && lens[2] == 3 && maxs[2] > mins[2];
    if (zstep) {
        const float INF = std::numeric_limits<float>::infinity();
        for (int k = 1; k <= 2; ++k) {
            uint32_t lo = ordered_key(mins[2]), hi = ordered_key(maxs[2]);      // first key in [lo, hi) with bin >= k
            while (lo < hi) {
                const uint32_t mid = lo + (hi - lo) / 2;
                if (host_vote_bin(from_ordered_key(mid), mins[2], maxs[2] - mins[2], 3.f, 3) >= k) hi = mid; else lo = mid + 1;
            }
            zt[k - 1] = (lo == ordered_key(maxs[2])) ? INF : from_ordered_key(lo);
        }
    }
#ifndef ICPF_FUSED_WIDE
#define ICPF_FUSED_WIDE 1024
#endif
    // tier 1: u32 counters + u32 max planes (4 lz + 8 bytes per column); tier 2, only for the pairs tier 1 flagged:
    // u16 (2 lz + 4 bytes per column, twice the columns); whatever is still flagged takes the global-memory kernels
    for (int tier = 0; tier < 2; ++tier) {
        const size_t per_col = tier == 0 ? (size_t)lens[2] * 4 + 8 : (size_t)lens[2] * 2 + 4;
        int cap_cols = (int)((budget - fixed - 16) / per_col);
        if (cap_cols > lens[0] * lens[1]) cap_cols = lens[0] * lens[1];
        const size_t smem = fixed + 16 +
                            (tier == 0 ? (size_t)cap_cols * lens[2] * 4 + (size_t)(cap_cols + (cap_cols & 1)) * 8
                                       : ((size_t)(cap_cols * lens[2] + 1) / 2) * 4 + (size_t)(cap_cols + (cap_cols & 1)) * 4);
        void (*kernel)(FusedHistArgs);
        const int mode = zstep ? 2 : (fastdiv ? 1 : 0);
        if (tier == 0)
            kernel = wide ? (mode == 2 ? hist_fused_kernel<ICPF_FUSED_WIDE, 2, false>
                                       : mode == 1 ? hist_fused_kernel<ICPF_FUSED_WIDE, 1, false> : hist_fused_kernel<ICPF_FUSED_WIDE, 0, false>)
                          : (mode == 2 ? hist_fused_kernel<256, 2, false>
                                       : mode == 1 ? hist_fused_kernel<256, 1, false> : hist_fused_kernel<256, 0, false>);
        else
            kernel = wide ? (mode == 2 ? hist_fused_kernel<ICPF_FUSED_WIDE, 2, true>
                                       : mode == 1 ? hist_fused_kernel<ICPF_FUSED_WIDE, 1, true> : hist_fused_kernel<ICPF_FUSED_WIDE, 0, true>)
                          : (mode == 2 ? hist_fused_kernel<256, 2, true>
                                       : mode == 1 ? hist_fused_kernel<256, 1, true> : hist_fused_kernel<256, 0, true>);
        cudaError_t err = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (err != cudaSuccess) return (int)err;
        FusedHistArgs a{reinterpret_cast<const float4*>(X), reinterpret_cast<const float4*>(Y), N,
                        mins[0], mins[1], mins[2], maxs[0], maxs[1], maxs[2], lens[0], lens[1], lens[2],
                        auto_swap, cap_cols, out_idx, out_votes, need_global, tier, zt[0], zt[1]};
        ICPF_LAUNCH(kernel, P, wide ? ICPF_FUSED_WIDE : 256, smem, stream)(a);
        err = cudaGetLastError();
        if (err != cudaSuccess) return (int)err;
    }
    return ICPF_OK;
}

}  // namespace icpf
