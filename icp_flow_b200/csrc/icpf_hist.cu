// icpf_hist.cu -- histogram-vote translation initialisation (utils_hist.py:33-124) as three stream-ordered kernels:
//
//   hist_votes_kernel   all-pairs difference histogram, bit-compatible with the reference's only CUDA kernel
//                       (hist_cuda/cpp/hist_cuda_core.cuh:35-62): one vote per (X_i, Y_j) with both flags > 0 whose
//                       difference lies in the half-open range, bin = floor((v - min) / (max - min) * len) in fp32
//   hist_peaks_kernel   3-D non-maximum suppression (11^3 window == all three z bins) + top-5 (utils_hist.py:21-29)
//   hist_score_kernel   the 5 decoded translations + the zero translation, scored by the smaller of the two mean
//                       unbounded NN distances, arg-min -> 4x4 translation pose (utils_hist.py:78-122)
//
// X rows are staged through shared memory (one coalesced pass over HBM); votes are fp32 atomic adds of 1.0 into an
// L2-resident histogram (exact: counts stay far below 2^24), cheap rejects first (|dz| < tau discards most pairs).
#include "icpf_internal.h"
#include "icpf_pair.cuh"
#include "icpf_gridnn.cuh"

namespace icpf {

// ------------------------------------------------------------------------------------------------ votes
constexpr int kVoteThreads = 256;
constexpr int kVoteTile = 1024;   // Y rows per shared-memory tile

struct HistGeom {
    float min_x, min_y, min_z, max_x, max_y, max_z;
    int len_x, len_y, len_z;
};

// hist(X, Y): votes X_i - Y_j.  If `auto_swap`, the roles follow utils_match.py:139-146 / utils_hist.py:69, i.e. the
// call is hist(dst_, src_) with (src_, dst_) = (src, dst) swapped whenever src has more valid rows than dst.
__global__ void __launch_bounds__(kVoteThreads) hist_votes_kernel(const float4* __restrict__ X,
                                                                  const float4* __restrict__ Y, int NX, int NY,
                                                                  HistGeom gm, float* __restrict__ bins, int auto_swap,
                                                                  const int* __restrict__ need) {
    __shared__ float4 tile[kVoteTile];
    __shared__ int s_cnt[2];
    const int b = blockIdx.y;
    if (need != nullptr && need[b] == 0) return;      // this pair was handled by the fused shared-memory kernel
    const float4* xb = X + (size_t)b * NX;
    const float4* yb = Y + (size_t)b * NY;
    int nx = NX, ny = NY;
    if (auto_swap) {
        // here X = dst and Y = src of the caller: swap when n_valid(src) > n_valid(dst)
        if (threadIdx.x < 2) s_cnt[threadIdx.x] = 0;
        __syncthreads();
        int cx = 0, cy = 0;
        for (int i = threadIdx.x; i < NX; i += kVoteThreads) cx += xb[i].w > 0.f;
        for (int i = threadIdx.x; i < NY; i += kVoteThreads) cy += yb[i].w > 0.f;
        cx = __reduce_add_sync(FULL_MASK, cx);
        cy = __reduce_add_sync(FULL_MASK, cy);
        if ((threadIdx.x & 31) == 0) {
            atomicAdd(&s_cnt[0], cx);
            atomicAdd(&s_cnt[1], cy);
        }
        __syncthreads();
        if (s_cnt[1] > s_cnt[0]) {
            const float4* t = xb; xb = yb; yb = t;
            nx = NY; ny = NX;
        }
        __syncthreads();
    }
    float* hb = bins + (size_t)b * gm.len_x * gm.len_y * gm.len_z;
    const float rx = __fsub_rn(gm.max_x, gm.min_x), ry = __fsub_rn(gm.max_y, gm.min_y),
                rz = __fsub_rn(gm.max_z, gm.min_z);
    const float flx = (float)gm.len_x, fly = (float)gm.len_y, flz = (float)gm.len_z;
    const int i = blockIdx.x * kVoteThreads + threadIdx.x;
    float4 xi = make_float4(0.f, 0.f, 0.f, 0.f);
    if (i < nx) xi = xb[i];
    const bool live = (i < nx) && (xi.w > 0.f);
    for (int base = 0; base < ny; base += kVoteTile) {
        const int n = min(kVoteTile, ny - base);
        __syncthreads();
        for (int j = threadIdx.x; j < n; j += kVoteThreads) tile[j] = yb[base + j];
        __syncthreads();
        if (!live) continue;
        for (int j = 0; j < n; ++j) {
            const float4 yj = tile[j];
            const float vz = __fsub_rn(xi.z, yj.z);
            if (!(yj.w > 0.f) || !(vz >= gm.min_z && vz < gm.max_z)) continue;
            const float vx = __fsub_rn(xi.x, yj.x), vy = __fsub_rn(xi.y, yj.y);
            if (vx >= gm.min_x && vx < gm.max_x && vy >= gm.min_y && vy < gm.max_y) {
                int px = __float2int_rd(__fmul_rn(__fdiv_rn(__fsub_rn(vx, gm.min_x), rx), flx));
                int py = __float2int_rd(__fmul_rn(__fdiv_rn(__fsub_rn(vy, gm.min_y), ry), fly));
                int pz = __float2int_rd(__fmul_rn(__fdiv_rn(__fsub_rn(vz, gm.min_z), rz), flz));
                // (v - min) can round up to (max - min) for v one ulp below max; the reference would then write out
                // of the pair's histogram -- clamp instead
                px = min(px, gm.len_x - 1); py = min(py, gm.len_y - 1); pz = min(pz, gm.len_z - 1);
                atomicAdd(hb + ((size_t)px * gm.len_y + py) * gm.len_z + pz, 1.0f);
            }
        }
    }
}

__global__ void __launch_bounds__(256) hist_zero_flagged_kernel(float* __restrict__ bins, size_t per_pair, const int* __restrict__ need) {
    if (need[blockIdx.x] == 0) return;
    float* h = bins + (size_t)blockIdx.x * per_pair;
    for (size_t i = threadIdx.x; i < per_pair; i += blockDim.x) h[i] = 0.f;
}

int launch_hist_votes(const float* X, const float* Y, int B, int NX, int NY, const float* mins, const float* maxs,
                      const int* lens, float* bins, int auto_swap, const int* need, cudaStream_t stream) {
    if (B == 0) return ICPF_OK;
    HistGeom gm{mins[0], mins[1], mins[2], maxs[0], maxs[1], maxs[2], lens[0], lens[1], lens[2]};
    const size_t per_pair = (size_t)lens[0] * lens[1] * lens[2];
    if (need == nullptr) {
        cudaError_t err = cudaMemsetAsync(bins, 0, (size_t)B * per_pair * sizeof(float), stream);    // at::zeros of hist_cuda.cu:59
        if (err != cudaSuccess) return (int)err;
    } else {
        // fall-back after the fused kernel: only the flagged pairs vote here, so only their histograms are zero-filled
        // (a chunk is up to 64 MB; on a batch the fused kernel handled completely this is B CTAs that return at once)
        ICPF_LAUNCH(hist_zero_flagged_kernel, B, 256, 0, stream)(bins, per_pair, need);
    }
    const int nmax = auto_swap ? max(NX, NY) : NX;
    dim3 grid((nmax + kVoteThreads - 1) / kVoteThreads, B);
    ICPF_LAUNCH(hist_votes_kernel, grid, kVoteThreads, 0, stream)(reinterpret_cast<const float4*>(X),
                                                         reinterpret_cast<const float4*>(Y), NX, NY, gm, bins,
                                                         auto_swap, need);
    return (int)cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------ NMS + top-k
constexpr int kPeakThreads = 256;
constexpr int kTopK = 5;        // utils_hist.py:21
constexpr int kNmsHalf = 5;     // kernel_size 11 -> +-5 bins; the z extent (3 bins) is always inside the window

// key = (vote count bits, ~flat index): larger = more votes, ties -> lowest flat index (torch.topk's CPU order)
__device__ __forceinline__ unsigned long long peak_key(float v, int idx) {
    return ((unsigned long long)__float_as_uint(v) << 32) | (unsigned int)(0x7fffffff - idx);
}

// `scratch` (may be NULL): [B, 2, lx*ly] floats in global memory for windows whose two max planes do not fit shared
// memory (translation_frame beyond ~8 m: the reference widens it with the frame gap, main.py:200).
__global__ void __launch_bounds__(kPeakThreads) hist_peaks_kernel(const float* __restrict__ bins, int lx, int ly,
                                                                  int lz, int* __restrict__ out_idx,
                                                                  float* __restrict__ out_votes,
                                                                  const int* __restrict__ need,
                                                                  float* __restrict__ scratch) {
    ICPF_DYN_SHARED float sm[];
    __shared__ unsigned long long s_best[kPeakThreads / 32];
    __shared__ unsigned long long s_pick[kTopK];
    const int b = blockIdx.x, tid = threadIdx.x;
    if (need != nullptr && need[b] == 0) return;      // handled by the fused kernel
    float* colmax = scratch ? scratch + (size_t)b * 2 * lx * ly : sm;      // [lx*ly] max over z
    float* rowmax = colmax + (size_t)lx * ly;                               // [lx*ly] max over the y window
    const float* hb = bins + (size_t)b * lx * ly * lz;
    const int ncol = lx * ly;
    for (int c = tid; c < ncol; c += kPeakThreads) {
        float m = hb[(size_t)c * lz];
        for (int z = 1; z < lz; ++z) m = fmaxf(m, hb[(size_t)c * lz + z]);
        colmax[c] = m;
    }
    __syncthreads();
    for (int c = tid; c < ncol; c += kPeakThreads) {
        const int x = c / ly, y = c - x * ly;
        float m = colmax[c];
        for (int d = max(0, y - kNmsHalf); d <= min(ly - 1, y + kNmsHalf); ++d) m = fmaxf(m, colmax[x * ly + d]);
        rowmax[c] = m;
    }
    __syncthreads();
    // each thread keeps its own top-k of surviving positive bins (value desc, index asc)
    unsigned long long top[kTopK];
#pragma unroll
    for (int k = 0; k < kTopK; ++k) top[k] = 0ull;
    for (int c = tid; c < ncol; c += kPeakThreads) {
        const int x = c / ly, y = c - x * ly;
        float m = rowmax[c];
        for (int d = max(0, x - kNmsHalf); d <= min(lx - 1, x + kNmsHalf); ++d) m = fmaxf(m, rowmax[d * ly + y]);
        if (!(m > 0.f)) continue;
        for (int z = 0; z < lz; ++z) {
            const float v = hb[(size_t)c * lz + z];
            if (v == m) {          // x == max_pool3d(x): plateaus all survive
                unsigned long long key = peak_key(v, c * lz + z);
#pragma unroll
                for (int k = 0; k < kTopK; ++k) {
                    if (key > top[k]) { const unsigned long long t = top[k]; top[k] = key; key = t; }
                }
            }
        }
    }
    // k rounds of block arg-max over the heads of the per-thread lists
    for (int round = 0; round < kTopK; ++round) {
        unsigned long long best = top[0];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const unsigned long long other = __shfl_xor_sync(FULL_MASK, best, o);
            best = other > best ? other : best;
        }
        if ((tid & 31) == 0) s_best[tid >> 5] = best;
        __syncthreads();
        if (tid == 0) {
            unsigned long long m = s_best[0];
            for (int w = 1; w < kPeakThreads / 32; ++w) m = s_best[w] > m ? s_best[w] : m;
            s_pick[round] = m;
        }
        __syncthreads();
        const unsigned long long win = s_pick[round];
        if (win != 0ull && top[0] == win) {      // keys are unique (they embed the flat index): pop it
#pragma unroll
            for (int k = 0; k + 1 < kTopK; ++k) top[k] = top[k + 1];
            top[kTopK - 1] = 0ull;
        }
        __syncthreads();
    }
    if (tid == 0) {
        // fewer than k positive peaks: torch.topk pads with zero-valued entries of x * mask; those are arbitrary
        // zero-vote bins in the reference (implementation-defined order) -- take the lowest flat indices
        int picked[kTopK];
        int np = 0;
        for (int k = 0; k < kTopK; ++k) {
            if (s_pick[k] != 0ull) {
                picked[np] = 0x7fffffff - (int)(unsigned int)(s_pick[k] & 0xffffffffu);
                out_votes[(size_t)b * kTopK + np] = __uint_as_float((unsigned int)(s_pick[k] >> 32));
                ++np;
            }
        }
        int fill = 0;
        const int npos = np;
        while (np < kTopK) {
            bool used = false;
            for (int k = 0; k < npos; ++k) used = used || (picked[k] == fill);
            if (!used) {
                picked[np] = fill;
                out_votes[(size_t)b * kTopK + np] = 0.f;
                ++np;
            }
            ++fill;
        }
        for (int k = 0; k < kTopK; ++k) out_idx[(size_t)b * kTopK + k] = picked[k];
    }
}

size_t hist_peaks_scratch_floats(int lx, int ly) {
    return ((size_t)lx * ly * 2 * sizeof(float) > (size_t)200 * 1024) ? (size_t)lx * ly * 2 : 0;
}

int launch_hist_peaks(const float* bins, int B, int lx, int ly, int lz, int* out_idx, float* out_votes,
                      const int* need, float* scratch, cudaStream_t stream) {
    if (B == 0) return ICPF_OK;
    const bool global = hist_peaks_scratch_floats(lx, ly) > 0;
    if (global && scratch == nullptr) return ICPF_E_WORKSPACE;
    const size_t smem = global ? 0 : (size_t)lx * ly * 2 * sizeof(float);
    cudaError_t err = cudaFuncSetAttribute(hist_peaks_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess) return (int)err;
    ICPF_LAUNCH(hist_peaks_kernel, B, kPeakThreads, smem, stream)(bins, lx, ly, lz, out_idx, out_votes, need,
                                                         global ? scratch : nullptr);
    return (int)cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------ candidate scoring
constexpr int kCand = kTopK + 1;   // + the zero translation, last (utils_hist.py:83)

struct ScoreArgs {
    const float* src;        // [P,N,4]
    const float* dst;        // [P,N,4]
    int N;
    const int* cand_idx;     // [P,5] flat histogram indices
    const float* bins_x;     // [lx] bin starts (torch.arange of the reference, utils_hist.py:63-65)
    const float* bins_y;     // [ly]
    const float* bins_z;     // [lz]
    int lx, ly, lz;
    float half_bin;          // args.thres_dist // 2  (utils_hist.py:78; 0.0 for tau = 0.1)
    float tau;               // bin width = thres_dist: only sizes the cells of the NN grids
    int auto_swap;
    float* out_pose;         // [P,16]
    float* out_scores;       // [P,6] (may be NULL)
    int* out_which;          // [P]   (may be NULL)
    int* defer;              // deferral scratch (score_defer_words(P) words) or NULL: see hist_score_kernel
    int P;
};

// Deferral scratch, in 32-bit words: [0] items queued, [1] items taken, [64 ..) the queue (p * 8 + k, at most 5 per
// pair), then 16 words per pair: scores[6], best score of the first kernel, evaluated-mask, items still pending.
constexpr int kSmCount = 148;        // B200
constexpr int kDeferHeader = 64;
constexpr int kDeferPairWords = 16;
#ifndef ICPF_DEFER_BUDGET
#define ICPF_DEFER_BUDGET 2
#endif
constexpr int kDeferBudget = ICPF_DEFER_BUDGET;      // chunks of kThreads rows a candidate may cost in the first kernel (power of two; 0: no deferral)
size_t score_defer_words(int P) { return (size_t)kDeferHeader + (size_t)P * (kTopK + kDeferPairWords); }

// errors.min(dim=-1) over the evaluated candidates (first minimum) -> 4x4 translation pose (utils_hist.py:104-122)
__device__ __forceinline__ void score_select(const ScoreArgs& a, int p, const float (&score)[kCand], unsigned int evalmask,
                                             const float (*t)[3]) {
    const float INF = __int_as_float(0x7f800000);
    int which = 0;
    float best = 0.f;
    bool have = false;
    for (int k = 0; k < kCand; ++k) {
        const bool ev = (evalmask >> k) & 1u;
        const float e = ev ? score[k] : INF;        // +inf for the candidates that were excluded
        if (a.out_scores) a.out_scores[(size_t)p * kCand + k] = e;
        if (!ev) continue;
        if (!have || e < best) { best = e; which = k; have = true; }      // errors.min(dim=-1): first minimum
    }
    if (a.out_which) a.out_which[p] = which;
    float* o = a.out_pose + (size_t)p * 16;
    for (int i = 0; i < 16; ++i) o[i] = (i % 5 == 0) ? 1.f : 0.f;
    o[3] = t[which][0]; o[7] = t[which][1]; o[11] = t[which][2];
}

// GRIDNN: both clouds are counting-sorted into uniform grids in shared memory (icpf_gridnn.cuh) and the exact scores
// come from grid searches instead of full scans -- the same minima, hence the same bits, at O(n) instead of O(n^2).
// The rows in storage order (they fix the order of the sums) are streamed from global memory, coalesced and
// L2-resident, so a pair costs 2 (N + 257) * 16 B of shared memory (N = 1024: 41 KB, 5 CTAs per SM).
// !GRIDNN (clusters whose grids do not fit shared memory, N > ~7000): full scans over the rows in global memory.
//
// Deferral (GRIDNN, a.defer != NULL).  A pair of unrelated clusters has no candidate that ends the others early: all six
// are evaluated in full, about 30 times the work of a matching pair, and as ONE CTA per pair those few pairs were the
// whole tail of the launch (C3: 5 % of the pairs, 3.4 of 3.8 ms).  So the first kernel gives every candidate after the
// top peak a budget of kDeferBudget chunks per direction; a candidate that is neither finished nor excluded by then is
// queued as an item (pair, candidate) and the pair's selection is left open.  The second kernel (ITEM, persistent CTAs
// pulling from the queue) evaluates one item per CTA -- the candidates of a slow pair side by side on up to five SMs --
// against the best score of the first kernel, and whichever CTA finishes a pair's last item selects.  A candidate is
// either exact (same rows, same order of the sums: the same bits however it was scheduled) or provably above an exact
// score, so the selected translation does not depend on the split; which excluded candidates read +inf does not depend
// on timing either (the limit of an item is fixed by the first kernel).
template <bool GRIDNN, bool ITEM>
__device__ __forceinline__ void score_pair(const ScoreArgs& a, const int p, const int only_k) {
    __shared__ float s_t[kCand][3];
    __shared__ float s_part[kWarps][2][kCand];
    __shared__ float s_scratch[kWarps * 4];
    __shared__ float s_box[kWarps][12];
    __shared__ float s_lb[kCand];
    __shared__ float s_score[kCand];
    const int tid = threadIdx.x;
    // this pair's words of the deferral scratch
    int* pairw = (GRIDNN && a.defer != nullptr) ? a.defer + kDeferHeader + (size_t)a.P * kTopK + (size_t)p * kDeferPairWords
                                                : nullptr;
    const float4* S = reinterpret_cast<const float4*>(a.src) + (size_t)p * a.N;
    const float4* D = reinterpret_cast<const float4*>(a.dst) + (size_t)p * a.N;
    float cnt[2] = {0.f, 0.f};
    for (int q = tid; q < a.N; q += kThreads) {
        cnt[0] += (S[q].w > 0.f) ? 1.f : 0.f;
        cnt[1] += (D[q].w > 0.f) ? 1.f : 0.f;
    }
    block_allreduce_sum<2, kWarps>(cnt, s_scratch);
    int n_s = (int)cnt[0], n_d = (int)cnt[1];
    if (a.auto_swap && n_s > n_d) {     // always register the smaller cloud onto the larger one (utils_match.py:139-146)
        const float4* t = S; S = D; D = t;
        const int n = n_s; n_s = n_d; n_d = n;
    }
    if (tid < kCand) {
        float tx = 0.f, ty = 0.f, tz = 0.f;
        if (tid < kTopK) {
            // flat -> (i // d // w % h, i // d % w, i % d)  (utils_hist.py:78)
            const int flat = a.cand_idx[(size_t)p * kTopK + tid];
            const int iz = flat % a.lz, iy = (flat / a.lz) % a.ly, ix = (flat / a.lz / a.ly) % a.lx;
            tx = __fadd_rn(a.bins_x[ix], a.half_bin);
            ty = __fadd_rn(a.bins_y[iy], a.half_bin);
            tz = __fadd_rn(a.bins_z[iz], a.half_bin);
        }
        s_t[tid][0] = tx; s_t[tid][1] = ty; s_t[tid][2] = tz;
    }
    __syncthreads();
    const float INF = __int_as_float(0x7f800000);

    // ---- a cheap lower bound of every candidate's score: a point's NN distance is at least its distance to the
    // other cloud's bounding box, so  score_k >= min(mean_i dist(src_i + t_k, bbox(dst)), mean_i dist(dst_i - t_k,
    // bbox(src))).  Candidates whose bound exceeds an exactly evaluated score can never be the arg-min and are skipped
    // (the zero-vote "filler" bins of pairs with fewer than five peaks sit metres away: result-identical, O(n) instead
    // of O(n^2) for most candidates on real frames).
    if constexpr (!ITEM) {
    float blo[6] = {INF, INF, INF, INF, INF, INF}, bhi[6] = {-INF, -INF, -INF, -INF, -INF, -INF};
    for (int i = tid; i < n_s; i += kThreads) {
        const float4 v = S[i];
        blo[0] = fminf(blo[0], v.x); blo[1] = fminf(blo[1], v.y); blo[2] = fminf(blo[2], v.z);
        bhi[0] = fmaxf(bhi[0], v.x); bhi[1] = fmaxf(bhi[1], v.y); bhi[2] = fmaxf(bhi[2], v.z);
    }
    for (int i = tid; i < n_d; i += kThreads) {
        const float4 v = D[i];
        blo[3] = fminf(blo[3], v.x); blo[4] = fminf(blo[4], v.y); blo[5] = fminf(blo[5], v.z);
        bhi[3] = fmaxf(bhi[3], v.x); bhi[4] = fmaxf(bhi[4], v.y); bhi[5] = fmaxf(bhi[5], v.z);
    }
#pragma unroll
    for (int k = 0; k < 6; ++k) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            blo[k] = fminf(blo[k], __shfl_xor_sync(FULL_MASK, blo[k], o));
            bhi[k] = fmaxf(bhi[k], __shfl_xor_sync(FULL_MASK, bhi[k], o));
        }
    }
    if ((tid & 31) == 0) {
#pragma unroll
        for (int k = 0; k < 6; ++k) {
            s_box[tid >> 5][k] = blo[k];
            s_box[tid >> 5][6 + k] = bhi[k];
        }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 6; ++k) {
        blo[k] = s_box[0][k];
        bhi[k] = s_box[0][6 + k];
        for (int w = 1; w < kWarps; ++w) {
            blo[k] = fminf(blo[k], s_box[w][k]);
            bhi[k] = fmaxf(bhi[k], s_box[w][6 + k]);
        }
    }
    {
        float lf[kCand], lb[kCand];
#pragma unroll
        for (int k = 0; k < kCand; ++k) lf[k] = lb[k] = 0.f;
        for (int i = tid; i < n_s; i += kThreads) {
            const float4 v = S[i];
#pragma unroll
            for (int k = 0; k < kCand; ++k) {
                const float x = v.x + s_t[k][0], y = v.y + s_t[k][1], z = v.z + s_t[k][2];
                const float dx = fmaxf(fmaxf(blo[3] - x, x - bhi[3]), 0.f), dy = fmaxf(fmaxf(blo[4] - y, y - bhi[4]), 0.f),
                            dz = fmaxf(fmaxf(blo[5] - z, z - bhi[5]), 0.f);
                lf[k] += sqrtf(dx * dx + dy * dy + dz * dz);
            }
        }
        for (int i = tid; i < n_d; i += kThreads) {
            const float4 v = D[i];
#pragma unroll
            for (int k = 0; k < kCand; ++k) {
                const float x = v.x - s_t[k][0], y = v.y - s_t[k][1], z = v.z - s_t[k][2];
                const float dx = fmaxf(fmaxf(blo[0] - x, x - bhi[0]), 0.f), dy = fmaxf(fmaxf(blo[1] - y, y - bhi[1]), 0.f),
                            dz = fmaxf(fmaxf(blo[2] - z, z - bhi[2]), 0.f);
                lb[k] += sqrtf(dx * dx + dy * dy + dz * dz);
            }
        }
#pragma unroll
        for (int k = 0; k < kCand; ++k) {
            lf[k] = warp_sum(lf[k]);
            lb[k] = warp_sum(lb[k]);
        }
        if ((tid & 31) == 0) {
#pragma unroll
            for (int k = 0; k < kCand; ++k) {
                s_part[tid >> 5][0][k] = lf[k];
                s_part[tid >> 5][1][k] = lb[k];
            }
        }
        __syncthreads();
        if (tid < kCand) {
            float f = 0.f, bsum_ = 0.f;
            for (int w = 0; w < kWarps; ++w) { f += s_part[w][0][tid]; bsum_ += s_part[w][1][tid]; }
            // generous slack: the bound must stay below the exact fp32 score it is compared with
            s_lb[tid] = fminf(f / (float)n_s, bsum_ / (float)n_d) * 0.9999f - 1e-5f;
            s_score[tid] = INF;
        }
        __syncthreads();
    }
    }   // !ITEM

    // ---- NN grids over both clouds (after the role swap), built once per pair
    GridInfo gS, gD;
    const float4* sortS = nullptr;
    const float4* sortD = nullptr;
    const unsigned short* runS = nullptr;
    const unsigned short* runD = nullptr;
    const bool use_grid = GRIDNN && n_s > 0 && n_d > 0;
    if constexpr (GRIDNN) {
        float4* extra = g_tile;
        float* grid_scratch = reinterpret_cast<float*>(extra + 2 * gridnn_units(a.N));
        GridTiles td{D, extra, reinterpret_cast<uint32_t*>(extra + a.N), grid_scratch};
        GridTiles ts{S, extra + gridnn_units(a.N), reinterpret_cast<uint32_t*>(extra + gridnn_units(a.N) + a.N),
                     grid_scratch};
        if (use_grid) {
            gD = build_grid(td, n_d, a.tau);
            gS = build_grid(ts, n_s, a.tau);
        }
        sortD = td.sorted_p; runD = reinterpret_cast<const unsigned short*>(td.cells_p);
        sortS = ts.sorted_p; runS = reinterpret_cast<const unsigned short*>(ts.cells_p);
    }

    // ---- exact scores
    __shared__ unsigned char s_eval[kCand];      // 1: s_score[k] is the exact score, 0: k is provably not the arg-min
    if (tid < kCand) s_eval[tid] = 0;
    __syncthreads();
    if (GRIDNN && use_grid) {
        // One candidate at a time through the NN grids.  The top peak is evaluated in full; every other candidate that
        // its bounding-box bound cannot exclude is evaluated with EARLY TERMINATION: score_k = min(mean_f, mean_b) and
        // both means are sums of non-negative terms, so as soon as the partial sums prove mean_f > best and
        // mean_b > best the candidate cannot be the arg-min and its exact value is never needed (the zero translation
        // of a cluster that moved by metres is dismissed after its first 128 rows).  A candidate that survives has the
        // exact score of the full evaluation, accumulated in the same order -> the selected translation is the same.
        __shared__ float s_chk[kWarps];
        const int lane = tid & 31, warp = tid >> 5;
        auto block_total = [&](float v) -> float {       // uniform: every thread gets the same total
            __syncthreads();
            v = warp_sum(v);
            if (lane == 0) s_chk[warp] = v;
            __syncthreads();
            float t = s_chk[0];
            for (int w = 1; w < kWarps; ++w) t += s_chk[w];
            return t;
        };
        // mean over the rows of Q of the NN distance in the other cloud; stops (returns true in `exceeded`) once the
        // partial sum proves mean > limit.  BACK: rows of D against (S + t), else rows of (S + t) against D.
        // `budget` > 0: give up (`over`) when the candidate is still open after that many chunks.
        auto mean_nn = [&](bool back, float tx, float ty, float tz, float limit, bool can_stop, int budget, bool& exceeded,
                           bool& over) -> float {
            const int nq = back ? n_d : n_s;
            float acc = 0.f;
            exceeded = false;
            over = false;
            int chunk = 0;
            for (int i0 = 0; i0 < nq; i0 += kThreads, ++chunk) {
                const int i = i0 + tid;
                const bool active = i < nq;
                float m;
                if (back) {
                    const float4 v = active ? D[i] : make_float4(0.f, 0.f, 0.f, 0.f);
                    m = nn_unbounded_grid_warp<true>(gS, sortS, runS, n_s, active, v.x - tx, v.y - ty, v.z - tz, v.x, v.y,
                                                     v.z, tx, ty, tz);
                } else {
                    const float4 v = active ? S[i] : make_float4(0.f, 0.f, 0.f, 0.f);
                    const float qx = __fadd_rn(v.x, tx), qy = __fadd_rn(v.y, ty), qz = __fadd_rn(v.z, tz);
                    m = nn_unbounded_grid_warp<false>(gD, sortD, runD, n_d, active, qx, qy, qz, qx, qy, qz, 0.f, 0.f, 0.f);
                }
                if (active) acc += sqrtf(m);
                // checks after chunks 0, 1, 3, 7, ... (never after the last one)
                if (can_stop && ((chunk + 1) & chunk) == 0 && i0 + kThreads < nq) {
                    if (block_total(acc) * 0.999f > limit * (float)nq) {
                        exceeded = true;
                        break;
                    }
                    if (budget > 0 && chunk + 1 >= budget) {
                        over = true;
                        break;
                    }
                }
            }
            if (exceeded || over) return 0.f;
            // the deterministic final sum of the full evaluation: lanes by butterfly, warps in order
            acc = warp_sum(acc);
            __syncthreads();
            if (lane == 0) s_chk[warp] = acc;
            __syncthreads();
            float t = 0.f;
            for (int w = 0; w < kWarps; ++w) t += s_chk[w];
            return __fdiv_rn(t, (float)nq);
        };
        if constexpr (ITEM) {
            // one queued candidate against the best score of the first kernel
            const int k = only_k;
            const float best = __int_as_float(__ldcg(pairw + 6));
            const float tx = s_t[k][0], ty = s_t[k][1], tz = s_t[k][2];
            bool fx, bx, ov;
            const float ef = mean_nn(false, tx, ty, tz, best, true, 0, fx, ov);
            const float eb = mean_nn(true, tx, ty, tz, fx ? best : fminf(best, ef), true, 0, bx, ov);
            float score = INF;
            bool exact = false;
            if (!fx && !bx) { score = fminf(ef, eb); exact = true; }
            else if (!fx && bx) { if (ef <= best) { score = ef; exact = true; } }
            else if (fx && !bx) { if (eb <= best) { score = eb; exact = true; } }
            if (tid == 0) {
                if (exact) {
                    pairw[k] = __float_as_int(score);
                    atomicOr(pairw + 7, 1 << k);
                }
                __threadfence();
                if (atomicSub(pairw + 8, 1) == 1) {       // the pair's last item: select
                    __threadfence();
                    float sc[kCand];
                    for (int c = 0; c < kCand; ++c) sc[c] = __int_as_float(__ldcg(pairw + c));
                    score_select(a, p, sc, (unsigned int)__ldcg(pairw + 7), s_t);
                }
            }
            __syncthreads();
            return;
        }
        float best = INF;
        const int order[kCand] = {0, kCand - 1, 1, 2, 3, 4};
        const int budget = (pairw != nullptr) ? kDeferBudget : 0;
        int ndefer = 0, deferred[kCand];
        for (int o = 0; o < kCand; ++o) {
            const int k = order[o];
            const bool first = (o == 0);
            if (!first && s_lb[k] > best) continue;                 // (NaN bounds are evaluated, never skipped)
            const float tx = s_t[k][0], ty = s_t[k][1], tz = s_t[k][2];
            bool fx, bx, ov = false;
            const float ef = mean_nn(false, tx, ty, tz, best, !first, first ? 0 : budget, fx, ov);
            // the backward mean matters only below min(best, ef)
            float eb = 0.f;
            bx = false;
            if (!ov) eb = mean_nn(true, tx, ty, tz, fx ? best : fminf(best, ef), !first, first ? 0 : budget, bx, ov);
            if (ov) {                                               // still open at its budget: an item for the second kernel
                deferred[ndefer++] = k;
                continue;
            }
            float score = INF;
            bool exact = false;
            if (!fx && !bx) { score = fminf(ef, eb); exact = true; }          // torch.minimum
            else if (!fx && bx) { if (ef <= best) { score = ef; exact = true; } }   // mean_b > ef: the minimum is ef
            else if (fx && !bx) { if (eb <= best) { score = eb; exact = true; } }   // mean_f > best >= eb
            if (exact) {
                if (tid == 0) { s_score[k] = score; s_eval[k] = 1; }
                best = fminf(best, score);
            }
        }
        __syncthreads();
        if (ndefer > 0) {
            if (tid == 0) {
                unsigned int mask = 0;
                for (int k = 0; k < kCand; ++k) {
                    pairw[k] = __float_as_int(s_eval[k] ? s_score[k] : INF);
                    mask |= s_eval[k] ? (1u << k) : 0u;
                }
                pairw[6] = __float_as_int(best);
                pairw[7] = (int)mask;
                pairw[8] = ndefer;
                const int at = atomicAdd(a.defer, ndefer);
                for (int j = 0; j < ndefer; ++j) a.defer[kDeferHeader + at + j] = p * 8 + deferred[j];
            }
            return;
        }
        __syncthreads();
    } else {
        // full scans, two candidates per sweep: round 1 = the top peak and the zero translation, round 2 = whatever the bound cannot exclude
        int list[kCand];
        int nlist = 2;
        list[0] = 0;
        list[1] = kCand - 1;
        for (int round = 0; round < 2; ++round) {
            if (round == 1) {
                const float best1 = fminf(s_score[0], s_score[kCand - 1]);
                nlist = 0;
                for (int k = 1; k < kCand - 1; ++k) {
                    if (!(s_lb[k] > best1)) list[nlist++] = k;      // NaN bounds are evaluated, never skipped
                }
            }
            // forward: NN of (src_i + t_k) among the dst rows; backward: NN of dst_i among the (src_j + t_k)
            for (int l0 = 0; l0 < nlist; l0 += 2) {
                const int k0 = list[l0], k1 = (l0 + 1 < nlist) ? list[l0 + 1] : list[l0];
                const float t0x = s_t[k0][0], t0y = s_t[k0][1], t0z = s_t[k0][2];
                const float t1x = s_t[k1][0], t1y = s_t[k1][1], t1z = s_t[k1][2];
                float f0 = 0.f, f1 = 0.f, b0 = 0.f, b1 = 0.f;
                for (int i = tid; i < n_s; i += kThreads) {
                    const float4 v = S[i];
                    const float q0x = __fadd_rn(v.x, t0x), q0y = __fadd_rn(v.y, t0y), q0z = __fadd_rn(v.z, t0z);
                    const float q1x = __fadd_rn(v.x, t1x), q1y = __fadd_rn(v.y, t1y), q1z = __fadd_rn(v.z, t1z);
                    float m0 = INF, m1 = INF;
#pragma unroll 4
                    for (int j = 0; j < n_d; ++j) {
                        const float4 c = D[j];
                        m0 = fminf(m0, sqdist(q0x, q0y, q0z, c.x, c.y, c.z));
                        m1 = fminf(m1, sqdist(q1x, q1y, q1z, c.x, c.y, c.z));
                    }
                    f0 += sqrtf(m0);
                    f1 += sqrtf(m1);
                }
                for (int i = tid; i < n_d; i += kThreads) {
                    const float4 v = D[i];
                    float m0 = INF, m1 = INF;
#pragma unroll 4
                    for (int j = 0; j < n_s; ++j) {
                        const float4 c = S[j];
                        m0 = fminf(m0, sqdist(v.x, v.y, v.z, __fadd_rn(c.x, t0x), __fadd_rn(c.y, t0y), __fadd_rn(c.z, t0z)));
                        m1 = fminf(m1, sqdist(v.x, v.y, v.z, __fadd_rn(c.x, t1x), __fadd_rn(c.y, t1y), __fadd_rn(c.z, t1z)));
                    }
                    b0 += sqrtf(m0);
                    b1 += sqrtf(m1);
                }
                f0 = warp_sum(f0); f1 = warp_sum(f1); b0 = warp_sum(b0); b1 = warp_sum(b1);
                __syncthreads();
                if ((tid & 31) == 0) {
                    s_part[tid >> 5][0][0] = f0; s_part[tid >> 5][0][1] = f1;
                    s_part[tid >> 5][1][0] = b0; s_part[tid >> 5][1][1] = b1;
                }
                __syncthreads();
                if (tid < 2) {        // warps added in a fixed order: deterministic
                    float f = 0.f, bb = 0.f;
                    for (int w = 0; w < kWarps; ++w) { f += s_part[w][0][tid]; bb += s_part[w][1][tid]; }
                    const float ef = __fdiv_rn(f, (float)n_s), eb = __fdiv_rn(bb, (float)n_d);
                    s_score[tid == 0 ? k0 : k1] = fminf(ef, eb);        // torch.minimum
                }
                __syncthreads();
            }
        }
        if (tid == 0) {
            const float best1 = fminf(s_score[0], s_score[kCand - 1]);
            for (int k = 0; k < kCand; ++k) s_eval[k] = (k == 0) || (k == kCand - 1) || !(s_lb[k] > best1);
        }
        __syncthreads();
    }
    if (tid == 0) {
        float sc[kCand];
        unsigned int mask = 0;
        for (int k = 0; k < kCand; ++k) {
            sc[k] = s_score[k];
            mask |= s_eval[k] ? (1u << k) : 0u;
        }
        score_select(a, p, sc, mask, s_t);
    }
}

#ifndef ICPF_SCORE_MIN_CTAS
#define ICPF_SCORE_MIN_CTAS 5      // (96 registers instead of 128 for the item kernel: 5 CTAs per SM, as many as the shared memory allows)
#endif
template <bool GRIDNN, bool ITEM>
__global__ void __launch_bounds__(kThreads, ICPF_SCORE_MIN_CTAS) hist_score_kernel(ScoreArgs a) {
    if constexpr (ITEM) {
        __shared__ int s_item;
        for (;;) {
            if (threadIdx.x == 0) s_item = atomicAdd(a.defer + 1, 1);
            __syncthreads();
            const int item = s_item;
            if (item >= __ldcg(a.defer)) return;
            const int code = a.defer[kDeferHeader + item];
            score_pair<GRIDNN, true>(a, code >> 3, code & 7);
            __syncthreads();
        }
    } else {
        score_pair<GRIDNN, false>(a, blockIdx.x, -1);
    }
}

int launch_hist_score(const float* src, const float* dst, int P, int N, const int* cand_idx, const float* bins_x,
                      const float* bins_y, const float* bins_z, int lx, int ly, int lz, float half_bin, float tau,
                      int auto_swap, float* out_pose, float* out_scores, int* out_which, int* defer, cudaStream_t stream) {
    if (P == 0) return ICPF_OK;
    const size_t with_grids = ((size_t)2 * gridnn_units(N) + up16(kRedFloats * 4)) * 16;
    const bool gridnn = with_grids <= (size_t)227 * 1024 && tau > 0.f;
    const size_t smem = gridnn ? with_grids : 0;
    if (!gridnn || kDeferBudget == 0) defer = nullptr;
    auto kernel = gridnn ? hist_score_kernel<true, false> : hist_score_kernel<false, false>;
    cudaError_t err = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess) return (int)err;
    ScoreArgs a{src, dst, N, cand_idx, bins_x, bins_y, bins_z, lx, ly, lz, half_bin, tau, auto_swap, out_pose,
                out_scores, out_which, defer, P};
    if (defer != nullptr) {
        err = cudaMemsetAsync(defer, 0, 2 * sizeof(int), stream);
        if (err != cudaSuccess) return (int)err;
    }
    ICPF_LAUNCH(kernel, P, kThreads, smem, stream)(a);
    err = cudaGetLastError();
    if (err != cudaSuccess || defer == nullptr) return (int)err;
    // the queued (pair, candidate) items: persistent CTAs, as many as the device holds at this shared-memory size
    auto items = hist_score_kernel<true, true>;
    err = cudaFuncSetAttribute(items, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess) return (int)err;
    const int per_sm = (int)std::max<size_t>(1, std::min<size_t>(8, ((size_t)227 * 1024) / (smem + 2048)));
    const long long want = (long long)P * kTopK;
    const int grid = (int)std::min<long long>(want, (long long)kSmCount * per_sm);
    ICPF_LAUNCH(items, grid, kThreads, smem, stream)(a);
    return (int)cudaGetLastError();
}

}  // namespace icpf
