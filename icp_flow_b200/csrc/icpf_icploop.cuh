// icpf_icploop.cuh -- the ICP iteration loop of one pair, fused into ONE pass over the src rows and ONE block
// reduction per iteration.
//
// Reference loop (utils_icp_pytorch3d.py:153-214), per iteration k with the current transform (R_k, T_k):
//     NN of Xt = X0 R_k + T_k in Y  ->  gate  ->  Kabsch on (X0, NN)  ->  (R_{k+1}, T_{k+1})
//     rmse_k = rms | X0 R_{k+1} + T_{k+1} - NN_k |  over the gated rows;  rel_k = (rmse_{k-1} - rmse_k) / rmse_{k-1}
// Restructured, result-identical:
//   * rmse_k needs the NEW transform and the OLD correspondences -- exactly what the correspondence search of
//     iteration k+1 has in hand (it transforms every row with (R_{k+1}, T_{k+1}) and, for the cache test, reloads
//     the old neighbour), so the rmse numerator is accumulated there and only the last iteration needs a pass
//     of its own; the convergence flag of iteration k is therefore recorded one iteration late.
//   * centroids and the centred cross-covariance come from one pass of raw moments about a pivot:
//         H = (S_xy - S_x S_y^T / W) / W,   mu = pivot + S / W,
//     with the pivots set to the previous iteration's centroids, so S_x, S_y ~ 0 and nothing cancels (the
//     reference's two-pass centring, utils_icp_pytorch3d.py:314-336, to fp32 rounding).
//   * the 16 moments are reduced with a transposing butterfly (16 shuffles instead of 80), the partials of the four
//     warps meet in shared memory, and one thread solves the 3x3 problem while the other CTAs of the SM keep the
//     pipes busy.
//   * rows whose cached neighbour cannot be proven (MODE 3) are compacted per warp and searched afterwards with
//     dense lanes.
// All three search modes execute the same accumulation code in the same order, so they agree bit for bit.
#pragma once

#include "icpf_pair.cuh"
#ifdef ICPF_PHASE_CLOCKS
#include <cstdio>
#endif
#if defined(ICPF_ITER_STATS) && defined(ICPF_SIMT_EMU)
// emulator-only instrumentation (tools/iter_stats_simt.py): per iteration {rows re-validated, rows searched, warps that summed again, pairs}
extern "C" { long long icpf_dbg_iter_stats[128][4]; }
#endif

namespace icpf {

// layout of the per-pair loop state a paused run leaves behind (floats; see icp_iterations)
enum : int { S_PIV = 0, S_FRAME = 6, S_FLAGS = 15, S_HPREV = 16, S_WPREV = 25, S_RMSE = 26, S_CONV = 27, S_STATS = 31,
             kIcpStateFloats = 36 };

struct IcpResult {
    float r[9];                   // valid in every thread
    float t[3];                   // valid in every thread
    // the fields below are maintained by thread 0 only
    float rmse;
    int iters;                    // iterations executed by this pair
    unsigned long long conv_lo;   // bit k: relative rmse <= thr at iteration k      (k < 64)
    unsigned long long conv_hi;   //                                                 (64 <= k < 128)
    float searches;               // statistics: rows searched (dense first iteration + deferred rows), dense passes
    int refreshes;
    float prev_rmse;
    bool have_prev;
    float w_prev;                 // clamp(sum of weights) of the last iteration executed
    bool stopped;                 // left the loop at its bitwise fixed point (not at max_it)
    bool tail_ok;                 // ... and every later iteration passes the relative-rmse test (0 <= thr, rmse > 0, finite)
};

__device__ __forceinline__ void set_conv_bit(IcpResult& r, int k) {
    if (k < 64) r.conv_lo |= 1ull << k;
    else if (k < 128) r.conv_hi |= 1ull << (k - 64);
}

// rel_k = (rmse_{k-1} - rmse_k) / rmse_{k-1}  (1 for the first iteration), utils_icp_pytorch3d.py:195-198
__device__ __forceinline__ void record_rmse(IcpResult& r, int k, float rmse, float rel_thr) {
    const float rel = r.have_prev ? __fdiv_rn(r.prev_rmse - rmse, r.prev_rmse) : 1.0f;
    if (rel <= rel_thr) set_conv_bit(r, k);
    r.rmse = rmse;
    r.prev_rmse = rmse;
    r.have_prev = true;
}

// Transposing butterfly: on return lane L holds, in v[0], the warp-wide sum of the input slot
//   idx(L) = 8*bit4(L) + 4*bit3(L) + 2*bit2(L) + bit1(L)     (both lanes of a pair hold the same value)
__device__ __forceinline__ int reduce16_slot(int lane) {
    return ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1);
}

__device__ __forceinline__ float warp_reduce16(float (&v)[16], int lane) {
    bool up = (lane & 16) != 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const float send = up ? v[i] : v[i + 8], keep = up ? v[i + 8] : v[i];
        v[i] = keep + __shfl_xor_sync(FULL_MASK, send, 16);
    }
    up = (lane & 8) != 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float send = up ? v[i] : v[i + 4], keep = up ? v[i + 4] : v[i];
        v[i] = keep + __shfl_xor_sync(FULL_MASK, send, 8);
    }
    up = (lane & 4) != 0;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const float send = up ? v[i] : v[i + 2], keep = up ? v[i + 2] : v[i];
        v[i] = keep + __shfl_xor_sync(FULL_MASK, send, 4);
    }
    up = (lane & 2) != 0;
    {
        const float send = up ? v[0] : v[1], keep = up ? v[1] : v[0];
        v[0] = keep + __shfl_xor_sync(FULL_MASK, send, 2);
    }
    v[0] += __shfl_xor_sync(FULL_MASK, v[0], 1);
    return v[0];
}

// The ICP loop for the pair held in `tl`.
//   MODE 1: candidates = tl.dst(), brute force.  MODE 2: candidates = tl.sorted() + grid `g` (build_grid must have run).
//   MODE 3: grid + correspondence cache.  A row whose cached best candidate is provably still its strict nearest
//           neighbour skips the search: with m the distance the row moved since the cache reference and B the cached
//           lower bound on the distance of every other point, (d_best + m) < B implies every other point is farther
//           than d_best; and B > tau + m alone proves that only the cached candidate can pass the gate.  The gate
//           always uses the exactly recomputed d_best^2, so the results are bit-identical to MODE 1/2.
//   n_s / n_d : valid-row counts (knn `lengths`); tau2 = fp32(thres^2); pivot0 = any point near the clouds
//   init_R / init_T (may be NULL) = init_transform of the reference: used for the first correspondence search only.
//   hist (may be NULL) = this pair's [hist_depth][13] record of (R, T, rmse) after each iteration: when the batch stop
//           falls on an iteration this pair went beyond, its state there is read back instead of re-running the pair.
//   state (may be NULL) = this pair's kIcpStateFloats words of loop state (pivots, warm-start frame, previous H, rmse
//           bookkeeping) as save_icp_state() left them.  RESUME and resume_it >= 2: the loop CONTINUES at iteration
//           resume_it from that state and the record -- the correspondences of iteration resume_it - 1 are searched
//           again under its recorded transform (a search is a function of the transform alone), everything else is
//           restored, so the continued run is the uninterrupted one bit for bit.
template <int MODE, bool RESUME, class Tiles>
__device__ __forceinline__ IcpResult icp_iterations(const Tiles& tl, const GridInfo& g, int n_s, int n_d, float tau2,
                                           int max_it, float rel_thr, bool early_exit, const float* init_R,
                                           const float* init_T, float pivx, float pivy, float pivz,
                                           float* __restrict__ hist = nullptr, int hist_depth = 0,
                                           float* __restrict__ state = nullptr, int resume_it = 0) {
    constexpr bool GRID = MODE >= 2;
    constexpr bool CACHE = MODE == 3;
    using NW = NnWord<Tiles::kPosBits>;
    constexpr unsigned int kNnPosMask = NW::kPosMask, kNnNone = NW::kNone, kNnMasked = NW::kMasked;
    const float INF = __int_as_float(0x7f800000);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float4* __restrict__ cand = GRID ? tl.sorted() : tl.dst();
    const unsigned short* cell_runs = reinterpret_cast<const unsigned short*>(tl.cells());
    unsigned int* __restrict__ nnw = tl.nn();
    float* part = tl.red() + kScrPart;
    float* total = tl.red() + kScrTotal;
    float* bc = tl.bcast();
    const float tau_hi = sqrtf(tau2) * 1.0001f + 1e-6f;
    const int nbatch = (n_s + kThreads - 1) / kThreads;     // row q = b * kThreads + tid

    IcpResult res;
    res.rmse = 0.f;
    res.iters = 0;
    res.conv_lo = res.conv_hi = 0ull;
    res.searches = 0.f;
    res.refreshes = 0;
    res.prev_rmse = 0.f;
    res.have_prev = false;
    res.tail_ok = false;
    if (tid < 9) bc[B_R + tid] = init_R ? init_R[tid] : ((tid % 4 == 0) ? 1.f : 0.f);
    if (tid < 3) bc[B_T + tid] = init_T ? init_T[tid] : 0.f;
    if (tid == 0) {
        bc[B_PX] = pivx; bc[B_PX + 1] = pivy; bc[B_PX + 2] = pivz;
        bc[B_PY] = pivx; bc[B_PY + 1] = pivy; bc[B_PY + 2] = pivz;
        bc[B_EXIT] = 0.f;
        bc[B_ZEROSTEP] = 0.f;
        *reinterpret_cast<int*>(bc + B_SEARCH) = 0;
        bc[B_PIVMOVED] = 0.f;
#pragma unroll
        for (int w = 0; w < kWarps; ++w) reinterpret_cast<int*>(bc + B_DIRTY)[w] = 0;
        reinterpret_cast<KabschState*>(bc + B_KABSCH)->warm = false;     // first solve of this pair starts cold
    }
    __syncthreads();
    float W_prev = 1.f;                                 // thread 0: clamp(sum of weights) of the previous iteration
    bool done = (n_s <= 0 || n_d <= 0 || max_it <= 0);  // engine-defined: nothing to align -> identity
    // (RESUME is a template parameter so that the kernel of the first pass carries none of this in its loop)
    const bool resume = RESUME && (state != nullptr) && (hist != nullptr) && (resume_it >= 2) && (resume_it <= hist_depth) && !done;
    const int first_it = resume ? resume_it - 1 : 0;    // resumed: iteration resume_it - 1 only repeats its search
    if (resume) {
        if (tid < 12) bc[B_R + tid] = hist[(resume_it - 1) * 13 + tid];                                   // (R, T) now
        if (tid < 12) bc[B_RC + tid] = hist[(resume_it - 1) * 13 + tid] - hist[(resume_it - 2) * 13 + tid];  // last step
        if (tid == 0) {
            KabschState* kst = reinterpret_cast<KabschState*>(bc + B_KABSCH);
#pragma unroll
            for (int i = 0; i < 6; ++i) bc[B_PX + i] = state[S_PIV + i];
#pragma unroll
            for (int i = 0; i < 9; ++i) kst->v[i] = state[S_FRAME + i];
            const unsigned int fl = __float_as_uint(state[S_FLAGS]);
            kst->warm = (fl & 1u) != 0u;
            kst->changed = (fl & 2u) != 0u;
            res.have_prev = (fl & 4u) != 0u;
#pragma unroll
            for (int i = 0; i < 9; ++i) bc[B_HPREV + i] = state[S_HPREV + i];
            W_prev = state[S_WPREV];
            res.prev_rmse = res.rmse = state[S_RMSE];
            res.conv_lo = (unsigned long long)__float_as_uint(state[S_CONV]) |
                          ((unsigned long long)__float_as_uint(state[S_CONV + 1]) << 32);
            res.conv_hi = (unsigned long long)__float_as_uint(state[S_CONV + 2]) |
                          ((unsigned long long)__float_as_uint(state[S_CONV + 3]) << 32);
            res.searches = state[S_STATS];
            res.refreshes = (int)state[S_STATS + 1];
            res.iters = resume_it;
        }
        __syncthreads();
    }

#ifdef ICPF_PHASE_CLOCKS
    long long pc[6] = {0, 0, 0, 0, 0, 0}, pt0 = clock64();
#define ICPF_PC(i) { const long long now_ = clock64(); pc[i] += now_ - pt0; pt0 = now_; }
#else
#define ICPF_PC(i)
#endif
    // The moment sums of a warp's rows are a pure function of (correspondences, gate flags, pivots): a warp whose rows
    // kept all three since it last summed them finds its 16 partial sums still in the scratch, bit for bit.
    bool have_partials = false;
    for (int it = first_it; !done && it < max_it; ++it) {
        // the first iteration of a resumed run only rebuilds the correspondences of iteration resume_it - 1, under the
        // transform that iteration searched with (on record); its solve is on record too
        const bool presearch = resume && (it == first_it);
        float R[9], T[3];
        if (presearch) {
#pragma unroll
            for (int i = 0; i < 9; ++i) R[i] = hist[(resume_it - 2) * 13 + i];
#pragma unroll
            for (int i = 0; i < 3; ++i) T[i] = hist[(resume_it - 2) * 13 + 9 + i];
        } else {
#pragma unroll
            for (int i = 0; i < 9; ++i) R[i] = bc[B_R + i];
#pragma unroll
            for (int i = 0; i < 3; ++i) T[i] = bc[B_T + i];
        }
        // The cache is re-anchored every iteration: bounds are kept relative to the row's position in the previous
        // iteration, so the motion that counts is the last step; a row is searched again only when its own bound is used up.
        const bool refresh = !CACHE || (it == first_it);      // dense search of every row (always, without the cache)
        float sq = 0.f;
        int nsearch = 0;
        bool chg = !(CACHE && !refresh) || !have_partials || (bc[B_PIVMOVED] != 0.f);     // this lane saw a reason to sum again
        unsigned short* dlist = GRID ? tl.defer() : nullptr;

        // ---------------- pass A: rmse numerator of the previous iteration + correspondence search of this one
        if (MODE == 1) {
            constexpr int QB = 4;
            for (int b0 = 0; b0 < nbatch; b0 += QB) {
                float qx[QB], qy[QB], qz[QB], best[QB];
                int bidx[QB];
                float4 x0[QB];
#pragma unroll
                for (int k = 0; k < QB; ++k) {
                    const int q = (b0 + k) * kThreads + tid;
                    const bool valid = q < n_s;
                    x0[k] = valid ? tl.src()[q] : make_float4(0.f, 0.f, 0.f, 0.f);
                    apply_rt(R, T, x0[k].x, x0[k].y, x0[k].z, qx[k], qy[k], qz[k]);
                    const unsigned int wold = (it > first_it && valid) ? nnw[q] : (kNnMasked | kNnNone);
                    if (!(wold & kNnMasked)) {
                        const float4 c = cand[wold & kNnPosMask];
                        sq += sqdist(qx[k], qy[k], qz[k], c.x, c.y, c.z);
                    }
                }
                nn_brute<QB>(cand, n_d, qx, qy, qz, best, bidx);
#pragma unroll
                for (int k = 0; k < QB; ++k) {
                    const int q = (b0 + k) * kThreads + tid;
                    if (q < n_s) nnw[q] = NW::pack(bidx[k], 0.f, (best[k] <= tau2) && (x0[k].w > 0.f));
                }
            }
        } else {
            if (CACHE && !refresh && bc[B_ZEROSTEP] != 0.f) {
                // The transform is the previous iteration's bit for bit: every row sits where it sat, so every cached
                // correspondence, gate flag and bound stands as it is; only the rmse numerator is due.
                for (int b = 0; b < nbatch; ++b) {
                    const int q = b * kThreads + tid;
                    if (q >= n_s) continue;
                    const unsigned int wold = nnw[q];
                    if (wold & kNnMasked) continue;
                    const float4 x0 = tl.src()[q];
                    const float4 c = cand[wold & kNnPosMask];
                    float qx, qy, qz;
                    apply_rt(R, T, x0.x, x0.y, x0.z, qx, qy, qz);
                    sq += sqdist(qx, qy, qz, c.x, c.y, c.z);
                }
            } else if (CACHE && !refresh) {
                // the step (R_k - R_{k-1}, T_k - T_{k-1}) thread 0 left in the broadcast block: bounds are kept relative to
                // the row's position in the previous iteration and re-anchored every iteration
                float dr[9], dt[3];
#pragma unroll
                for (int i = 0; i < 9; ++i) dr[i] = bc[B_RC + i];
#pragma unroll
                for (int i = 0; i < 3; ++i) dt[i] = bc[B_TC + i];
                const float anchor_slack = 0.1f * g.pad;      // fp32 rounding of the two positions whose distance is m
                const unsigned int lt = (1u << lane) - 1u;
                int* dcount = reinterpret_cast<int*>(bc + B_SEARCH);
                for (int b = 0; b < nbatch; ++b) {
                    const int q = b * kThreads + tid;
                    const bool valid = q < n_s;
                    const float4 x0 = valid ? tl.src()[q] : make_float4(0.f, 0.f, 0.f, 0.f);
                    const unsigned int wold = valid ? nnw[q] : (kNnMasked | kNnNone);
                    float qx, qy, qz;
                    apply_rt(R, T, x0.x, x0.y, x0.z, qx, qy, qz);
                    const unsigned int pos = wold & kNnPosMask;
                    const bool has = pos != kNnNone;
                    const float4 c = cand[has ? pos : 0u];
                    const float d2 = has ? sqdist(qx, qy, qz, c.x, c.y, c.z) : INF;
                    if (!(wold & kNnMasked)) sq += d2;          // rmse numerator: new transform, old correspondence
                    // distance the row moved in this step (inflated), and what is left of its bound
                    const float mx = fmaf(x0.z, dr[6], fmaf(x0.y, dr[3], x0.x * dr[0])) + dt[0];
                    const float my = fmaf(x0.z, dr[7], fmaf(x0.y, dr[4], x0.x * dr[1])) + dt[1];
                    const float mz = fmaf(x0.z, dr[8], fmaf(x0.y, dr[5], x0.x * dr[2])) + dt[2];
                    const float rem = NW::bound(wold) - (sqrtf(fmaf(mz, mz, fmaf(my, my, mx * mx))) * 1.0002f + anchor_slack);
                    // (a) every other point is provably beyond the gate: only the cached candidate can pass it;
                    // (b) the cached candidate is provably still the strict nearest neighbour: d_best < rem, tested on
                    //     the squares with 1.5e-4 of head-room for the rounding of either side
                    const bool hit = (rem > tau_hi) || (rem > 0.f && d2 * 1.0003f < rem * rem);
                    const bool masked = !((d2 <= tau2) && (x0.w > 0.f));           // (no candidate: d2 = inf)
                    if (hit) {
                        nnw[q] = NW::pack(has ? (int)pos : -1, rem, !masked);       // re-anchored at the new position
                        chg = chg || (masked != ((wold & kNnMasked) != 0u));      // the gate flag changed
                    }
                    const bool need = valid && !hit;
                    const unsigned int vote = __ballot_sync(FULL_MASK, need);
                    if (vote != 0u) {
                        // rows whose cached neighbour failed go to ONE list per CTA (warp-aggregated atomic), so that
                        // the searches afterwards run with dense lanes
                        int base = 0;
                        if (lane == 0) base = atomicAdd(dcount, __popc(vote));
                        base = __shfl_sync(FULL_MASK, base, 0);
                        if (need) dlist[base + __popc(vote & lt)] = (unsigned short)q;
                    }
                }
                __syncthreads();
                ICPF_PC(0)
                nsearch = *dcount;
                if (nsearch > 0) {
                    // ---------------- searches of the rows whose cached neighbour could not be proven
                    for (int i = tid; i < nsearch; i += kThreads) {
                        const int q = dlist[i];
                        const float4 x0 = tl.src()[q];
                        float qx, qy, qz;
                        apply_rt(R, T, x0.x, x0.y, x0.z, qx, qy, qz);
                        const NnTop2 nn = grid_search(g, cand, cell_runs, qx, qy, qz);
                        const unsigned int wnew = NW::pack(nn.pos1, fminf(sqrtf(nn.d2), nn.box) * 0.9999f - g.pad, (nn.d1 <= tau2) && (x0.w > 0.f));
                        if (((wnew ^ nnw[q]) & (kNnPosMask | kNnMasked)) != 0u) reinterpret_cast<int*>(bc + B_DIRTY)[(q >> 5) & (kWarps - 1)] = 1;
                        nnw[q] = wnew;
                    }
                    __syncthreads();
                    if (reinterpret_cast<int*>(bc + B_DIRTY)[warp] != 0) chg = true;
                    __syncwarp();
                    if (lane == 0) reinterpret_cast<int*>(bc + B_DIRTY)[warp] = 0;
                }
                ICPF_PC(1)
            } else {
                for (int b = 0; b < nbatch; ++b) {
                    const int q = b * kThreads + tid;
                    if (q >= n_s) continue;
                    const float4 x0 = tl.src()[q];
                    float qx, qy, qz;
                    apply_rt(R, T, x0.x, x0.y, x0.z, qx, qy, qz);
                    if (it > first_it) {
                        const unsigned int wold = nnw[q];
                        if (!(wold & kNnMasked)) {
                            const float4 c = cand[wold & kNnPosMask];
                            sq += sqdist(qx, qy, qz, c.x, c.y, c.z);
                        }
                    }
                    // g.pad (>= 1e-4 m, >= 16 ulp of the largest coordinate) covers the fp32 rounding of the
                    // transformed positions whose separation m is bounded analytically
                    const NnTop2 nn = grid_search(g, cand, cell_runs, qx, qy, qz);
                    nnw[q] = NW::pack(nn.pos1, CACHE ? fminf(sqrtf(nn.d2), nn.box) * 0.9999f - g.pad : 0.f, (nn.d1 <= tau2) && (x0.w > 0.f));
                }
            }
        }

        if (presearch) {
            __syncthreads();          // the correspondence words are complete before the next iteration reads them
            continue;
        }

        // ---------------- pass B: raw moments about the pivots (same code and order in every mode)
        const bool sum_again = __any_sync(FULL_MASK, chg) != 0;
#if defined(ICPF_ITER_STATS) && defined(ICPF_SIMT_EMU)
        if (lane == 0 && it < 128) {
            icpf_dbg_iter_stats[it][0] += (bc[B_ZEROSTEP] != 0.f && warp == 0) ? 1 : 0;
            icpf_dbg_iter_stats[it][1] += (warp == 0) ? nsearch : 0;
            icpf_dbg_iter_stats[it][2] += sum_again ? 1 : 0;
            icpf_dbg_iter_stats[it][3] += (warp == 0) ? 1 : 0;
        }
#endif
        if (sum_again) {
            float mom[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) mom[i] = 0.f;
            const float px = bc[B_PX], py = bc[B_PX + 1], pz = bc[B_PX + 2];
            const float ux = bc[B_PY], uy = bc[B_PY + 1], uz = bc[B_PY + 2];
            for (int b = 0; b < nbatch; ++b) {
                const int q = b * kThreads + tid;
                const unsigned int w = (q < n_s) ? nnw[q] : kNnMasked;
                if (w & kNnMasked) continue;
                const float4 x = tl.src()[q];
                const float4 y = cand[w & kNnPosMask];
                const float ax = x.x - px, ay = x.y - py, az = x.z - pz;
                const float bx = y.x - ux, by = y.y - uy, bz = y.z - uz;
                mom[0] += 1.f;
                mom[1] += ax; mom[2] += ay; mom[3] += az;
                mom[4] += bx; mom[5] += by; mom[6] += bz;
                mom[7] = fmaf(ax, bx, mom[7]);   mom[8] = fmaf(ax, by, mom[8]);   mom[9] = fmaf(ax, bz, mom[9]);
                mom[10] = fmaf(ay, bx, mom[10]); mom[11] = fmaf(ay, by, mom[11]); mom[12] = fmaf(ay, bz, mom[12]);
                mom[13] = fmaf(az, bx, mom[13]); mom[14] = fmaf(az, by, mom[14]); mom[15] = fmaf(az, bz, mom[15]);
            }
            const float msum = warp_reduce16(mom, lane);
            if ((lane & 1) == 0) part[warp * kSums + reduce16_slot(lane)] = msum;
            have_partials = true;
        }
        sq = warp_sum(sq);
        if (lane == 0) {
            part[warp * kSums + 16] = sq;
            part[warp * kSums + 17] = (warp == 0) ? (float)nsearch : 0.f;      // (the CTA's count)
        }
        __syncthreads();
        ICPF_PC(2)

        // ---------------- one warp folds the partials, one thread solves the 3x3 problem
        if (warp == 0) {
            if (lane < kSums) {
                float s = part[lane];
#pragma unroll
                for (int w = 1; w < kWarps; ++w) s += part[w * kSums + lane];
                total[lane] = s;
            }
            __syncwarp();
            if (lane == 0) {
                if (it > (resume ? resume_it : 0)) {        // (a resumed run recorded iteration resume_it - 1 before it paused)
                    record_rmse(res, it - 1, sqrtf(__fdiv_rn(total[16], W_prev)), rel_thr);
                    if (hist != nullptr && it - 1 < hist_depth) hist[(it - 1) * 13 + 12] = res.rmse;
                }
                const float W = fmaxf(total[0], 1e-9f);
                const float sx0 = total[1], sx1 = total[2], sx2 = total[3];
                const float sy0 = total[4], sy1 = total[5], sy2 = total[6];
                const float iW = __frcp_rn(W);
                const float mx0 = sx0 * iW, mx1 = sx1 * iW, mx2 = sx2 * iW;
                const float my0 = sy0 * iW, my1 = sy1 * iW, my2 = sy2 * iW;
                float h[9];
                h[0] = fmaf(-sx0, my0, total[7]) * iW;  h[1] = fmaf(-sx0, my1, total[8]) * iW;
                h[2] = fmaf(-sx0, my2, total[9]) * iW;  h[3] = fmaf(-sx1, my0, total[10]) * iW;
                h[4] = fmaf(-sx1, my1, total[11]) * iW; h[5] = fmaf(-sx1, my2, total[12]) * iW;
                h[6] = fmaf(-sx2, my0, total[13]) * iW; h[7] = fmaf(-sx2, my1, total[14]) * iW;
                h[8] = fmaf(-sx2, my2, total[15]) * iW;
                // The reference's rotation is a pure function of H (torch.svd, utils_icp_pytorch3d.py:339).  The
                // warm-started solve below is a function of (H, stored frame), and two frames that are both converged
                // to working precision can hand over to each other for ever (R flips by an ulp, the rmse by 1e-6 of its
                // value -- the size of the reference's own stopping threshold).  So when H repeats bit for bit the
                // previous R stands: same H, same R, as in the reference, and the pair is at its fixed point.
                KabschState* kst = reinterpret_cast<KabschState*>(bc + B_KABSCH);
#ifdef ICPF_NO_SAME_H       // A/B switch (tools/ab_kernel.py): always solve
                bool same_h = false;
#else
                bool same_h = (it > 0);
#endif
#pragma unroll
                for (int i = 0; i < 9; ++i) same_h = same_h && (__float_as_uint(h[i]) == __float_as_uint(bc[B_HPREV + i]));
                Rot3 rot;
                if (same_h) {
#pragma unroll
                    for (int i = 0; i < 9; ++i) rot.r[i] = bc[B_R + i];
                    kst->changed = false;
                } else {
                    rot = kabsch_rotation(h, kst);
#pragma unroll
                    for (int i = 0; i < 9; ++i) bc[B_HPREV + i] = h[i];
                }
                // centroids: mu = pivot + S / W   (zero-weight iteration: both are the pivots, T = 0 like the reference)
                const float cx0 = bc[B_PX] + mx0, cx1 = bc[B_PX + 1] + mx1, cx2 = bc[B_PX + 2] + mx2;
                const float cy0 = bc[B_PY] + my0, cy1 = bc[B_PY + 1] + my1, cy2 = bc[B_PY + 2] + my2;
                float t[3];
                const bool empty = !(total[0] > 0.f);
                t[0] = empty ? 0.f : cy0 - fmaf(cx2, rot.r[6], fmaf(cx1, rot.r[3], cx0 * rot.r[0]));
                t[1] = empty ? 0.f : cy1 - fmaf(cx2, rot.r[7], fmaf(cx1, rot.r[4], cx0 * rot.r[1]));
                t[2] = empty ? 0.f : cy2 - fmaf(cx2, rot.r[8], fmaf(cx1, rot.r[5], cx0 * rot.r[2]));
                // The pair has reached its fixed point only when the WHOLE state repeats: the transform, the pivots of
                // the moment sums and the warm-start frame of the 3x3 solve -- then every later iteration is this one.
                // (with no gated correspondence the pivots and the frame are left untouched, so only (R,T) = (I,0) counts)
                // The pivots only have to be NEAR the centroids (so that S_x, S_y stay small); they follow them while
                // they are more than a millimetre off and then stay put, which makes H a function of the
                // correspondences alone once the alignment has settled.
                const float off = fmaxf(fmaxf(fmaxf(fabsf(mx0), fabsf(mx1)), fabsf(mx2)),
                                        fmaxf(fmaxf(fabsf(my0), fabsf(my1)), fabsf(my2)));
                const bool move_pivot = !empty && (off > 1e-3f);
                bool same = (it > 0) && !reinterpret_cast<KabschState*>(bc + B_KABSCH)->changed && !move_pivot;
#pragma unroll
                for (int i = 0; i < 9; ++i) same = same && (__float_as_uint(rot.r[i]) == __float_as_uint(bc[B_R + i]));
#pragma unroll
                for (int i = 0; i < 3; ++i) same = same && (__float_as_uint(t[i]) == __float_as_uint(bc[B_T + i]));
                bc[B_PIVMOVED] = move_pivot ? 1.f : 0.f;
                if (move_pivot) {
                    bc[B_PX] = cx0; bc[B_PX + 1] = cx1; bc[B_PX + 2] = cx2;
                    bc[B_PY] = cy0; bc[B_PY + 1] = cy1; bc[B_PY + 2] = cy2;
                }
                W_prev = W;
                res.iters = it + 1;
                if (CACHE) {
                    res.searches += refresh ? (float)n_s : total[17];
                    res.refreshes += refresh ? 1 : 0;
                    // the step the rows take between this search and the next one
                    bool zero = true;
#pragma unroll
                    for (int i = 0; i < 9; ++i) {
                        const float d = rot.r[i] - bc[B_R + i];
                        bc[B_RC + i] = d;
                        zero = zero && (d == 0.f);
                    }
#pragma unroll
                    for (int i = 0; i < 3; ++i) {
                        const float d = t[i] - bc[B_T + i];
                        bc[B_TC + i] = d;
                        zero = zero && (d == 0.f);
                    }
                    bc[B_ZEROSTEP] = zero ? 1.f : 0.f;
                }
#pragma unroll
                for (int i = 0; i < 9; ++i) bc[B_R + i] = rot.r[i];
#pragma unroll
                for (int i = 0; i < 3; ++i) bc[B_T + i] = t[i];
                *reinterpret_cast<int*>(bc + B_SEARCH) = 0;
                bc[B_EXIT] = (early_exit && same) ? 1.f : 0.f;
            }
        }
        __syncthreads();
        ICPF_PC(3)
        done = bc[B_EXIT] != 0.f;
        // (R, T) after this iteration; thread 0 rewrites the broadcast block only behind the next iteration's barriers
        if (hist != nullptr && it < hist_depth && tid < 12) hist[it * 13 + tid] = bc[B_R + tid];
    }

#ifdef ICPF_PHASE_CLOCKS
    if (tid == 0 && (blockIdx.x % 251) == 0)
        printf("PC block %d: A %lld D %lld B %lld S %lld (cycles over all iterations)\n", (int)blockIdx.x, pc[0], pc[1], pc[2], pc[3]);
#endif
    // ---------------- rmse of the last iteration executed (final transform against its own correspondences)
#pragma unroll
    for (int i = 0; i < 9; ++i) res.r[i] = bc[B_R + i];
#pragma unroll
    for (int i = 0; i < 3; ++i) res.t[i] = bc[B_T + i];
    const int iters = __shfl_sync(FULL_MASK, res.iters, 0);   // thread 0's count, needed below by warp 0 only
    if (n_s > 0 && n_d > 0 && max_it > 0) {
        float sq = 0.f;
        for (int b = 0; b < nbatch; ++b) {
            const int q = b * kThreads + tid;
            const unsigned int w = (q < n_s) ? nnw[q] : kNnMasked;
            if (w & kNnMasked) continue;
            const float4 x = tl.src()[q];
            const float4 c = cand[w & kNnPosMask];
            float qx, qy, qz;
            apply_rt(res.r, res.t, x.x, x.y, x.z, qx, qy, qz);
            sq += sqdist(qx, qy, qz, c.x, c.y, c.z);
        }
        sq = warp_sum(sq);
        if (lane == 0) part[warp * kSums + 16] = sq;
        __syncthreads();
        if (tid == 0) {
            float s = part[16];
#pragma unroll
            for (int w = 1; w < kWarps; ++w) s += part[w * kSums + 16];
            const float rmse = sqrtf(__fdiv_rn(s, W_prev));
            record_rmse(res, iters - 1, rmse, rel_thr);
            if (hist != nullptr && iters - 1 < hist_depth) hist[(iters - 1) * 13 + 12] = rmse;
            // stopped at a bitwise fixed point: every later iteration repeats this state, so its relative rmse is
            // (rmse - rmse) / rmse = 0 (NaN when rmse == 0)
            res.tail_ok = (rmse > 0.f) && (0.0f <= rel_thr) && (rmse < INF);
            if (iters < max_it && res.tail_ok) {
                for (int k = iters; k < max_it && k < 128; ++k) set_conv_bit(res, k);
            }
        }
    }
    if (tid == 0) {
        res.w_prev = W_prev;
        res.stopped = done;
        res.tail_ok = res.tail_ok && done;
    }
    return res;
}

// Thread 0: leave the loop state of a finished (or paused) run in `state` for a later continuation (icp_iterations,
// `state` / `resume_it`); S_FLAGS bit 3 = stopped at its fixed point, bit 4 = its later iterations pass the stop test.  `bc` is the pair's broadcast block.
__device__ __forceinline__ void save_icp_state(const IcpResult& res, const float* bc, float* __restrict__ state) {
    const KabschState* kst = reinterpret_cast<const KabschState*>(bc + B_KABSCH);
#pragma unroll
    for (int i = 0; i < 6; ++i) state[S_PIV + i] = bc[B_PX + i];
#pragma unroll
    for (int i = 0; i < 9; ++i) state[S_FRAME + i] = kst->v[i];
    state[S_FLAGS] = __uint_as_float((kst->warm ? 1u : 0u) | (kst->changed ? 2u : 0u) | (res.have_prev ? 4u : 0u) |
                                     (res.stopped ? 8u : 0u) | (res.tail_ok ? 16u : 0u));
#pragma unroll
    for (int i = 0; i < 9; ++i) state[S_HPREV + i] = bc[B_HPREV + i];
    state[S_WPREV] = res.w_prev;
    state[S_RMSE] = res.prev_rmse;
    state[S_CONV] = __uint_as_float((unsigned int)res.conv_lo);
    state[S_CONV + 1] = __uint_as_float((unsigned int)(res.conv_lo >> 32));
    state[S_CONV + 2] = __uint_as_float((unsigned int)res.conv_hi);
    state[S_CONV + 3] = __uint_as_float((unsigned int)(res.conv_hi >> 32));
    state[S_STATS] = res.searches;
    state[S_STATS + 1] = (float)res.refreshes;
}

}  // namespace icpf
