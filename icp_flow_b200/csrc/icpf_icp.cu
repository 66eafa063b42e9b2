// icpf_icp.cu -- batched ICP loop: one CTA per (src, dst) cluster pair, tiles TMA-staged into shared memory once,
// all iterations (NN search, gating, Kabsch sums, closed-form rotation, re-transform, rmse) run on-chip.
//
// Replaces utils_icp_pytorch3d.iterative_closest_point (/root/reference/utils_icp_pytorch3d.py:37-225) and its
// batch-coupled stopping rule (:209) without any host synchronisation:
//   kernel 1  every pair iterates to its bitwise fixed point (or max_iterations) and records, per iteration, whether
//             its relative-rmse test passed (128-bit mask);
//   kernel 2  AND-reduces the masks over the batch -> k* = first iteration at which the reference's `.all()` fires;
//   kernel 3  gives the pairs that were still moving at k* their state at k*, read back from the per-iteration record
//             (R, T, rmse; 52 B per pair and iteration) kernel 1 keeps -- no pair is run twice for it.
#include "icpf_internal.h"
#include <type_traits>

#include "icpf_icploop.cuh"

namespace icpf {

static_assert(kIcpStateWords == kIcpStateFloats, "state record");

struct IcpArgs {
    const float* src;
    const float* dst;
    const float* init_R;   // [P,9] or NULL
    const float* init_T;   // [P,3] or NULL
    const float* init_pose;  // [P,16] or NULL: 4x4 applied to the moved cloud before ICP (utils_icp.py:21)
    int auto_swap;           // 1: register the cloud with fewer valid rows onto the other (utils_match.py:139-146)
    int P, N;
    float tau;             // fp32(thres)
    float tau2;            // fp32(thres^2)
    float cell_factor;     // grid cell size in units of the padded gate radius
    int max_it;
    float rel_thr;
    int early_exit;
    float* out_R;
    float* out_T;
    float* out_rmse;
    float* out_pose;       // [P,16] or NULL
    int* iters;            // [P] (never NULL inside the library)
    uint32_t* conv;        // [P,4]
    int* stats;            // [P,2] {full searches, cache refreshes} of the first pass
    const int* decided;    // second full pass: runs only while *decided == 0; NULL otherwise
    float* hist;           // [P, kIcpHistDepth, 13] (R, T, rmse) after each iteration, or NULL
    float* state;          // [P, kIcpStateWords] loop state of the pairs the capped first pass paused, or NULL
    int cap;               // first pass: iteration cap (<= max_it)
    unsigned char* big_ws; // global-memory variant: per-pair workspace (pair_global_ws_bytes(N) each)
    float* const* peer_pose;   // fused all-gather: DEVICE array of `peer_world` base pointers (peer-mapped [*,16]) or NULL
    int peer_world;
    int peer_row0;             // first row of this rank's block in the gathered buffer
};

// FULLPASS = the pass after a capped first pass: only the pairs still moving at the cap run, continuing from where the
// first pass paused them (a separate instantiation, so that the kernel of the first pass -- the one every call pays for --
// carries no continuation logic in its loop).
template <int MODE, bool BIG, bool FULLPASS>
__global__ void __launch_bounds__(kThreads, BIG ? 4 : 7) icp_pairs_kernel(IcpArgs a) {
    const int p = blockIdx.x;
    int max_it = a.max_it;
    bool early_exit = a.early_exit != 0;
    int resume_it = 0;
    if (FULLPASS) {
        if (*a.decided != 0) return;      // the capped first pass already found the batch stop
        // the capped pass left the loop state of every pair behind: bit 3 of its flags = stopped at its fixed point
        const unsigned int flags = a.state != nullptr ? __float_as_uint(a.state[(size_t)p * kIcpStateWords + S_FLAGS]) : 0u;
        const bool stopped = (flags & 8u) != 0u;
        if (a.iters[p] < a.cap || stopped) {
            // This pair stopped at its bitwise fixed point within the cap: every later iteration repeats that state, so
            // its transform stands and its convergence history continues by the tail rule (icpf_icploop.cuh; the capped
            // pass recorded whether the rule holds -- bit cap-1 is NOT a stand-in for it when the pair stopped exactly at
            // iteration cap-1: that bit is then the real test of that iteration) -- only the pairs still moving at the cap
            // go on, from where they paused.
            if (threadIdx.x == 0 && a.cap >= 1) {
                uint32_t* c = a.conv + (size_t)p * 4;
                const int last = a.cap - 1;
                const bool tail_ok = a.state != nullptr ? (flags & 16u) != 0u : (((c[last >> 5] >> (last & 31)) & 1u) != 0u);
                if (tail_ok) {
                    for (int k = a.cap; k < max_it && k < 128; ++k) c[k >> 5] |= 1u << (k & 31);
                }
            }
            return;
        }
#ifndef ICPF_NO_RESUME         // A/B switch (tools/ab_kernel.py): the full pass starts the pairs over
        if (a.state != nullptr && a.hist != nullptr) resume_it = a.cap;
#endif
    } else {
        max_it = min(max_it, a.cap);
    }
    constexpr bool GRID = MODE >= 2;
    typename std::conditional<BIG, PairTilesG, PairTiles>::type tl;
    if constexpr (BIG) {
        // rows stay in global memory; the workspace holds the sorted copy, the transformed src rows, nn words, list
        const size_t n = (size_t)(a.N + kThreads - 1) / kThreads * kThreads;
        unsigned char* w = a.big_ws + (size_t)p * pair_global_ws_bytes(a.N);
        tl.src_p = const_cast<float4*>(reinterpret_cast<const float4*>(a.src) + (size_t)p * a.N);
        tl.dst_p = const_cast<float4*>(reinterpret_cast<const float4*>(a.dst) + (size_t)p * a.N);
        tl.sorted_p = reinterpret_cast<float4*>(w);
        tl.nn_p = reinterpret_cast<unsigned int*>(w + n * 32);
        tl.defer_p = reinterpret_cast<unsigned short*>(w + n * 36);
        tl.defer_cap = (int)(n / kWarps);
        if (!GRID) tl.sorted_p = tl.dst_p;
    } else {
        tl = carve_pair_tiles<GRID>(a.N);
        if (threadIdx.x == 0) {
            mbar_init(tl.bar(), 1);
            fence_barrier_init();
        }
        __syncthreads();
        load_pair_tiles(tl, a.src + (size_t)p * a.N * 4, a.dst + (size_t)p * a.N * 4, a.N, 0);
    }

    // valid-row counts (knn `lengths`): number of rows with flag > 0  (utils_icp_pytorch3d.py:109-112)
    float cnt[2] = {0.f, 0.f};
    for (int q = threadIdx.x; q < a.N; q += kThreads) {
        cnt[0] += (tl.src()[q].w > 0.f) ? 1.f : 0.f;
        cnt[1] += (tl.dst()[q].w > 0.f) ? 1.f : 0.f;
    }
    block_allreduce_sum<2, kWarps>(cnt, tl.red() + kScrPart);
    int n_s = (int)cnt[0], n_d = (int)cnt[1];
    if (a.auto_swap && n_s > n_d) {
        tl.template swap_clouds<GRID>();
        const int n = n_s; n_s = n_d; n_d = n;
    }
    const float4 piv = tl.dst()[0];      // first pivot of the moment sums: any point of the fixed cloud
    __syncthreads();                   // scratch and the raw dst rows are reused below
    if (a.init_pose != nullptr) {
        // src' = [x y z 1] pose^T with the flag carried through (utils_helper.py:76-87)
        if (threadIdx.x < 12) tl.bcast()[threadIdx.x] = a.init_pose[(size_t)p * 16 + threadIdx.x];
        __syncthreads();
        float m[12];
#pragma unroll
        for (int i = 0; i < 12; ++i) m[i] = tl.bcast()[i];
        if constexpr (BIG) {
            // the inputs are read-only: the moved cloud goes to the workspace
            const size_t n = (size_t)(a.N + kThreads - 1) / kThreads * kThreads;
            float4* moved = reinterpret_cast<float4*>(a.big_ws + (size_t)p * pair_global_ws_bytes(a.N) + n * 16);
            for (int q = threadIdx.x; q < n_s; q += kThreads) moved[q] = transform_row(m, tl.src()[q]);
            tl.src_p = moved;
        } else {
            for (int q = threadIdx.x; q < n_s; q += kThreads) tl.src()[q] = transform_row(m, tl.src()[q]);
        }
        __syncthreads();
    }

    GridInfo g;
    if (GRID && n_s > 0 && n_d > 0) g = build_grid(tl, n_d, a.tau, a.cell_factor);
    const IcpResult r = icp_iterations<MODE, FULLPASS>(tl, g, n_s, n_d, a.tau2, max_it, a.rel_thr, early_exit,
                                       a.init_R ? a.init_R + (size_t)p * 9 : nullptr,
                                       a.init_T ? a.init_T + (size_t)p * 3 : nullptr, piv.x, piv.y, piv.z,
                                       a.hist != nullptr ? a.hist + (size_t)p * kIcpHistDepth * kIcpHistFloats : nullptr,
                                       kIcpHistDepth,
                                       (FULLPASS && a.state != nullptr) ? a.state + (size_t)p * kIcpStateWords : nullptr,
                                       resume_it);

    // the final (R, T) also sit in the broadcast block (no dynamic register indexing)
    if (threadIdx.x < 9) a.out_R[(size_t)p * 9 + threadIdx.x] = tl.bcast()[B_R + threadIdx.x];
    if (threadIdx.x < 3) a.out_T[(size_t)p * 3 + threadIdx.x] = tl.bcast()[B_T + threadIdx.x];
    if (a.out_pose && threadIdx.x < 16) {
        // column-convention 4x4 [[R^T, T],[0,1]]  (utils_icp.py:60-65)
        const int row = threadIdx.x >> 2, col = threadIdx.x & 3;
        float v;
        if (row == 3) v = (col == 3) ? 1.f : 0.f;
        else if (col == 3) v = tl.bcast()[B_T + row];
        else v = tl.bcast()[B_R + col * 3 + row];
        a.out_pose[(size_t)p * 16 + threadIdx.x] = v;
    }
    if (a.peer_pose != nullptr && threadIdx.x < 16) {
        // Fused all-gather: the 64 B transform goes straight into every rank's gathered buffer through the
        // NVLink-mapped peer pointers (plain st.global on peer addresses); a cross-rank barrier after the kernel makes
        // the rows visible to their consumers.  No NCCL collective on the data path.
        const int row = threadIdx.x >> 2, col = threadIdx.x & 3;
        float v;
        if (row == 3) v = (col == 3) ? 1.f : 0.f;
        else if (col == 3) v = tl.bcast()[B_T + row];
        else v = tl.bcast()[B_R + col * 3 + row];
        for (int w = 0; w < a.peer_world; ++w) a.peer_pose[w][(size_t)(a.peer_row0 + p) * 16 + threadIdx.x] = v;
    }
    if (!FULLPASS && a.state != nullptr && threadIdx.x == 0) {
        save_icp_state(r, tl.bcast(), a.state + (size_t)p * kIcpStateWords);       // a capped pass: the full pass may go on
    }
    if (threadIdx.x == 0) {
        if (a.out_rmse) a.out_rmse[p] = r.rmse;
        a.iters[p] = r.iters;
        a.stats[(size_t)p * 2 + 0] = (int)r.searches;
        a.stats[(size_t)p * 2 + 1] = r.refreshes;
        a.conv[(size_t)p * 4 + 0] = (uint32_t)r.conv_lo;
        a.conv[(size_t)p * 4 + 1] = (uint32_t)(r.conv_lo >> 32);
        a.conv[(size_t)p * 4 + 2] = (uint32_t)r.conv_hi;
        a.conv[(size_t)p * 4 + 3] = (uint32_t)(r.conv_hi >> 32);
    }
}

// State at the batch stop from the per-iteration record: a pair that executed more iterations than the batch did
// (batch[0] = k* + 1 <= kIcpHistDepth) takes (R, T, rmse) of iteration k* and reports k* + 1 iterations -- exactly what
// running it for k* + 1 iterations gives (the record holds the very values that run produced).  One thread per pair.
struct IcpSelectArgs {
    const float* hist;
    const int* batch;
    int P;
    int* iters;
    float* out_R;
    float* out_T;
    float* out_rmse;
    float* out_pose;
    float* const* peer_pose;
    int peer_world;
    int peer_row0;
};

__device__ __forceinline__ void select_pair(const IcpSelectArgs& a, int p, int b) {
    if (a.iters[p] <= b || b < 1 || b > kIcpHistDepth) return;
    const float* h = a.hist + ((size_t)p * kIcpHistDepth + (b - 1)) * kIcpHistFloats;
    float r[9], t[3];
#pragma unroll
    for (int i = 0; i < 9; ++i) r[i] = h[i];
#pragma unroll
    for (int i = 0; i < 3; ++i) t[i] = h[9 + i];
#pragma unroll
    for (int i = 0; i < 9; ++i) a.out_R[(size_t)p * 9 + i] = r[i];
#pragma unroll
    for (int i = 0; i < 3; ++i) a.out_T[(size_t)p * 3 + i] = t[i];
    if (a.out_rmse) a.out_rmse[p] = h[12];
    a.iters[p] = b;
    // column-convention 4x4 [[R^T, T],[0,1]]  (utils_icp.py:60-65)
    float m[16];
#pragma unroll
    for (int row = 0; row < 3; ++row) {
#pragma unroll
        for (int col = 0; col < 3; ++col) m[row * 4 + col] = r[col * 3 + row];
        m[row * 4 + 3] = t[row];
    }
    m[12] = m[13] = m[14] = 0.f;
    m[15] = 1.f;
    if (a.out_pose) {
#pragma unroll
        for (int i = 0; i < 16; ++i) a.out_pose[(size_t)p * 16 + i] = m[i];
    }
    if (a.peer_pose != nullptr) {
        for (int w = 0; w < a.peer_world; ++w) {
#pragma unroll
            for (int i = 0; i < 16; ++i) a.peer_pose[w][(size_t)(a.peer_row0 + p) * 16 + i] = m[i];
        }
    }
}

__global__ void __launch_bounds__(128) icp_select_batch_kernel(IcpSelectArgs a) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p < a.P) select_pair(a, p, a.batch[0]);
}

#ifndef ICPF_RESOLVE_THREADS
#define ICPF_RESOLVE_THREADS 1024
#endif
constexpr int kResolveThreads = ICPF_RESOLVE_THREADS;     // one CTA: the launch is a latency chain, more threads = fewer trips
// AND of the per-pair convergence masks -> first iteration k* where every pair passes (utils_icp_pytorch3d.py:209).
// batch[0] = iterations the reference loop would have executed, batch[1] = converged flag.
// `limit` = iterations the masks cover (the cap of the first pass, or max_it); `decided` (may be NULL) is set to 1 when
// the answer is final and left 0 when the capped pass could not tell (a later full pass decides); `and_out` (may be NULL)
// receives the AND of the masks -- what a caller with several shards exchanges (IcpPhase, icpf_internal.h).
// `sel.hist != NULL`: once the stop is final, the same launch hands the pairs that went beyond it their state at the stop
// (select_pair) -- the call ends with one launch instead of two.
__global__ void __launch_bounds__(kResolveThreads) icp_resolve_batch_kernel(const uint32_t* conv, int P, int max_it, int limit,
                                                                int batch_stop, int* batch, int* decided,
                                                                uint32_t* and_out, IcpSelectArgs sel) {
    __shared__ uint32_t s_and[4];
    const bool settled = decided != nullptr && limit == max_it && *decided != 0;     // second resolve, nothing left to do
    if (settled) {
        if (sel.hist != nullptr) {
            const int b = batch[0];
            for (int p = threadIdx.x; p < P; p += blockDim.x) select_pair(sel, p, b);
        }
        return;
    }
    if (threadIdx.x < 4) s_and[threadIdx.x] = 0xffffffffu;
    __syncthreads();
    uint32_t m[4] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu};
    for (int p = threadIdx.x; p < P; p += blockDim.x) {
#pragma unroll
        for (int i = 0; i < 4; ++i) m[i] &= conv[(size_t)p * 4 + i];
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        for (int o = 16; o > 0; o >>= 1) m[i] &= __shfl_xor_sync(FULL_MASK, m[i], o);
        if ((threadIdx.x & 31) == 0) atomicAnd(&s_and[i], m[i]);
    }
    __syncthreads();
    if (and_out != nullptr && threadIdx.x < 4) and_out[threadIdx.x] = s_and[threadIdx.x];
    if (threadIdx.x == 0) {
        int kstar = -1;
        if (batch_stop) {
            for (int i = 0; i < 4 && kstar < 0; ++i) {
                if (s_and[i]) kstar = i * 32 + (__ffs(s_and[i]) - 1);
            }
        }
        int verdict = 1;
        if (kstar >= 0 && kstar < limit) {
            batch[0] = kstar + 1;
            batch[1] = 1;
        } else if (limit >= max_it) {
            batch[0] = max_it;
            batch[1] = 0;
        } else {
            batch[0] = limit;          // provisional: no pair is re-run, the full pass follows
            batch[1] = 0;
            verdict = 0;
        }
        if (decided) *decided = verdict;
    }
    if (sel.hist != nullptr) {
        __syncthreads();
        const int b = batch[0];
        for (int p = threadIdx.x; p < P; p += blockDim.x) select_pair(sel, p, b);
    }
}

// phase 2 of a phased call: the batch stop found over ALL shards, as given by the caller
__global__ void icp_set_batch_kernel(int* batch, int iters, int converged) {
    batch[0] = iters;
    batch[1] = converged;
}

size_t icp_big_workspace_bytes(int P, int N) {
    return pair_needs_global(N) ? (size_t)P * pair_global_ws_bytes(N) : 0;
}

// One CTA per (destination rank, 256-row slice): 16-byte stores of this rank's contiguous block into the peer buffer.
__global__ void __launch_bounds__(256) peer_push_kernel(const float4* __restrict__ local, float* const* peer_pose, int row0,
                                                        int n_vec) {
    float4* dst = reinterpret_cast<float4*>(peer_pose[blockIdx.y]) + (size_t)row0 * 4;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_vec; i += gridDim.x * blockDim.x) dst[i] = local[i];
}

int launch_peer_push(const float* local_pose, float* const* peer_pose_dev, int world, int row0, int P, cudaStream_t stream) {
    if (P == 0 || world == 0) return ICPF_OK;
    const int n_vec = P * 4;
    const int gx = max(1, min(8, (n_vec + 255) / 256));
    ICPF_LAUNCH(peer_push_kernel, dim3(gx, world), 256, 0, stream)(reinterpret_cast<const float4*>(local_pose), peer_pose_dev, row0, n_vec);
    return (int)cudaGetLastError();
}

// Compact rows (xyz of the valid rows, CSR offsets) -> the reference's padded [B,N,4] layout (pad_segment).
__global__ void __launch_bounds__(256) expand_rows_kernel(const float* __restrict__ rows, const int* __restrict__ offsets, int N,
                                                          float4* __restrict__ out) {
    const int b = blockIdx.x;
    const int lo = offsets[b], n = min(max(offsets[b + 1] - lo, 0), N);
    const float* r = rows + (size_t)lo * 3;
    float4* o = out + (size_t)b * N;
    for (int i = threadIdx.x; i < N; i += blockDim.x) {
        o[i] = (i < n) ? make_float4(r[3 * i], r[3 * i + 1], r[3 * i + 2], 1.0f) : make_float4(1e8f, 1e8f, 1e8f, 0.0f);
    }
}

int launch_expand_rows(const float* rows, const int32_t* offsets, int B, int N, float* out, cudaStream_t stream) {
    if (B == 0) return ICPF_OK;
    ICPF_LAUNCH(expand_rows_kernel, B, 256, 0, stream)(rows, offsets, N, reinterpret_cast<float4*>(out));
    return (int)cudaGetLastError();
}

int launch_icp(const float* src, const float* dst, const float* init_R, const float* init_T, const float* init_pose,
               int auto_swap, int P, int N, const icpf_params& prm, float* out_R, float* out_T,
               float* out_rmse, float* out_pose, int* out_iters, uint32_t* out_conv, int* out_batch, void* workspace,
               size_t workspace_bytes, cudaStream_t stream, const IcpPhase* phase, const icpf_icp_ext* ext) {
    if (P == 0) return ICPF_OK;
    // nn_mode: 0 auto (grid + cache whenever the tiles fit in shared memory), 1 brute force, 2 grid, 3 grid + cache
    if (N > kMaxRows) return ICPF_E_UNSUPPORTED;
    const bool grid = prm.nn_mode != 1;
    // clusters whose tiles do not fit shared memory run the global-memory variant (same code, rows in L2)
    const bool big = pair_needs_global(N);
    const size_t smem = big ? pair_global_smem_bytes() : pair_smem_bytes(N, grid);
    // workspace: iters [P] | conv [P,4] | batch [2] + flags | stats [P,2] | history  (icpf_internal.h)
    const size_t need = icp_workspace_bytes(P) + icp_big_workspace_bytes(P, N);
    if (workspace == nullptr || workspace_bytes < need) return ICPF_E_WORKSPACE;
    unsigned char* ws = static_cast<unsigned char*>(workspace);
    int* iters = out_iters ? out_iters : reinterpret_cast<int*>(ws);
    uint32_t* conv = out_conv ? out_conv : reinterpret_cast<uint32_t*>(ws + icp_ws_off_conv(P));
    int* batch = out_batch ? out_batch : reinterpret_cast<int*>(ws + icp_ws_off_batch(P));

    auto kernel = big ? (!grid ? icp_pairs_kernel<1, true, false>
                               : (prm.nn_mode == 2 ? icp_pairs_kernel<2, true, false> : icp_pairs_kernel<3, true, false>))
                      : (!grid ? icp_pairs_kernel<1, false, false>
                               : (prm.nn_mode == 2 ? icp_pairs_kernel<2, false, false> : icp_pairs_kernel<3, false, false>));
    auto kernel_full = big ? (!grid ? icp_pairs_kernel<1, true, true>
                                    : (prm.nn_mode == 2 ? icp_pairs_kernel<2, true, true> : icp_pairs_kernel<3, true, true>))
                           : (!grid ? icp_pairs_kernel<1, false, true>
                                    : (prm.nn_mode == 2 ? icp_pairs_kernel<2, false, true> : icp_pairs_kernel<3, false, true>));
    cudaError_t err = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess) return (int)err;
    err = cudaFuncSetAttribute(kernel_full, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess) return (int)err;

    IcpArgs a;
    a.src = src; a.dst = dst; a.init_R = init_R; a.init_T = init_T; a.P = P; a.N = N;
    a.init_pose = init_pose; a.auto_swap = auto_swap;
    a.tau = (float)prm.thres_dist;
    a.tau2 = (float)(prm.thres_dist * prm.thres_dist);   // python: thres**2 in double, compared in fp32
    a.cell_factor = prm.reserved[0] > 0 ? (float)prm.reserved[0] * 1e-3f : kCellFactor;
    a.max_it = prm.max_iterations;
    a.rel_thr = prm.relative_rmse_thr;
    a.early_exit = prm.early_exit;
    a.out_R = out_R; a.out_T = out_T; a.out_rmse = out_rmse; a.out_pose = out_pose;
    a.iters = iters; a.conv = conv; a.decided = nullptr;
    // Pairs that never reach a bitwise fixed point (limit cycles of a few ulp, non-convergent clusters) would run all
    // max_iterations in the first pass although the reference's batch stop fires after 10-25: cap the first pass and
    // fall back to a full pass only when the batch stop was not found below the cap.
    const int kFirstPassCap = 32;
    const bool capped = prm.batch_stop && prm.early_exit && prm.max_iterations > kFirstPassCap;
    a.cap = capped ? kFirstPassCap : prm.max_iterations;
    a.big_ws = big ? ws + icp_workspace_bytes(P) : nullptr;
    const bool peer = ext != nullptr && ext->peer_pose_dev != nullptr && ext->peer_world > 0;
    a.peer_pose = peer ? reinterpret_cast<float* const*>(ext->peer_pose_dev) : nullptr;
    a.peer_world = peer ? ext->peer_world : 0;
    a.peer_row0 = peer ? ext->peer_row0 : 0;
    cudaEvent_t prof_start = ext ? static_cast<cudaEvent_t>(ext->start_event) : nullptr;
    cudaEvent_t prof_stop = ext ? static_cast<cudaEvent_t>(ext->stop_event) : nullptr;
    int* decided = reinterpret_cast<int*>(ws + icp_ws_off_batch(P)) + 8;
    a.stats = reinterpret_cast<int*>(ws + icp_ws_off_stats(P));
    // With a batch stop every pair records (R, T, rmse) after each iteration; the pairs still moving at the stop read
    // their state back from that record (icp_select_batch_kernel) instead of being run again.
    a.hist = prm.batch_stop ? reinterpret_cast<float*>(ws + icp_ws_off_hist(P)) : nullptr;
    // ... and a capped first pass leaves its loop state, so that the full pass continues the pairs still moving at the
    // cap instead of starting them over
    a.state = capped ? reinterpret_cast<float*>(ws + icp_ws_off_state(P)) : nullptr;
    const int ph = phase ? phase->phase : -1;          // -1: the whole call at once (one device holds the batch)
    uint32_t* and_out = phase ? phase->and_out : nullptr;
    // state at the batch stop for the pairs that went beyond it (read back from the record): the last stage of the call
    IcpSelectArgs sa;
    sa.hist = a.hist; sa.batch = batch; sa.P = P; sa.iters = iters;
    sa.out_R = out_R; sa.out_T = out_T; sa.out_rmse = out_rmse; sa.out_pose = out_pose;
    sa.peer_pose = a.peer_pose; sa.peer_world = a.peer_world; sa.peer_row0 = a.peer_row0;
    IcpSelectArgs no_sel = sa;
    no_sel.hist = nullptr;
    // a whole call resolves the stop and reads the state back in ONE launch after its last pass
    const bool fuse_sel = (ph < 0) && prm.batch_stop;
    if (ph <= 0) {
        if (prof_start && prof_stop) cudaEventRecord(prof_start, stream);
        ICPF_LAUNCH(kernel, P, kThreads, smem, stream)(a);
        err = cudaGetLastError();
        if (prof_start && prof_stop) cudaEventRecord(prof_stop, stream);
        if (err != cudaSuccess) return (int)err;
        ICPF_LAUNCH(icp_resolve_batch_kernel, 1, kResolveThreads, 0, stream)(conv, P, prm.max_iterations, a.cap, prm.batch_stop, batch,
                                                        capped ? decided : nullptr, and_out,
                                                        (fuse_sel && !capped) ? sa : no_sel);
        err = cudaGetLastError();
        if (err != cudaSuccess) return (int)err;
        if (ph == 0) return ICPF_OK;
    }
    if (ph == 1 || (ph < 0 && capped)) {
        if (ph == 1) {
            // the shards decided together that the stop lies beyond the capped pass
            err = cudaMemsetAsync(decided, 0, sizeof(int), stream);
            if (err != cudaSuccess) return (int)err;
        }
        a.decided = decided;
        ICPF_LAUNCH(kernel_full, P, kThreads, smem, stream)(a);
        a.decided = nullptr;
        ICPF_LAUNCH(icp_resolve_batch_kernel, 1, kResolveThreads, 0, stream)(conv, P, prm.max_iterations, prm.max_iterations, prm.batch_stop,
                                                        batch, decided, and_out, fuse_sel ? sa : no_sel);
        err = cudaGetLastError();
        if (err != cudaSuccess) return (int)err;
        if (ph == 1) return ICPF_OK;
    }
    if (ph == 2) {
        ICPF_LAUNCH(icp_set_batch_kernel, 1, 1, 0, stream)(batch, phase->batch_iters, phase->converged);
        err = cudaGetLastError();
        if (err != cudaSuccess) return (int)err;
    }
    if (prm.batch_stop && !fuse_sel) {
        ICPF_LAUNCH(icp_select_batch_kernel, (P + 127) / 128, 128, 0, stream)(sa);
        err = cudaGetLastError();
        if (err != cudaSuccess) return (int)err;
    }
    return ICPF_OK;
}

}  // namespace icpf
