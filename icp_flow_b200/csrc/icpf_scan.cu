// icpf_scan.cu -- the callers either side of the per-pair path, scan-level:
//
//   cluster index    the reference addresses a cluster as `points[labels == l]` -- one boolean mask over the whole scan
//                    per use (utils_check.py:24-25, utils_match.py:85-86).  Here one stable counting sort per scan
//                    builds a CSR index (rows of a cluster in scan order) plus per-cluster statistics.
//   sanity_check     utils_check.py:21-49 on the statistics: one thread per candidate pair, ordered compaction.
//   gather / pad     utils_match.py:84-91 + pad_segment (utils_helper.py:185-196): one CTA per (pair, side) writes the
//                    padded [P, max_points, 4] batch straight from the index.
//   flow recovery    utils_flow.py:57-69: per point  T_cluster(label) * pose * p - p.
//
// All of it is HBM-/latency-bound integer and gather work on O(10^5) points per scan; nothing here is a contraction.
#include "icpf_internal.h"

namespace icpf {

namespace {

constexpr int kScanThreads = 256;
constexpr int kScanWarps = kScanThreads / 32;

// a label addresses a cluster slot when it is a non-negative integer below n_labels (ground / unclustered points
// carry -1e8 / -1, utils_flow.py:28-31)
__device__ __forceinline__ int label_slot(float l, int n_labels) {
    if (!(l >= 0.f) || !(l < (float)n_labels)) return -1;
    const int li = (int)l;
    return ((float)li == l) ? li : -1;
}

__global__ void __launch_bounds__(kScanThreads)
cluster_count_kernel(const float* __restrict__ labels, int n, int n_labels, int chunk, int* __restrict__ blockhist) {
    const int b = blockIdx.x;
    const int lo = b * chunk, hi = min(n, lo + chunk);
    int* hist = blockhist + (size_t)b * n_labels;
    for (int i = lo + threadIdx.x; i < hi; i += kScanThreads) {
        const int s = label_slot(labels[i], n_labels);
        if (s >= 0) atomicAdd(&hist[s], 1);
    }
}

// totals[l] = points of label l over all blocks
__global__ void __launch_bounds__(kScanThreads)
cluster_totals_kernel(const int* __restrict__ blockhist, int n_blocks, int n_labels, int* __restrict__ offsets) {
    const int l = blockIdx.x * kScanThreads + threadIdx.x;
    if (l >= n_labels) return;
    int t = 0;
#pragma unroll 8
    for (int b = 0; b < n_blocks; ++b) t += blockhist[(size_t)b * n_labels + l];       // (independent loads, batched by the unroll)
    offsets[l] = t;
}

// in-place exclusive scan of offsets[0, n_labels) (one block); offsets[n_labels] = number of indexed points
__global__ void __launch_bounds__(1024) cluster_offsets_kernel(int* __restrict__ offsets, int n_labels) {
    __shared__ int wsum[32];
    __shared__ int running;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) running = 0;
    __syncthreads();
    for (int base = 0; base < n_labels; base += 1024) {
        const int l = base + tid;
        const int v = (l < n_labels) ? offsets[l] : 0;
        int incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int u = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += u;
        }
        if (lane == 31) wsum[warp] = incl;
        __syncthreads();
        int before = running;
        for (int w = 0; w < warp; ++w) before += wsum[w];
        if (l < n_labels) offsets[l] = before + incl - v;
        __syncthreads();
        if (tid == 1023) running = before + incl;
        __syncthreads();
    }
    if (tid == 0) offsets[n_labels] = running;
}

// blockhist[b][l] <- first output position of block b's points of label l
__global__ void __launch_bounds__(kScanThreads)
cluster_bases_kernel(int* __restrict__ blockhist, int n_blocks, int n_labels, const int* __restrict__ offsets) {
    const int l = blockIdx.x * kScanThreads + threadIdx.x;
    if (l >= n_labels) return;
    int off = offsets[l];
    // eight loads in flight per step: one thread walks all blocks of its label, and a load-add-store chain per block was
    // 293 dependent L2 round trips (73 us per scan on a 150 k-point frame)
    constexpr int kBatch = 8;
    for (int b0 = 0; b0 < n_blocks; b0 += kBatch) {
        int v[kBatch];
#pragma unroll
        for (int k = 0; k < kBatch; ++k) v[k] = (b0 + k < n_blocks) ? blockhist[(size_t)(b0 + k) * n_labels + l] : 0;
#pragma unroll
        for (int k = 0; k < kBatch; ++k) {
            if (b0 + k < n_blocks) blockhist[(size_t)(b0 + k) * n_labels + l] = off;
            off += v[k];
        }
    }
}

// stable scatter: points of one label keep their scan order (points[labels == l] of the reference)
__global__ void __launch_bounds__(kScanThreads)
cluster_scatter_kernel(const float* __restrict__ labels, int n, int n_labels, int chunk, int* __restrict__ blockhist,
                       int* __restrict__ order) {
    const int b = blockIdx.x;
    const int lo = b * chunk, hi = min(n, lo + chunk);
    int* cursor = blockhist + (size_t)b * n_labels;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int t0 = lo; t0 < hi; t0 += kScanThreads) {
        const int i = t0 + threadIdx.x;
        const int s = (i < hi) ? label_slot(labels[i], n_labels) : -1;
        // lanes of one warp with the same label form a group; lanes without a slot get a private key
        const unsigned int grp = __match_any_sync(0xffffffffu, s >= 0 ? s : -1 - lane);
        const int rank = __popc(grp & ((1u << lane) - 1u));
        const int leader = __ffs(grp) - 1;
        int base = 0;
        for (int w = 0; w < kScanWarps; ++w) {         // warps in scan order: earlier rows first
            if (w == warp && s >= 0 && rank == 0) {
                base = atomicAdd(&cursor[s], __popc(grp));
            }
            __syncthreads();
        }
        base = __shfl_sync(0xffffffffu, base, leader);
        if (s >= 0) order[base + rank] = i;
    }
}

// per-cluster statistics in a fixed order (deterministic): stats[l] = {mean x, y, z, sorted |max - min| extents (3), 0, 0}
// Means are accumulated in fp64 and rounded once; min / max are exact, the extent is the reference's fp32 subtraction
// (get_bbox_tensor, utils_helper.py:166-170).
__global__ void __launch_bounds__(128)
cluster_stats_kernel(const float* __restrict__ points, int stride, const int* __restrict__ order,
                     const int* __restrict__ offsets, int n_labels, float* __restrict__ stats) {
    const int l = blockIdx.x;
    const int s = offsets[l], e = offsets[l + 1];
    const float INF = __int_as_float(0x7f800000);
    double sx = 0.0, sy = 0.0, sz = 0.0;
    float lo[3] = {INF, INF, INF}, hi[3] = {-INF, -INF, -INF};
    for (int j0 = s + threadIdx.x; j0 < e; j0 += 4 * 128) {
        int idx[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) idx[u] = (j0 + u * 128 < e) ? order[j0 + u * 128] : -1;
        float px[4], py[4], pz[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (idx[u] >= 0) {
                const float* p = points + (size_t)idx[u] * stride;
                px[u] = p[0]; py[u] = p[1]; pz[u] = p[2];
            }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (idx[u] >= 0) {
                sx += (double)px[u]; sy += (double)py[u]; sz += (double)pz[u];
                lo[0] = fminf(lo[0], px[u]); lo[1] = fminf(lo[1], py[u]); lo[2] = fminf(lo[2], pz[u]);
                hi[0] = fmaxf(hi[0], px[u]); hi[1] = fmaxf(hi[1], py[u]); hi[2] = fmaxf(hi[2], pz[u]);
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        sx += __shfl_xor_sync(0xffffffffu, sx, o);
        sy += __shfl_xor_sync(0xffffffffu, sy, o);
        sz += __shfl_xor_sync(0xffffffffu, sz, o);
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            lo[k] = fminf(lo[k], __shfl_xor_sync(0xffffffffu, lo[k], o));
            hi[k] = fmaxf(hi[k], __shfl_xor_sync(0xffffffffu, hi[k], o));
        }
    }
    __shared__ double ssum[4][3];
    __shared__ float sbox[4][6];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) {
        ssum[warp][0] = sx; ssum[warp][1] = sy; ssum[warp][2] = sz;
#pragma unroll
        for (int k = 0; k < 3; ++k) { sbox[warp][k] = lo[k]; sbox[warp][3 + k] = hi[k]; }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        float* out = stats + (size_t)l * 8;
        const int n = e - s;
        if (n <= 0) {
#pragma unroll
            for (int k = 0; k < 8; ++k) out[k] = 0.f;
            return;
        }
        double t[3];
        float ext[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            t[k] = ssum[0][k] + ssum[1][k] + ssum[2][k] + ssum[3][k];
            const float a = fminf(fminf(sbox[0][k], sbox[1][k]), fminf(sbox[2][k], sbox[3][k]));
            const float c = fmaxf(fmaxf(sbox[0][3 + k], sbox[1][3 + k]), fmaxf(sbox[2][3 + k], sbox[3][3 + k]));
            ext[k] = fabsf(__fsub_rn(c, a));
            out[k] = (float)(t[k] / (double)n);
        }
        // sorted([x, y, z]) ascending
        float a = ext[0], c = ext[1], d = ext[2], tmp;
        if (a > c) { tmp = a; a = c; c = tmp; }
        if (c > d) { tmp = c; c = d; d = tmp; }
        if (a > c) { tmp = a; a = c; c = tmp; }
        out[3] = a; out[4] = c; out[5] = d;
        out[6] = 0.f; out[7] = 0.f;
    }
}

// ---------------------------------------------------------------------------------------------- sanity_check
struct SanityGates {
    int min_cluster_size;
    float translation_frame;
    float thres_box;
};

__device__ __forceinline__ int pair_slot(long long l, int n_labels) {
    return (l >= 0 && l < (long long)n_labels) ? (int)l : -1;
}

// utils_check.py:21-49, one thread per candidate pair; kept pairs are written in input order
__global__ void __launch_bounds__(1024)
sanity_check_kernel(const int* __restrict__ src_offsets, const float* __restrict__ src_stats, int n_src,
                    const int* __restrict__ dst_offsets, const float* __restrict__ dst_stats, int n_dst,
                    const long long* __restrict__ pairs, int P, SanityGates g, int* __restrict__ out_keep,
                    long long* __restrict__ out_pairs, int* __restrict__ out_count, int cross_nd) {
    __shared__ int wsum[32];
    __shared__ int running;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) running = 0;
    __syncthreads();
    for (int base = 0; base < P; base += 1024) {
        const int p = base + tid;
        bool keep = false;
        long long ls = 0, ld = 0;
        if (p < P) {
            if (cross_nd > 0) {
                // all-against-all of two label lists (utils_match.py:48-50): `pairs` = [src list | dst list], src-major
                ls = pairs[p / cross_nd];
                ld = pairs[P / cross_nd + p % cross_nd];
            } else {
                ls = pairs[2 * p];
                ld = pairs[2 * p + 1];
            }
            const int a = pair_slot(ls, n_src), b = pair_slot(ld, n_dst);
            if (a >= 0 && b >= 0) {          // `min(pair) < 0: continue`; a label without points has length 0
                const int na = src_offsets[a + 1] - src_offsets[a], nb = dst_offsets[b + 1] - dst_offsets[b];
                const float* sa = src_stats + (size_t)a * 8;
                const float* sb = dst_stats + (size_t)b * 8;
                keep = min(na, nb) >= g.min_cluster_size && min(na, nb) > 0;
                // torch.linalg.norm((mean_dst - mean_src)[0:2]) > translation_frame  (fp32)
                const float dx = __fsub_rn(sb[0], sa[0]), dy = __fsub_rn(sb[1], sa[1]);
                const float nrm = sqrtf(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)));
                keep = keep && !(nrm > g.translation_frame);
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const float lo = fminf(sa[3 + k], sb[3 + k]), hi = fmaxf(sa[3 + k], sb[3 + k]);
                    keep = keep && !(lo < __fmul_rn(g.thres_box, hi));
                }
            }
            if (out_keep != nullptr) out_keep[p] = keep ? 1 : 0;
        }
        const unsigned int vote = __ballot_sync(0xffffffffu, keep);
        const int rank = __popc(vote & ((1u << lane) - 1u));
        if (lane == 0) wsum[warp] = __popc(vote);
        __syncthreads();
        int before = running;
        for (int w = 0; w < warp; ++w) before += wsum[w];
        if (keep) {
            out_pairs[2 * (before + rank)] = ls;
            out_pairs[2 * (before + rank) + 1] = ld;
        }
        __syncthreads();
        if (tid == 1023) running = before + __popc(vote);
        __syncthreads();
    }
    if (tid == 0) *out_count = running;
}

// ---------------------------------------------------------------------------------------------- gather + pad
struct ScanView {
    const float* points;
    const int* order;
    const int* offsets;
    int stride;
    int n_labels;
};

// One CTA per (pair, side): rows (x, y, z, 1) of the cluster in scan order, then (1e8, 1e8, 1e8, 0) -- pad_segment.
// A cluster with more than max_points rows takes the rows the caller drew (sample_rows: positions inside the cluster,
// the reference's torch.randperm(len)[:max_points]); without a sample it keeps its first max_points rows.
__global__ void __launch_bounds__(kScanThreads)
gather_pairs_kernel(ScanView src, ScanView dst, const long long* __restrict__ pairs, int max_points,
                    const int* __restrict__ sample_rows, const long long* __restrict__ sample_offsets,
                    float4* __restrict__ out_src, float4* __restrict__ out_dst) {
    const int p = blockIdx.x, side = blockIdx.y;
    const ScanView v = side ? dst : src;
    float4* out = (side ? out_dst : out_src) + (size_t)p * max_points;
    const int slot = pair_slot(pairs[2 * p + side], v.n_labels);
    int s = 0, cnt = 0;
    if (slot >= 0) {
        s = v.offsets[slot];
        cnt = v.offsets[slot + 1] - s;
    }
    const long long so = sample_offsets ? sample_offsets[2 * p + side] : -1;
    const int rows = min(cnt, max_points);
    for (int r = threadIdx.x; r < max_points; r += kScanThreads) {
        float4 o = make_float4(1e8f, 1e8f, 1e8f, 0.f);
        if (r < rows) {
            const int j = (so >= 0) ? sample_rows[so + r] : r;
            const float* q = v.points + (size_t)v.order[s + j] * v.stride;
            o = make_float4(q[0], q[1], q[2], 1.f);
        }
        out[r] = o;
    }
}

// ---------------------------------------------------------------------------------------------- flow recovery
constexpr int kFlowTable = 16384;      // label -> pair look-up table in shared memory (u16), labels beyond it are scanned

// utils_flow.py:57-69:  T_per_point = (pairs[:,0] == label ? transformations[k] : I) @ pose;  flow = T_per_point p - p
__global__ void __launch_bounds__(kScanThreads)
flow_kernel(const float* __restrict__ points, int stride, const float* __restrict__ labels, int n,
            const float* __restrict__ pair_labels, int pair_stride, const float* __restrict__ transforms, int K,
            const float* __restrict__ pose, float* __restrict__ flow) {
    __shared__ unsigned int table[kFlowTable / 2];      // packed u16 halves, 0 = no pair, k + 1 otherwise
    __shared__ float spose[16];
    __shared__ int overflow;                             // some pair label does not fit the table
    for (int i = threadIdx.x; i < kFlowTable / 2; i += kScanThreads) table[i] = 0u;
    if (threadIdx.x < 16) spose[threadIdx.x] = pose ? pose[threadIdx.x] : ((threadIdx.x % 5 == 0) ? 1.f : 0.f);
    if (threadIdx.x == 0) overflow = 0;
    __syncthreads();
    for (int k = threadIdx.x; k < K; k += kScanThreads) {
        const float l = pair_labels[(size_t)k * pair_stride];
        const int li = (int)l;
        if (l >= 0.f && l < (float)kFlowTable && (float)li == l) {
            // the highest row wins, like the sequential index_put of the reference on duplicated labels
            const int sh = (li & 1) * 16;
            unsigned int* w = &table[li >> 1];
            unsigned int old = *w;
            while (true) {
                const unsigned int cur = (old >> sh) & 0xffffu;
                if (cur >= (unsigned int)(k + 1)) break;
                const unsigned int upd = (old & ~(0xffffu << sh)) | ((unsigned int)(k + 1) << sh);
                const unsigned int seen = atomicCAS(w, old, upd);
                if (seen == old) break;
                old = seen;
            }
        } else {
            overflow = 1;
        }
    }
    __syncthreads();
    const bool scan_all = overflow != 0;
    for (int i = blockIdx.x * kScanThreads + threadIdx.x; i < n; i += gridDim.x * kScanThreads) {
        const float l = labels[i];
        int k = -1;
        const int li = (int)l;
        if (l >= 0.f && l < (float)kFlowTable && (float)li == l) {
            k = (int)((table[li >> 1] >> ((li & 1) * 16)) & 0xffffu) - 1;
        } else if (scan_all) {
            for (int j = 0; j < K; ++j)
                if (pair_labels[(size_t)j * pair_stride] == l) k = j;
        }
        float m[12];
        if (k >= 0) {
            const float* T = transforms + (size_t)k * 16;
#pragma unroll
            for (int r = 0; r < 3; ++r)
#pragma unroll
                for (int c = 0; c < 4; ++c)
                    m[r * 4 + c] = fmaf(T[r * 4 + 3], spose[12 + c],
                                        fmaf(T[r * 4 + 2], spose[8 + c],
                                             fmaf(T[r * 4 + 1], spose[4 + c], T[r * 4] * spose[c])));
        } else {
#pragma unroll
            for (int j = 0; j < 12; ++j) m[j] = spose[j];      // eye(4) @ pose
        }
        const float* q = points + (size_t)i * stride;
        const float x = q[0], y = q[1], z = q[2];
        float* o = flow + (size_t)i * 3;
        o[0] = __fsub_rn(fmaf(1.0f, m[3], fmaf(z, m[2], fmaf(y, m[1], x * m[0]))), x);
        o[1] = __fsub_rn(fmaf(1.0f, m[7], fmaf(z, m[6], fmaf(y, m[5], x * m[4]))), y);
        o[2] = __fsub_rn(fmaf(1.0f, m[11], fmaf(z, m[10], fmaf(y, m[9], x * m[8]))), z);
    }
}

inline int scan_blocks(int n, int n_labels) {
    int b = (n + 511) / 512;
    b = b < 1 ? 1 : b;
    b = b > 296 ? 296 : b;
    const long long cap = (1ll << 22) / (n_labels > 0 ? n_labels : 1);     // <= 16 MB of per-block counters
    if (b > cap) b = (int)(cap < 1 ? 1 : cap);
    return b;
}

}  // namespace

size_t cluster_index_workspace_bytes(int n, int n_labels) {
    return align_up((size_t)scan_blocks(n, n_labels) * (size_t)(n_labels > 0 ? n_labels : 1) * 4, 256);
}

int launch_cluster_index(const float* points, int stride, const float* labels, int n, int n_labels, int* order,
                         int* offsets, float* stats, void* ws, size_t ws_bytes, cudaStream_t stream) {
    const int B = scan_blocks(n, n_labels);
    const size_t need = cluster_index_workspace_bytes(n, n_labels);
    if (!ws || ws_bytes < need) return ICPF_E_WORKSPACE;
    int* blockhist = static_cast<int*>(ws);
    cudaError_t e = cudaMemsetAsync(blockhist, 0, (size_t)B * n_labels * 4, stream);
    if (e != cudaSuccess) return (int)e;
    const int chunk = ((n + B - 1) / B + kScanThreads - 1) / kScanThreads * kScanThreads;
    const int lb = (n_labels + kScanThreads - 1) / kScanThreads;
    if (n > 0) ICPF_LAUNCH(cluster_count_kernel, B, kScanThreads, 0, stream)(labels, n, n_labels, chunk, blockhist);
    ICPF_LAUNCH(cluster_totals_kernel, lb, kScanThreads, 0, stream)(blockhist, B, n_labels, offsets);
    ICPF_LAUNCH(cluster_offsets_kernel, 1, 1024, 0, stream)(offsets, n_labels);
    ICPF_LAUNCH(cluster_bases_kernel, lb, kScanThreads, 0, stream)(blockhist, B, n_labels, offsets);
    if (n > 0) ICPF_LAUNCH(cluster_scatter_kernel, B, kScanThreads, 0, stream)(labels, n, n_labels, chunk, blockhist, order);
    ICPF_LAUNCH(cluster_stats_kernel, n_labels, 128, 0, stream)(points, stride, order, offsets, n_labels, stats);
    return (int)cudaGetLastError();
}

int launch_sanity_check(const int* src_offsets, const float* src_stats, int n_src, const int* dst_offsets,
                        const float* dst_stats, int n_dst, const int64_t* pairs, int P, int min_cluster_size,
                        float translation_frame, float thres_box, int* out_keep, int64_t* out_pairs, int* out_count,
                        cudaStream_t stream, int cross_nd) {
    SanityGates g{min_cluster_size, translation_frame, thres_box};
    ICPF_LAUNCH(sanity_check_kernel, 1, 1024, 0, stream)(src_offsets, src_stats, n_src, dst_offsets, dst_stats, n_dst,
                                                reinterpret_cast<const long long*>(pairs), P, g, out_keep,
                                                reinterpret_cast<long long*>(out_pairs), out_count, cross_nd);
    return (int)cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------- selection (row f1)
// The rejection loop and selection of match_pairs (utils_match.py:70-75, 94-135) in one launch: the reference scatters
// the accepted registrations into [n_src, n_dst] matrices pair by pair, takes per src cluster the dst cluster of least
// min(error) (match_segments_descend: first minimum along the dst axis) and keeps it if that error is below
// thres_error.  Here: a 64-bit atomicMin per src cluster on (ordered error bits, dst index), the winning pair recovered
// by a second pass, an ordered compaction over the src clusters -- and the labels that stay unmatched on either side,
// which are the dynamic stage's candidates (utils_match.py:43-47).  One CTA; the work is P + n_src + n_dst items.
struct SelectArgs {
    const long long* pairs;      // [P,2]
    int P;
    const long long* src_unq;    // [ns] sorted
    int ns;
    const long long* dst_unq;    // [nd] sorted
    int nd;
    const float* errors;         // [P,2]
    const float* inliers;
    const float* ratios;
    const float* ious;
    const int* accept;           // [P]
    const float* transforms;     // [P,16]
    float thres_error;
    float* out_rows;             // [ns,10]
    float* out_transforms;       // [ns,16]
    long long* out_src_left;     // [ns]
    long long* out_dst_left;     // [nd]
    int* out_counts;             // [3] rows, src labels left, dst labels left
    unsigned long long* best;    // [ns]  workspace
    int* winner;                 // [ns]
    int* dst_used;               // [nd]
};

__device__ __forceinline__ int find_label(const long long* __restrict__ sorted, int n, long long v) {
    int lo = 0, hi = n;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (sorted[mid] < v) lo = mid + 1; else hi = mid;
    }
    return (lo < n && sorted[lo] == v) ? lo : -1;
}

// torch `errors.min(-1)`: NaN if either entry is NaN
__device__ __forceinline__ float pair_min_error(const float* __restrict__ errors, int p) {
    const float e0 = errors[2 * p], e1 = errors[2 * p + 1];
    return (e0 != e0 || e1 != e1) ? __int_as_float(0x7fc00000) : fminf(e0, e1);
}

// order-preserving bits of a float; NaN sorts first (torch.argmin returns the position of a NaN)
__device__ __forceinline__ unsigned int error_key(float e) {
    if (e != e) return 0u;
    const unsigned int u = __float_as_uint(e);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

// ordered compaction step shared by the three output lists: returns the rank of this thread's item among the kept
// items of the whole block (items before `base` counted in `running`), or -1
__device__ __forceinline__ int block_rank(bool keep, int* wsum, int* running) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned int vote = __ballot_sync(0xffffffffu, keep);
    if (lane == 0) wsum[warp] = __popc(vote);
    __syncthreads();
    int before = *running;
    for (int w = 0; w < warp; ++w) before += wsum[w];
    const int rank = before + __popc(vote & ((1u << lane) - 1u));
    __syncthreads();
    if (tid == 1023) *running = before + __popc(vote);
    __syncthreads();
    return keep ? rank : -1;
}

__global__ void __launch_bounds__(1024) match_select_kernel(SelectArgs a) {
    __shared__ int wsum[32];
    __shared__ int running;
    const int tid = threadIdx.x;
    for (int s = tid; s < a.ns; s += 1024) { a.best[s] = ~0ull; a.winner[s] = -1; }
    for (int d = tid; d < a.nd; d += 1024) a.dst_used[d] = 0;
    __syncthreads();
    for (int pass = 0; pass < 2; ++pass) {
        for (int p = tid; p < a.P; p += 1024) {
            if (a.accept[p] == 0) continue;
            const int si = find_label(a.src_unq, a.ns, a.pairs[2 * p]), di = find_label(a.dst_unq, a.nd, a.pairs[2 * p + 1]);
            if (si < 0 || di < 0) continue;
            const unsigned long long key = ((unsigned long long)error_key(pair_min_error(a.errors, p)) << 32) | (unsigned int)di;
            if (pass == 0) atomicMin(&a.best[si], key);
            else if (key == a.best[si]) atomicMax(&a.winner[si], p);        // (a repeated pair: the last one stands)
        }
        __syncthreads();
    }
    // rows in src-cluster order
    if (tid == 0) running = 0;
    __syncthreads();
    for (int base = 0; base < a.ns; base += 1024) {
        const int s = base + tid;
        int w = -1;
        if (s < a.ns) {
            w = a.winner[s];
            if (w >= 0 && !(pair_min_error(a.errors, w) < a.thres_error)) w = -1;
            a.winner[s] = w;
        }
        const int k = block_rank(w >= 0, wsum, &running);
        if (k >= 0) {
            const int di = (int)(a.best[s] & 0xffffffffull);
            float* r = a.out_rows + (size_t)k * 10;
            r[0] = (float)a.src_unq[s];
            r[1] = (float)a.dst_unq[di];
            r[2] = a.errors[2 * w]; r[3] = a.errors[2 * w + 1];
            r[4] = a.inliers[2 * w]; r[5] = a.inliers[2 * w + 1];
            r[6] = a.ratios[2 * w]; r[7] = a.ratios[2 * w + 1];
            r[8] = a.ious[2 * w]; r[9] = a.ious[2 * w + 1];
            for (int i = 0; i < 16; ++i) a.out_transforms[(size_t)k * 16 + i] = a.transforms[(size_t)w * 16 + i];
            a.dst_used[di] = 1;
        }
    }
    if (tid == 0) { a.out_counts[0] = running; running = 0; }
    __syncthreads();
    for (int base = 0; base < a.ns; base += 1024) {
        const int s = base + tid;
        const int k = block_rank(s < a.ns && a.winner[s] < 0, wsum, &running);
        if (k >= 0) a.out_src_left[k] = a.src_unq[s];
    }
    if (tid == 0) { a.out_counts[1] = running; running = 0; }
    __syncthreads();
    for (int base = 0; base < a.nd; base += 1024) {
        const int d = base + tid;
        const int k = block_rank(d < a.nd && a.dst_used[d] == 0, wsum, &running);
        if (k >= 0) a.out_dst_left[k] = a.dst_unq[d];
    }
    if (tid == 0) a.out_counts[2] = running;
}

size_t match_select_workspace_bytes(int ns, int nd) { return (size_t)ns * 12 + (size_t)nd * 4 + 64; }

int launch_match_select(const int64_t* pairs, int P, const int64_t* src_unq, int ns, const int64_t* dst_unq, int nd,
                        const float* errors, const float* inliers, const float* ratios, const float* ious,
                        const int* accept, const float* transforms, float thres_error, float* out_rows,
                        float* out_transforms, int64_t* out_src_left, int64_t* out_dst_left, int* out_counts,
                        void* workspace, cudaStream_t stream) {
    unsigned char* w = static_cast<unsigned char*>(workspace);
    SelectArgs a{reinterpret_cast<const long long*>(pairs), P, reinterpret_cast<const long long*>(src_unq), ns,
                 reinterpret_cast<const long long*>(dst_unq), nd, errors, inliers, ratios, ious, accept, transforms,
                 thres_error, out_rows, out_transforms, reinterpret_cast<long long*>(out_src_left),
                 reinterpret_cast<long long*>(out_dst_left), out_counts, reinterpret_cast<unsigned long long*>(w),
                 reinterpret_cast<int*>(w + (size_t)ns * 8), reinterpret_cast<int*>(w + (size_t)ns * 12)};
    ICPF_LAUNCH(match_select_kernel, 1, 1024, 0, stream)(a);
    return (int)cudaGetLastError();
}

int launch_gather_pairs(const float* src_points, int src_stride, const int* src_order, const int* src_offsets, int n_src,
                        const float* dst_points, int dst_stride, const int* dst_order, const int* dst_offsets, int n_dst,
                        const int64_t* pairs, int P, int max_points, const int* sample_rows,
                        const int64_t* sample_offsets, float* out_src, float* out_dst, cudaStream_t stream) {
    ScanView s{src_points, src_order, src_offsets, src_stride, n_src};
    ScanView d{dst_points, dst_order, dst_offsets, dst_stride, n_dst};
    ICPF_LAUNCH(gather_pairs_kernel, dim3(P, 2), kScanThreads, 0, stream)(
        s, d, reinterpret_cast<const long long*>(pairs), max_points, sample_rows,
        reinterpret_cast<const long long*>(sample_offsets), reinterpret_cast<float4*>(out_src),
        reinterpret_cast<float4*>(out_dst));
    return (int)cudaGetLastError();
}

int launch_flow(const float* points, int stride, const float* labels, int n, const float* pair_labels, int pair_stride,
                const float* transforms, int K, const float* pose, float* flow, cudaStream_t stream) {
    int blocks = (n + kScanThreads - 1) / kScanThreads;
    blocks = blocks > 148 * 4 ? 148 * 4 : blocks;
    ICPF_LAUNCH(flow_kernel, blocks, kScanThreads, 0, stream)(points, stride, labels, n, pair_labels, pair_stride, transforms, K,
                                                     pose, flow);
    return (int)cudaGetLastError();
}

}  // namespace icpf
