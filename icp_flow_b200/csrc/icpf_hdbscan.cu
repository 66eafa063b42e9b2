// icpf_hdbscan.cu -- HDBSCAN over one scan (the clustering every script of the reference
// selects: utils_cluster.cluster_hdbscan, /root/reference/utils_cluster.py:10-29, main.sh / demo.sh --if_hdbscan).
//
// The reference calls the `hdbscan` package (0.8.x, environment.yml; not vendored, not installable here) with
// min_samples = None (-> min_cluster_size), alpha = 1, metric euclidean, cluster_selection_method 'eom',
// approx_min_span_tree = True.  What is restated is the published algorithm (Campello, Moulavi, Sander 2013; McInnes,
// Healy, Astels 2017) with an EXACT minimum spanning tree; the test oracle is scikit-learn's port of the same package
// (sklearn.cluster.HDBSCAN, which is installed) -- parity with the reference's own dependency is therefore UNPINNED
// and stated as such in DESIGN.md.
//
//   core distance      distance to the min_samples-th nearest point, the point itself included (fp64 on the fp32
//                      coordinates: dx*dx + dy*dy + dz*dz in that order, sqrt -- what a kd-tree query returns)
//   mutual reachability  max(core_a, core_b, d(a, b))
//   MST                of the complete mutual-reachability graph; single-linkage dendrogram of its sorted edges
//   condensed tree     a split whose two sides both hold >= min_cluster_size points makes two new clusters, a smaller
//                      side falls out of its cluster; stability = sum over the points of (lambda_leave - lambda_birth),
//                      lambda = 1 / distance; excess-of-mass selection, the root never selected; a point belongs to
//                      the selected cluster it falls out of (or out of a descendant of), else noise.
//
// GPU: the two O(n^2) stages.  Core distances: every thread keeps the k smallest squared distances of its point (k
// doubles per thread in shared memory) while all points stream through a shared-memory tile.  MST: Prim's algorithm
// exactly as the oracle runs it (sklearn/cluster/_hdbscan/_linkage.pyx, mst_from_data_matrix: start at point 0, strict
// `<` updates, the first minimum in index order joins) -- every thread relaxes one candidate against the point that
// just joined, the blocks' arg-minima are folded, the winner joins; all steps in one cooperative launch (or one launch
// per step when the scan is too large for every block to be resident).  The edge
// sequence, not only the tree, is reproduced: equal weights are everywhere in a mutual-reachability graph (every
// edge into a point of large core distance weighs that core distance), and which of them the dendrogram merges first
// decides whether a small group falls out of a cluster or splits it, so a different MST order (Boruvka, a different
// tie rule) changes a few labels per scan.  With the oracle's order, its arithmetic (fp64, dx*dx + dy*dy + dz*dz in
// that order, no contraction) and its sort of the edges (numpy argsort, called by the host wrapper) the labels are
// identical.  ~3.3 us per step: 0.17 s for 5*10^4 points, against 9 s for the oracle on one host core.
// The O(n alpha(n)) bookkeeping on the n - 1 MST edges (union-find dendrogram, condensation, selection) runs on the
// host, as it does in the reference.
#include "icpf_internal.h"
#include "icpf_common.cuh"

#ifndef ICPF_SIMT_EMU
#include <cooperative_groups.h>
#endif
#include <algorithm>
#include <cmath>
#include <limits>
#include <numeric>
#include <vector>

namespace icpf {

// ------------------------------------------------------------------------------------------------ host: MST -> labels
namespace {

struct HdbEdge {
    double w;
    int a, b;
};

struct UnionFind {
    std::vector<int> parent;
    explicit UnionFind(int n) : parent(n) { std::iota(parent.begin(), parent.end(), 0); }
    int find(int x) {
        int r = x;
        while (parent[r] != r) r = parent[r];
        while (parent[x] != r) { const int nx = parent[x]; parent[x] = r; x = nx; }
        return r;
    }
};

}  // namespace

// labels[n]: clusters numbered by their lowest point index, -1 = noise
int hdbscan_labels_host(const int* edge_a, const int* edge_b, const double* edge_w, int n, int min_cluster_size,
                        int presorted, int* labels) {
    if (n <= 0) return ICPF_OK;
    for (int i = 0; i < n; ++i) labels[i] = -1;
    if (n == 1) return ICPF_OK;
    const int m = n - 1;
    std::vector<HdbEdge> e(m);
    for (int i = 0; i < m; ++i) {
        if (edge_a[i] < 0 || edge_a[i] >= n || edge_b[i] < 0 || edge_b[i] >= n || !(edge_w[i] >= 0.0)) return ICPF_E_PARAM;
        e[i] = HdbEdge{edge_w[i], std::min(edge_a[i], edge_b[i]), std::max(edge_a[i], edge_b[i])};
        if (presorted && i > 0 && e[i].w < e[i - 1].w) return ICPF_E_PARAM;
    }
    if (!presorted) std::sort(e.begin(), e.end(), [](const HdbEdge& x, const HdbEdge& y) {
        if (x.w != y.w) return x.w < y.w;
        if (x.a != y.a) return x.a < y.a;
        return x.b < y.b;
    });
    // single-linkage dendrogram: leaves 0 .. n-1, internal node n + i made by edge i
    std::vector<int> left(m), right(m), size(m);
    {
        UnionFind uf(n);
        std::vector<int> node_of(n);          // dendrogram node of a union-find root
        std::vector<int> cnt(n, 1);
        std::iota(node_of.begin(), node_of.end(), 0);
        for (int i = 0; i < m; ++i) {
            const int ra = uf.find(e[i].a), rb = uf.find(e[i].b);
            if (ra == rb) return ICPF_E_PARAM;                      // not a spanning tree
            left[i] = node_of[ra];
            right[i] = node_of[rb];
            size[i] = cnt[ra] + cnt[rb];
            uf.parent[rb] = ra;
            cnt[ra] = size[i];
            node_of[ra] = n + i;
        }
    }
    auto node_size = [&](int v) { return v < n ? 1 : size[v - n]; };
    // condensed tree, top-down.  Cluster ids grow from the root (0), so children always have larger ids.
    std::vector<int> c_parent{-1};
    std::vector<double> c_birth{0.0}, c_stab{0.0};
    std::vector<int> point_cluster(n, 0);
    std::vector<std::pair<int, int>> stack;            // (dendrogram node, cluster)
    std::vector<int> sub;                              // scratch: subtree walk
    auto fall_out = [&](int v, int cluster, double lambda) {
        sub.clear();
        sub.push_back(v);
        while (!sub.empty()) {
            const int u = sub.back();
            sub.pop_back();
            if (u < n) {
                point_cluster[u] = cluster;
                c_stab[cluster] += lambda - c_birth[cluster];
            } else {
                sub.push_back(left[u - n]);
                sub.push_back(right[u - n]);
            }
        }
    };
    stack.emplace_back(n + m - 1, 0);
    while (!stack.empty()) {
        const int v = stack.back().first, c = stack.back().second;
        stack.pop_back();
        if (v < n) {                                   // a single point left in its cluster: it never leaves
            point_cluster[v] = c;
            continue;
        }
        const int l = left[v - n], r = right[v - n];
        const double d = e[v - n].w;
        const double lambda = d > 0.0 ? 1.0 / d : std::numeric_limits<double>::infinity();
        const int nl = node_size(l), nr = node_size(r);
        if (nl >= min_cluster_size && nr >= min_cluster_size) {
            for (int side = 0; side < 2; ++side) {
                const int child = side ? r : l, cs = side ? nr : nl;
                const int id = (int)c_parent.size();
                c_parent.push_back(c);
                c_birth.push_back(lambda);
                c_stab.push_back(0.0);
                c_stab[c] += (lambda - c_birth[c]) * (double)cs;
                stack.emplace_back(child, id);
            }
        } else if (nl < min_cluster_size && nr < min_cluster_size) {
            fall_out(l, c, lambda);
            fall_out(r, c, lambda);
        } else if (nl < min_cluster_size) {
            fall_out(l, c, lambda);
            stack.emplace_back(r, c);
        } else {
            fall_out(r, c, lambda);
            stack.emplace_back(l, c);
        }
    }
    // excess of mass, bottom-up (children have larger ids); the root is never a cluster
    const int nc = (int)c_parent.size();
    std::vector<double> child_sum(nc, 0.0);
    std::vector<char> selected(nc, 1);
    selected[0] = 0;
    for (int c = nc - 1; c >= 1; --c) {
        if (child_sum[c] > c_stab[c]) {
            selected[c] = 0;
            c_stab[c] = child_sum[c];
        }
        child_sum[c_parent[c]] += c_stab[c];
    }
    // a selected cluster unselects its descendants: top-down, the nearest selected ancestor-or-self wins
    std::vector<int> owner(nc, -1);
    for (int c = 1; c < nc; ++c) {
        const int up = owner[c_parent[c]];
        owner[c] = up >= 0 ? up : (selected[c] ? c : -1);
    }
    std::vector<int> label_of(nc, -1);
    int next = 0;
    for (int i = 0; i < n; ++i) {
        const int o = owner[point_cluster[i]];
        if (o < 0) continue;
        if (label_of[o] < 0) label_of[o] = next++;
        labels[i] = label_of[o];
    }
    return ICPF_OK;
}

// ------------------------------------------------------------------------------------------------ device: core distances
constexpr int kHdbThreads = 128;
constexpr int kHdbMaxK = 64;            // min_samples the shared-memory heaps hold
constexpr int kPrimThreads = 256;

__device__ __forceinline__ double hdb_sqdist(double ax, double ay, double az, double bx, double by, double bz) {
    const double dx = ax - bx, dy = ay - by, dz = az - bz;
    return __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
}

__global__ void __launch_bounds__(256) hdb_to_double_kernel(const float* __restrict__ pts, int stride, int n,
                                                            double* __restrict__ p64) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    p64[3 * (size_t)i] = (double)pts[(size_t)i * stride];
    p64[3 * (size_t)i + 1] = (double)pts[(size_t)i * stride + 1];
    p64[3 * (size_t)i + 2] = (double)pts[(size_t)i * stride + 2];
}

// core[i] = sqrt of the k-th smallest squared distance from point i to the points of the scan, itself included
__global__ void __launch_bounds__(kHdbThreads) hdb_core_kernel(const double* __restrict__ p64, int n, int k,
                                                               double* __restrict__ core) {
    ICPF_DYN_SHARED __align__(16) float4 fsm[];              // (the dynamic shared array of icpf_histfused.cu, as doubles)
    double* hsm = reinterpret_cast<double*>(fsm);
    double* heap = hsm + threadIdx.x;                         // heap[j * kHdbThreads]: column of this thread
    double* tile = hsm + (size_t)k * kHdbThreads;             // [kHdbThreads * 3]
    const int i = blockIdx.x * kHdbThreads + threadIdx.x;
    const bool live = i < n;
    const double INF = __longlong_as_double(0x7ff0000000000000LL);
    double qx = 0.0, qy = 0.0, qz = 0.0;
    if (live) { qx = p64[3 * (size_t)i]; qy = p64[3 * (size_t)i + 1]; qz = p64[3 * (size_t)i + 2]; }
    for (int j = 0; j < k; ++j) heap[(size_t)j * kHdbThreads] = INF;
    double worst = INF;            // largest of the k kept
    int worst_at = 0;
    for (int base = 0; base < n; base += kHdbThreads) {
        __syncthreads();
        const int t = base + threadIdx.x;
        if (t < n) {
            tile[3 * threadIdx.x] = p64[3 * (size_t)t];
            tile[3 * threadIdx.x + 1] = p64[3 * (size_t)t + 1];
            tile[3 * threadIdx.x + 2] = p64[3 * (size_t)t + 2];
        }
        __syncthreads();
        const int cnt = min(kHdbThreads, n - base);
        if (live) {
            for (int c = 0; c < cnt; ++c) {
                const double d = hdb_sqdist(qx, qy, qz, tile[3 * c], tile[3 * c + 1], tile[3 * c + 2]);
                if (d < worst) {
                    heap[(size_t)worst_at * kHdbThreads] = d;
                    worst = -1.0;
                    for (int j = 0; j < k; ++j) {
                        const double v = heap[(size_t)j * kHdbThreads];
                        if (v > worst) { worst = v; worst_at = j; }
                    }
                }
            }
        }
    }
    if (live) core[i] = sqrt(worst);
}

// One step of Prim's algorithm (see the header).  part_v / part_j: per block the smallest key of the previous step.
struct PrimArgs {
    const double* p64;
    const double* core;
    int n;
    double* min_reach;       // [n]
    int* source;             // [n]
    unsigned char* in_tree;  // [n]
    double* part_v;          // [2][blocks]
    int* part_j;             // [2][blocks]
    int blocks;
    int* edge_src;           // [n-1]
    int* edge_dst;
    double* edge_w;
    int* cur_node;           // [1] the point that joined last
};

__device__ __forceinline__ bool prim_less(double va, int ja, double vb, int jb) { return va < vb || (va == vb && ja < jb); }

__device__ __forceinline__ void hdb_prim_step(const PrimArgs& a, int step, int last) {
    __shared__ double s_v[kPrimThreads / 32];
    __shared__ int s_j[kPrimThreads / 32];
    __shared__ int s_cur;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const double BIG = 1.7976931348623157e308;      // DBL_MAX: the oracle's initial `new_reachability`
    // ---- the point that joins now: 0 at the first step, else the arg-min of the previous step's candidates
    if (warp == 0) {
        int cur = 0;
        if (step > 0) {
            const double* pv = a.part_v + (size_t)((step - 1) & 1) * a.blocks;
            const int* pj = a.part_j + (size_t)((step - 1) & 1) * a.blocks;
            double bv = BIG;
            int bj = 0x7fffffff;
            for (int b = lane; b < a.blocks; b += 32) {
                if (prim_less(pv[b], pj[b], bv, bj)) { bv = pv[b]; bj = pj[b]; }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const double ov = __shfl_xor_sync(FULL_MASK, bv, o);
                const int oj = __shfl_xor_sync(FULL_MASK, bj, o);
                if (prim_less(ov, oj, bv, bj)) { bv = ov; bj = oj; }
            }
            // (no candidate below DBL_MAX -- non-finite input -- : the oracle's `new_node` stays 0)
            cur = (bj == 0x7fffffff) ? 0 : bj;
            if (blockIdx.x == 0 && lane == 0) {
                a.edge_src[step - 1] = (bj == 0x7fffffff) ? 0 : a.source[cur];
                a.edge_dst[step - 1] = cur;
                a.edge_w[step - 1] = bv;
                a.in_tree[cur] = 1;
                *a.cur_node = cur;
            }
        } else if (blockIdx.x == 0 && lane == 0) {
            a.in_tree[0] = 1;
        }
        if (lane == 0) s_cur = cur;
    }
    __syncthreads();
    if (last) return;
    const int cur = s_cur;
    const double cx = a.p64[3 * (size_t)cur], cy = a.p64[3 * (size_t)cur + 1], cz = a.p64[3 * (size_t)cur + 2];
    const double ccore = a.core[cur];
    // ---- relax one candidate per thread against `cur`
    const int j = blockIdx.x * kPrimThreads + tid;
    double v = BIG;
    int vj = 0x7fffffff;
    if (j < a.n && j != cur && a.in_tree[j] == 0) {
        const double d = sqrt(hdb_sqdist(cx, cy, cz, a.p64[3 * (size_t)j], a.p64[3 * (size_t)j + 1], a.p64[3 * (size_t)j + 2]));
        const double cj = a.core[j];
        double mr = ccore > cj ? ccore : cj;
        mr = mr > d ? mr : d;
        double cand = a.min_reach[j];
        if (mr < cand) {
            a.min_reach[j] = mr;
            a.source[j] = cur;
            cand = mr;
        }
        if (cand < BIG) { v = cand; vj = j; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double ov = __shfl_xor_sync(FULL_MASK, v, o);
        const int oj = __shfl_xor_sync(FULL_MASK, vj, o);
        if (prim_less(ov, oj, v, vj)) { v = ov; vj = oj; }
    }
    if (lane == 0) { s_v[warp] = v; s_j[warp] = vj; }
    __syncthreads();
    if (tid == 0) {
        for (int w = 1; w < kPrimThreads / 32; ++w) {
            if (prim_less(s_v[w], s_j[w], v, vj)) { v = s_v[w]; vj = s_j[w]; }
        }
        a.part_v[(size_t)(step & 1) * a.blocks + blockIdx.x] = v;
        a.part_j[(size_t)(step & 1) * a.blocks + blockIdx.x] = vj;
    }
}

__global__ void __launch_bounds__(kPrimThreads) hdb_prim_step_kernel(PrimArgs a, int step, int last) {
    hdb_prim_step(a, step, last);
}

#ifndef ICPF_SIMT_EMU
// 16-byte words as ONE memory operation each (PTX scalar type .b128: single-copy atomic, unlike a v4 vector access,
// which the memory model treats as four).  Every word carries the tag of the step that published it, so a reader needs
// no ordering between words: it re-reads a word until the tag is the one it waits for.
__device__ __forceinline__ void hdb_publish(uint4* p, unsigned long long lo, unsigned int z, unsigned int tag) {
    asm volatile("{\n.reg .b128 t;\nmov.b128 t, {%1, %2};\nst.relaxed.gpu.global.b128 [%0], t;\n}" ::"l"(p), "l"(lo),
                 "l"(((unsigned long long)tag << 32) | z)
                 : "memory");
}
__device__ __forceinline__ bool hdb_observe(const uint4* p, unsigned int tag, unsigned long long& lo, unsigned int& z) {
    unsigned long long hi;
    asm volatile("{\n.reg .b128 t;\nld.relaxed.gpu.global.b128 t, [%2];\nmov.b128 {%0, %1}, t;\n}" : "=l"(lo), "=l"(hi) : "l"(p) : "memory");
    z = (unsigned int)hi;
    return (unsigned int)(hi >> 32) == tag;
}

// All steps in ONE cooperative launch (every block resident) when the scan fits the device, without grid-wide barriers:
//   every block publishes its best candidate as five tagged words (value + index, x, y, z, core distance);
//   block 0 polls the words of all blocks (one L2 round trip once they are there), folds them and publishes the winner --
//   index and coordinates -- as four tagged words; every block polls those four.
// Two L2 round trips and two small reductions per step (~2 us; a grid barrier plus the dependent loads was ~4.3 us, a
// launch per step ~6 us).  Words alternate between two sets: a word tagged t is overwritten with tag t + 2, which its
// writer can only reach after every reader of tag t has moved on.
constexpr int kPartWords = 5, kWinWords = 4;

__global__ void __launch_bounds__(kPrimThreads) hdb_prim_coop_kernel(PrimArgs a, uint4* __restrict__ words) {
    __shared__ double s_v[kPrimThreads / 32], s_x[kPrimThreads / 32][4];
    __shared__ int s_j[kPrimThreads / 32];
    __shared__ int s_cur, s_win;
    __shared__ double s_cpt[4];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const double BIG = 1.7976931348623157e308;
    const double INF = __longlong_as_double(0x7ff0000000000000LL);
    uint4* part = words;                                                     // [2][blocks][kPartWords]
    uint4* win = words + (size_t)2 * a.blocks * kPartWords;                  // [2][kWinWords]
    // the candidate this thread owns for the whole run lives in registers
    const int j = blockIdx.x * kPrimThreads + tid;
    const bool have = j < a.n;
    double px = 0.0, py = 0.0, pz = 0.0, cj = 0.0, reach = INF;
    int src = 1;
    bool in_tree = !have;
    if (have) {
        px = a.p64[3 * (size_t)j]; py = a.p64[3 * (size_t)j + 1]; pz = a.p64[3 * (size_t)j + 2];
        cj = a.core[j];
    }
    for (int step = 0; step < a.n; ++step) {
        const unsigned int tag = (unsigned int)step;                         // the words published during step - 1
        if (step > 0 && blockIdx.x == 0) {
            // ---- fold the candidates of all blocks
            const uint4* pw = part + (size_t)((step - 1) & 1) * a.blocks * kPartWords;
            double bv = BIG, bx[4] = {0.0, 0.0, 0.0, 0.0};
            int bj = 0x7fffffff;
            for (int b = tid; b < a.blocks; b += kPrimThreads) {
                unsigned long long lo[kPartWords];
                unsigned int z[kPartWords];
                bool ok[kPartWords];
                bool all = false;
                while (!all) {
                    all = true;
#pragma unroll
                    for (int k = 0; k < kPartWords; ++k) {
                        ok[k] = hdb_observe(pw + (size_t)b * kPartWords + k, tag, lo[k], z[k]);
                        all = all && ok[k];
                    }
                }
                const double v = __longlong_as_double((long long)lo[0]);
                const int vj = (int)z[0];
                if (prim_less(v, vj, bv, bj)) {
                    bv = v; bj = vj;
#pragma unroll
                    for (int k = 0; k < 4; ++k) bx[k] = __longlong_as_double((long long)lo[1 + k]);
                }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const double ov = __shfl_xor_sync(FULL_MASK, bv, o);
                const int oj = __shfl_xor_sync(FULL_MASK, bj, o);
                double ox[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) ox[k] = __shfl_xor_sync(FULL_MASK, bx[k], o);
                if (prim_less(ov, oj, bv, bj)) {
                    bv = ov; bj = oj;
#pragma unroll
                    for (int k = 0; k < 4; ++k) bx[k] = ox[k];
                }
            }
            if (lane == 0) {
                s_v[warp] = bv; s_j[warp] = bj;
#pragma unroll
                for (int k = 0; k < 4; ++k) s_x[warp][k] = bx[k];
            }
            __syncthreads();
            if (tid == 0) {
                int wbest = 0;
                for (int w = 1; w < kPrimThreads / 32; ++w) {
                    if (prim_less(s_v[w], s_j[w], s_v[wbest], s_j[wbest])) wbest = w;
                }
                const int cur = (s_j[wbest] == 0x7fffffff) ? 0 : s_j[wbest];
                a.edge_dst[step - 1] = cur;
                a.edge_w[step - 1] = s_v[wbest];
                uint4* ww = win + (size_t)((step - 1) & 1) * kWinWords;
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    hdb_publish(ww + k, (unsigned long long)__double_as_longlong(s_x[wbest][k]), (unsigned int)cur, tag);
            }
            __syncthreads();
        }
        // ---- the point that joins now and its coordinates
        if (warp == 0 && lane < 4) {
            if (step == 0) {
                s_cpt[lane] = lane < 3 ? a.p64[lane] : a.core[0];
                if (lane == 0) s_cur = 0;
            } else {
                const uint4* ww = win + (size_t)((step - 1) & 1) * kWinWords + lane;
                unsigned long long lo;
                unsigned int z;
                while (!hdb_observe(ww, tag, lo, z)) {}
                s_cpt[lane] = __longlong_as_double((long long)lo);
                if (lane == 0) s_cur = (int)z;
            }
        }
        __syncthreads();
        const int cur = s_cur;
        if (j == cur) {
            in_tree = true;
            if (step > 0) a.edge_src[step - 1] = src;          // the owner knows which point reached it
        }
        if (step == a.n - 1) break;
        double v = BIG;
        int vj = 0x7fffffff;
        if (!in_tree) {
            const double d = sqrt(hdb_sqdist(s_cpt[0], s_cpt[1], s_cpt[2], px, py, pz));
            const double ccore = s_cpt[3];
            double mr = ccore > cj ? ccore : cj;
            mr = mr > d ? mr : d;
            if (mr < reach) { reach = mr; src = cur; }
            if (reach < BIG) { v = reach; vj = j; }
        }
        const double own_v = v;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double ov = __shfl_xor_sync(FULL_MASK, v, o);
            const int oj = __shfl_xor_sync(FULL_MASK, vj, o);
            if (prim_less(ov, oj, v, vj)) { v = ov; vj = oj; }
        }
        if (lane == 0) { s_v[warp] = v; s_j[warp] = vj; }
        __syncthreads();
        if (tid == 0) {
            for (int w = 1; w < kPrimThreads / 32; ++w) {
                if (prim_less(s_v[w], s_j[w], v, vj)) { v = s_v[w]; vj = s_j[w]; }
            }
            s_win = vj;
        }
        __syncthreads();
        // ---- the block's best candidate publishes itself (it holds its coordinates in registers)
        uint4* pw = part + ((size_t)(step & 1) * a.blocks + blockIdx.x) * kPartWords;
        const unsigned int ntag = (unsigned int)(step + 1);
        if (s_win == 0x7fffffff) {
            if (tid == 0) {
                hdb_publish(pw, (unsigned long long)__double_as_longlong(BIG), 0x7fffffffu, ntag);
                for (int k = 1; k < kPartWords; ++k) hdb_publish(pw + k, 0ull, 0u, ntag);
            }
        } else if (have && !in_tree && j == s_win) {
            hdb_publish(pw, (unsigned long long)__double_as_longlong(own_v), (unsigned int)j, ntag);
            hdb_publish(pw + 1, (unsigned long long)__double_as_longlong(px), 0u, ntag);
            hdb_publish(pw + 2, (unsigned long long)__double_as_longlong(py), 0u, ntag);
            hdb_publish(pw + 3, (unsigned long long)__double_as_longlong(pz), 0u, ntag);
            hdb_publish(pw + 4, (unsigned long long)__double_as_longlong(cj), 0u, ntag);
        }
        // (the shared words of the next step are written behind its barriers; s_win is read before anyone can pass two)
        __syncthreads();
    }
}
#endif

__global__ void __launch_bounds__(256) hdb_init_kernel(int n, double* min_reach, int* source, unsigned char* in_tree) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    min_reach[i] = __longlong_as_double(0x7ff0000000000000LL);
    source[i] = 1;                      // (np.ones in the oracle; only read after it was written)
    in_tree[i] = 0;
}

// ------------------------------------------------------------------------------------------------ device: Boruvka (any-order tree)
// The minimum spanning tree under the strict total order (weight, min(a,b), max(a,b)) on the edges is unique, so any
// algorithm may build it: Boruvka rounds -- every point finds its lightest edge into another component (all points
// stream through shared memory, fp64), a 64-bit atomicMin per component picks the component's lightest, the chosen
// edges are added and the components merged.  <= log2 n rounds of n^2 candidate edges instead of n - 1 dependent steps:
// ~20 ms instead of ~160 ms for 5*10^4 points.  The tree has the oracle's weights but not its ORDER among equal weights,
// so the dendrogram -- and a few labels per scan -- may differ from the oracle's where its own result depends on that
// order (icpf_hdbscan_mst_f32: prim_order = 0).
struct BorArgs {
    const double* p64;
    const double* core;
    int n;
    int* comp;                    // [n] component (its root point) of every point
    int* comp_next;               // [n]
    int* parent;                  // [n] hook target of a root
    unsigned long long* best_w;   // [n] per root: bits of the lightest outgoing weight
    unsigned long long* best_e;   // [n] per root: (min << 32 | max) of the lightest outgoing edge
    double* cand_w;               // [n] per point
    int* cand_j;                  // [n]
    int* edge_src;
    int* edge_dst;
    double* edge_w;
    int* counters;                // [0] edges emitted, [1] components left
};

__device__ __forceinline__ bool bor_edge_less(int a0, int b0, int a1, int b1) {        // (min, max) lexicographic
    const int lo0 = min(a0, b0), hi0 = max(a0, b0), lo1 = min(a1, b1), hi1 = max(a1, b1);
    return lo0 < lo1 || (lo0 == lo1 && hi0 < hi1);
}

__global__ void __launch_bounds__(256) hdb_bor_init_kernel(BorArgs a, int first) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= a.n) return;
    if (first) a.comp[i] = i;
    a.best_w[i] = ~0ull;
    a.best_e[i] = ~0ull;
    if (i == 0) a.counters[1] = 0;
    if (i == 0 && first) a.counters[0] = 0;
}

__global__ void __launch_bounds__(kHdbThreads) hdb_bor_minedge_kernel(BorArgs a) {
    __shared__ double t_x[kHdbThreads], t_y[kHdbThreads], t_z[kHdbThreads], t_c[kHdbThreads];
    __shared__ int t_comp[kHdbThreads];
    const int i = blockIdx.x * kHdbThreads + threadIdx.x;
    const bool live = i < a.n;
    const double INF = __longlong_as_double(0x7ff0000000000000LL);
    double qx = 0.0, qy = 0.0, qz = 0.0, qc = 0.0, bw = INF, lim2 = INF;
    int ci = -1, bj = -1;
    if (live) {
        qx = a.p64[3 * (size_t)i]; qy = a.p64[3 * (size_t)i + 1]; qz = a.p64[3 * (size_t)i + 2];
        qc = a.core[i];
        ci = a.comp[i];
    }
    for (int base = 0; base < a.n; base += kHdbThreads) {
        __syncthreads();
        const int t = base + threadIdx.x;
        if (t < a.n) {
            t_x[threadIdx.x] = a.p64[3 * (size_t)t]; t_y[threadIdx.x] = a.p64[3 * (size_t)t + 1]; t_z[threadIdx.x] = a.p64[3 * (size_t)t + 2];
            t_c[threadIdx.x] = a.core[t];
            t_comp[threadIdx.x] = a.comp[t];
        }
        __syncthreads();
        const int cnt = min(kHdbThreads, a.n - base);
        if (live) {
            for (int c = 0; c < cnt; ++c) {
                if (t_comp[c] == ci) continue;
                // a candidate whose squared distance is clearly above the best weight squared cannot win or tie: the
                // square root (most of the cost of a candidate) is only taken for the few that can
                const double d2 = hdb_sqdist(qx, qy, qz, t_x[c], t_y[c], t_z[c]);
                if (d2 > lim2 || t_c[c] > bw) continue;
                const double d = sqrt(d2);
                double mr = qc > t_c[c] ? qc : t_c[c];
                mr = mr > d ? mr : d;
                if (mr < bw || (mr == bw && bor_edge_less(i, base + c, i, bj))) {
                    bw = mr; bj = base + c;
                    lim2 = bw * bw * 1.000000000000001 + 1e-300;
                }
            }
        }
    }
    if (live) {
        a.cand_w[i] = bw;
        a.cand_j[i] = bj;
        if (bj >= 0) atomicMin(&a.best_w[ci], (unsigned long long)__double_as_longlong(bw));     // weights >= 0: bits are monotone
    }
}

__global__ void __launch_bounds__(256) hdb_bor_pick_kernel(BorArgs a) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= a.n) return;
    const int j = a.cand_j[i];
    if (j < 0) return;
    const int ci = a.comp[i];
    if ((unsigned long long)__double_as_longlong(a.cand_w[i]) != a.best_w[ci]) return;
    const unsigned long long key = ((unsigned long long)(unsigned int)min(i, j) << 32) | (unsigned int)max(i, j);
    atomicMin(&a.best_e[ci], key);
}

__global__ void __launch_bounds__(256) hdb_bor_hook_kernel(BorArgs a) {
    const int c = blockIdx.x * 256 + threadIdx.x;
    if (c >= a.n) return;
    if (a.comp[c] != c) return;                        // roots only
    const unsigned long long e = a.best_e[c];
    if (e == ~0ull) { a.parent[c] = c; return; }       // (a single component left)
    const int u = (int)(e >> 32), v = (int)(e & 0xffffffffull);
    const int t = a.comp[a.comp[u] == c ? v : u];
    const bool mutual = a.best_e[t] == e;              // both components chose this edge: emit it once, the smaller id stays root
    if (!mutual || c < t) {
        const int at = atomicAdd(&a.counters[0], 1);
        if (at < a.n - 1) {
            a.edge_src[at] = u;
            a.edge_dst[at] = v;
            a.edge_w[at] = __longlong_as_double((long long)a.best_w[c]);
        }
    }
    a.parent[c] = (mutual && c < t) ? c : t;
}

__global__ void __launch_bounds__(256) hdb_bor_jump_kernel(BorArgs a) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= a.n) return;
    int r = a.comp[i];
    while (a.parent[r] != r) r = a.parent[r];          // (hooks form trees: under a strict edge order only mutual pairs cycle)
    a.comp_next[i] = r;
    if (r == i) atomicAdd(&a.counters[1], 1);
}

inline size_t hdb_up(size_t b) { return (b + 255) / 256 * 256; }

size_t hdbscan_workspace_bytes(int n) {
    const size_t blocks = ((size_t)n + kPrimThreads - 1) / kPrimThreads;
    return hdb_up((size_t)n * 24) + hdb_up((size_t)n * 8) + hdb_up((size_t)n * 4) + hdb_up((size_t)n) +
           hdb_up(blocks * 16) + hdb_up(blocks * 8) + 256 + hdb_up((blocks * 2 * 5 + 2 * 4) * 16) +
           3 * hdb_up((size_t)n * 8) + 4 * hdb_up((size_t)n * 4) + 256;      // Boruvka: best_w, best_e, cand_w | comp, comp_next, parent, cand_j
}

int launch_hdbscan_mst(const float* points, int stride, int n, int min_samples, int prim_order, double* out_core,
                       int* out_src, int* out_dst, double* out_w, void* workspace, cudaStream_t stream) {
    if (min_samples > kHdbMaxK) return ICPF_E_UNSUPPORTED;
    unsigned char* w = static_cast<unsigned char*>(workspace);
    const int blocks = (n + kHdbThreads - 1) / kHdbThreads;
    const int pblocks = (n + kPrimThreads - 1) / kPrimThreads;
    double* p64 = reinterpret_cast<double*>(w); w += hdb_up((size_t)n * 24);
    double* min_reach = reinterpret_cast<double*>(w); w += hdb_up((size_t)n * 8);
    int* source = reinterpret_cast<int*>(w); w += hdb_up((size_t)n * 4);
    unsigned char* in_tree = w; w += hdb_up((size_t)n);
    double* part_v = reinterpret_cast<double*>(w); w += hdb_up((size_t)pblocks * 16);
    int* part_j = reinterpret_cast<int*>(w); w += hdb_up((size_t)pblocks * 8);
    int* cur_node = reinterpret_cast<int*>(w); w += 256;
    void* words_raw = w;
    const int b256 = (n + 255) / 256;
    ICPF_LAUNCH(hdb_to_double_kernel, b256, 256, 0, stream)(points, stride, n, p64);
    const int k = min_samples < n ? min_samples : n;
    const size_t smem = ((size_t)k + 3) * kHdbThreads * sizeof(double);
    cudaError_t err = cudaFuncSetAttribute(hdb_core_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess) return (int)err;
    ICPF_LAUNCH(hdb_core_kernel, blocks, kHdbThreads, smem, stream)(p64, n, k, out_core);
    ICPF_LAUNCH(hdb_init_kernel, b256, 256, 0, stream)(n, min_reach, source, in_tree);
    if (!prim_order && n > 1) {
        unsigned char* wb = static_cast<unsigned char*>(words_raw) + hdb_up(((size_t)pblocks * 2 * 5 + 2 * 4) * 16);
        BorArgs b;
        b.p64 = p64; b.core = out_core; b.n = n;
        b.best_w = reinterpret_cast<unsigned long long*>(wb); wb += hdb_up((size_t)n * 8);
        b.best_e = reinterpret_cast<unsigned long long*>(wb); wb += hdb_up((size_t)n * 8);
        b.cand_w = reinterpret_cast<double*>(wb); wb += hdb_up((size_t)n * 8);
        b.comp = reinterpret_cast<int*>(wb); wb += hdb_up((size_t)n * 4);
        b.comp_next = reinterpret_cast<int*>(wb); wb += hdb_up((size_t)n * 4);
        b.parent = reinterpret_cast<int*>(wb); wb += hdb_up((size_t)n * 4);
        b.cand_j = reinterpret_cast<int*>(wb); wb += hdb_up((size_t)n * 4);
        b.counters = reinterpret_cast<int*>(wb);
        b.edge_src = out_src; b.edge_dst = out_dst; b.edge_w = out_w;
        int host_counters[2] = {0, n};
        for (int round = 0; round < 64 && host_counters[1] > 1; ++round) {
            ICPF_LAUNCH(hdb_bor_init_kernel, b256, 256, 0, stream)(b, round == 0 ? 1 : 0);
            ICPF_LAUNCH(hdb_bor_minedge_kernel, blocks, kHdbThreads, 0, stream)(b);
            ICPF_LAUNCH(hdb_bor_pick_kernel, b256, 256, 0, stream)(b);
            ICPF_LAUNCH(hdb_bor_hook_kernel, b256, 256, 0, stream)(b);
            ICPF_LAUNCH(hdb_bor_jump_kernel, b256, 256, 0, stream)(b);
            // the number of rounds depends on the data: this mode reads the component count back every round
            err = cudaMemcpyAsync(host_counters, b.counters, sizeof(host_counters), cudaMemcpyDeviceToHost, stream);
            if (err != cudaSuccess) return (int)err;
            err = cudaStreamSynchronize(stream);
            if (err != cudaSuccess) return (int)err;
            int* t = b.comp; b.comp = b.comp_next; b.comp_next = t;
        }
        if (host_counters[1] != 1 || host_counters[0] != n - 1) return ICPF_E_PARAM;       // non-finite input: no spanning tree
        return (int)cudaGetLastError();
    }
    PrimArgs a{p64, out_core, n, min_reach, source, in_tree, part_v, part_j, pblocks, out_src, out_dst, out_w, cur_node};
#ifndef ICPF_SIMT_EMU
    {
        int dev = 0, coop = 0, sms = 0, per_sm = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, hdb_prim_coop_kernel, kPrimThreads, 0);
        if (coop && (long long)per_sm * sms >= pblocks) {
            uint4* words = static_cast<uint4*>(words_raw);
            err = cudaMemsetAsync(words, 0, ((size_t)pblocks * 2 * kPartWords + 2 * kWinWords) * 16, stream);   // tag 0: nothing published
            if (err != cudaSuccess) return (int)err;
            void* args[] = {&a, &words};
            err = cudaLaunchCooperativeKernel((const void*)hdb_prim_coop_kernel, dim3(pblocks), dim3(kPrimThreads), args, 0, stream);
            if (err != cudaErrorCooperativeLaunchTooLarge) return (int)err;
            (void)cudaGetLastError();              // (the device is shared right now: one launch per step instead)
        }
    }
#endif
    for (int step = 0; step < n; ++step) {            // step n-1 only records the last edge
        ICPF_LAUNCH(hdb_prim_step_kernel, pblocks, kPrimThreads, 0, stream)(a, step, step == n - 1 ? 1 : 0);
    }
    return (int)cudaGetLastError();
}

}  // namespace icpf
