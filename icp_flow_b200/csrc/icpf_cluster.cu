// icpf_cluster.cu -- DBSCAN over one scan (the clustering that produces the labels the
// per-pair path consumes).
//
// Reference: utils_cluster.cluster_dbscan (/root/reference/utils_cluster.py:32-48) = Open3D's PointCloud.cluster_dbscan
// (eps = args.epsilon, min_points = args.min_cluster_size), a sequential region growing over pre-computed radius
// neighbourhoods.  Its result is a deterministic function of the points, which is what is computed here in parallel:
//   core      a point whose eps-ball (the point itself included) holds >= min_points points;
//   cluster   a connected component of the core points under "within eps"; clusters are numbered in the order the
//             sequential scan meets them, i.e. by their LOWEST core-point index;
//   border    a non-core point takes the cluster of the first cluster that reaches it -- clusters are grown one after
//             the other, so that is the LOWEST-numbered cluster with a core point within eps; otherwise noise (-1).
// Distances are evaluated in fp64 on the fp32 coordinates exactly as a kd-tree radius query does (dx*dx + dy*dy +
// dz*dz in that order, compared with eps*eps), so the labels equal sklearn's DBSCAN / Open3D's on the same points (the
// test oracle is sklearn: Open3D is not installable here; the two differ only for a pair of points EXACTLY eps apart).
//
// Structure: uniform grid of cells >= eps (dense table sized on the device from the bounding box), counting sort of
// the points by cell, three passes over the 27 neighbouring cells (count -> core flags, lock-free union-find over the
// core-core edges with "larger root hooks under smaller root" so that a component's root is its lowest core index,
// labels), two exclusive scans.  Integer / gather work on O(10^5) points, latency- and L2-bound.
#include "icpf_internal.h"
#include "icpf_common.cuh"

namespace icpf {

namespace {

constexpr int kDbThreads = 256;
constexpr int kDbMaxCells = 1 << 23;          // dense cell table (int32): at most 32 MB of workspace
constexpr int kScanBlock = 1024;

struct DbHeader {                              // lives at the start of the workspace
    int lo_bits[3], hi_bits[3];                // bounding box of the finite points (order-preserving int encoding)
    float lo[3];
    double inv_cell;
    int dim[3];
    int ncell;
    int scan_len;                              // ncell + 1: entries of the cell table the scan covers
    int num_clusters;
};

__device__ __forceinline__ int float_to_ordered(float f) {
    const int i = __float_as_int(f);
    return i >= 0 ? i : i ^ 0x7fffffff;
}
__device__ __forceinline__ float ordered_to_float(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff); }
__device__ __forceinline__ bool finite3(float x, float y, float z) {
    return (fabsf(x) < 3.0e38f) && (fabsf(y) < 3.0e38f) && (fabsf(z) < 3.0e38f);      // false for NaN / inf
}

__global__ void db_init_kernel(DbHeader* h) {
    for (int k = 0; k < 3; ++k) {
        h->lo_bits[k] = 0x7fffffff;
        h->hi_bits[k] = (int)0x80000000;
    }
    h->num_clusters = 0;
}

__global__ void __launch_bounds__(kDbThreads) db_bbox_kernel(const float* __restrict__ pts, int stride, int n, DbHeader* h) {
    __shared__ int s_lo[3], s_hi[3];
    if (threadIdx.x < 3) { s_lo[threadIdx.x] = 0x7fffffff; s_hi[threadIdx.x] = (int)0x80000000; }
    __syncthreads();
    int lo[3] = {0x7fffffff, 0x7fffffff, 0x7fffffff}, hi[3] = {(int)0x80000000, (int)0x80000000, (int)0x80000000};
    for (int i = blockIdx.x * kDbThreads + threadIdx.x; i < n; i += gridDim.x * kDbThreads) {
        const float x = pts[(size_t)i * stride], y = pts[(size_t)i * stride + 1], z = pts[(size_t)i * stride + 2];
        if (!finite3(x, y, z)) continue;
        const int v[3] = {float_to_ordered(x), float_to_ordered(y), float_to_ordered(z)};
        for (int k = 0; k < 3; ++k) { lo[k] = min(lo[k], v[k]); hi[k] = max(hi[k], v[k]); }
    }
    for (int k = 0; k < 3; ++k) { atomicMin(&s_lo[k], lo[k]); atomicMax(&s_hi[k], hi[k]); }
    __syncthreads();
    if (threadIdx.x < 3) { atomicMin(&h->lo_bits[threadIdx.x], s_lo[threadIdx.x]); atomicMax(&h->hi_bits[threadIdx.x], s_hi[threadIdx.x]); }
}

// grid geometry: cells of eps * 2^k (k as small as the table allows): two points within eps are in adjacent cells
__global__ void db_setup_kernel(DbHeader* h, double eps, int max_cells) {
    double cell = eps * (1.0 + 1e-9);
    double ext[3];
    bool empty = false;
    for (int k = 0; k < 3; ++k) {
        if (h->lo_bits[k] > h->hi_bits[k]) empty = true;
        h->lo[k] = empty ? 0.f : ordered_to_float(h->lo_bits[k]);
        ext[k] = empty ? 0.0 : (double)ordered_to_float(h->hi_bits[k]) - (double)h->lo[k];
    }
    for (int it = 0; it < 64; ++it) {
        double cells = 1.0;
        for (int k = 0; k < 3; ++k) cells *= floor(ext[k] / cell) + 1.0;
        if (cells <= (double)max_cells) break;
        cell *= 2.0;
    }
    h->inv_cell = 1.0 / cell;
    int nc = 1;
    for (int k = 0; k < 3; ++k) {
        h->dim[k] = (int)floor(ext[k] / cell) + 1;
        nc *= h->dim[k];
    }
    h->ncell = nc;
    h->scan_len = nc + 1;
}

__device__ __forceinline__ int db_cell(const DbHeader* h, float x, float y, float z, int& cx, int& cy, int& cz) {
    cx = min(h->dim[0] - 1, max(0, (int)floor(((double)x - (double)h->lo[0]) * h->inv_cell)));
    cy = min(h->dim[1] - 1, max(0, (int)floor(((double)y - (double)h->lo[1]) * h->inv_cell)));
    cz = min(h->dim[2] - 1, max(0, (int)floor(((double)z - (double)h->lo[2]) * h->inv_cell)));
    return (cx * h->dim[1] + cy) * h->dim[2] + cz;
}

// cell of every point (-1: non-finite, never anybody's neighbour) and the per-cell counts at table[cell + 1]
__global__ void __launch_bounds__(kDbThreads) db_count_kernel(const float* __restrict__ pts, int stride, int n, const DbHeader* h,
                                                              int* __restrict__ cell_of, int* __restrict__ table) {
    const int i = blockIdx.x * kDbThreads + threadIdx.x;
    if (i >= n) return;
    const float x = pts[(size_t)i * stride], y = pts[(size_t)i * stride + 1], z = pts[(size_t)i * stride + 2];
    int c = -1;
    if (finite3(x, y, z)) {
        int cx, cy, cz;
        c = db_cell(h, x, y, z, cx, cy, cz);
        atomicAdd(&table[c + 1], 1);
    }
    cell_of[i] = c;
}

// ---- exclusive scan of an int array in place, three launches; `len_dev` (may be NULL) = length known on the device only
__global__ void __launch_bounds__(kScanBlock) scan_reduce_kernel(const int* __restrict__ a, int len, const int* len_dev,
                                                                 int* __restrict__ block_sums) {
    __shared__ int wsum[32];
    const int L = len_dev ? *len_dev : len;
    const int i = blockIdx.x * kScanBlock + threadIdx.x;
    int v = (i < L) ? a[i] : 0;
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL_MASK, v, o);
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x < 32) {
        int s = wsum[threadIdx.x];
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(FULL_MASK, s, o);
        if (threadIdx.x == 0) block_sums[blockIdx.x] = s;
    }
}

__global__ void __launch_bounds__(kScanBlock) scan_blocksums_kernel(int* __restrict__ block_sums, int nblocks, int* total) {
    __shared__ int wsum[32];
    __shared__ int running;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) running = 0;
    __syncthreads();
    for (int base = 0; base < nblocks; base += kScanBlock) {
        const int i = base + tid;
        const int v = (i < nblocks) ? block_sums[i] : 0;
        int incl = v;
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(FULL_MASK, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) wsum[warp] = incl;
        __syncthreads();
        int woff = 0;
        for (int w = 0; w < warp; ++w) woff += wsum[w];
        const int excl = running + woff + incl - v;
        if (i < nblocks) block_sums[i] = excl;
        __syncthreads();
        if (tid == kScanBlock - 1) running = excl + v;
        __syncthreads();
    }
    if (tid == 0 && total) *total = running;
}

__global__ void __launch_bounds__(kScanBlock) scan_apply_kernel(int* __restrict__ a, int len, const int* len_dev,
                                                                const int* __restrict__ block_sums) {
    __shared__ int wsum[32];
    const int L = len_dev ? *len_dev : len;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int i = blockIdx.x * kScanBlock + tid;
    if (blockIdx.x * kScanBlock >= L) return;
    const int v = (i < L) ? a[i] : 0;
    int incl = v;
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(FULL_MASK, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) wsum[warp] = incl;
    __syncthreads();
    int woff = 0;
    for (int w = 0; w < warp; ++w) woff += wsum[w];
    if (i < L) a[i] = block_sums[blockIdx.x] + woff + incl - v;
}

void exclusive_scan(int* a, int len_max, const int* len_dev, int* block_sums, int* total, cudaStream_t stream) {
    const int nb = (len_max + kScanBlock - 1) / kScanBlock;
    ICPF_LAUNCH(scan_reduce_kernel, nb, kScanBlock, 0, stream)(a, len_max, len_dev, block_sums);
    ICPF_LAUNCH(scan_blocksums_kernel, 1, kScanBlock, 0, stream)(block_sums, nb, total);
    ICPF_LAUNCH(scan_apply_kernel, nb, kScanBlock, 0, stream)(a, len_max, len_dev, block_sums);
}

// counting sort: order[pos] = point index, grouped by cell; afterwards cell c = order[table[c] .. table[c + 1])
__global__ void __launch_bounds__(kDbThreads) db_scatter_kernel(const int* __restrict__ cell_of, int n, int* __restrict__ table,
                                                                int* __restrict__ order) {
    const int i = blockIdx.x * kDbThreads + threadIdx.x;
    if (i >= n) return;
    const int c = cell_of[i];
    if (c < 0) return;
    order[atomicAdd(&table[c + 1], 1)] = i;
}

// squared distance exactly as a kd-tree radius query evaluates it on float64 copies of the fp32 coordinates
__device__ __forceinline__ double db_dist2(float ax, float ay, float az, float bx, float by, float bz) {
    const double dx = (double)ax - (double)bx, dy = (double)ay - (double)by, dz = (double)az - (double)bz;
    return __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
}

// visit every point j within eps of point i (i itself included); `visit(j)` returns false to stop early
template <class Visit>
__device__ __forceinline__ void db_neighbours(const float* __restrict__ pts, int stride, const DbHeader* h,
                                              const int* __restrict__ table, const int* __restrict__ order, int i, double eps2,
                                              Visit visit) {
    const float x = pts[(size_t)i * stride], y = pts[(size_t)i * stride + 1], z = pts[(size_t)i * stride + 2];
    int cx, cy, cz;
    db_cell(h, x, y, z, cx, cy, cz);
    const int x0 = max(cx - 1, 0), x1 = min(cx + 1, h->dim[0] - 1);
    const int y0 = max(cy - 1, 0), y1 = min(cy + 1, h->dim[1] - 1);
    const int z0 = max(cz - 1, 0), z1 = min(cz + 1, h->dim[2] - 1);
    for (int ix = x0; ix <= x1; ++ix) {
        for (int iy = y0; iy <= y1; ++iy) {
            // the z cells of a column are consecutive table entries: one run of the sorted order
            const int base = (ix * h->dim[1] + iy) * h->dim[2];
            const int s = table[base + z0], e = table[base + z1 + 1];
            for (int k = s; k < e; ++k) {
                const int j = order[k];
                const float qx = pts[(size_t)j * stride], qy = pts[(size_t)j * stride + 1], qz = pts[(size_t)j * stride + 2];
                if (db_dist2(x, y, z, qx, qy, qz) <= eps2) {
                    if (!visit(j)) return;
                }
            }
        }
    }
}

__global__ void __launch_bounds__(kDbThreads) db_core_kernel(const float* __restrict__ pts, int stride, int n, const DbHeader* h,
                                                             const int* __restrict__ table, const int* __restrict__ order,
                                                             const int* __restrict__ cell_of, double eps2, int min_points,
                                                             int* __restrict__ parent, unsigned char* __restrict__ core) {
    const int t = blockIdx.x * kDbThreads + threadIdx.x;
    if (t >= n) return;
    // threads walk the points in cell order: neighbouring threads read neighbouring cells
    const int nsorted = table[h->ncell];
    if (t >= nsorted) return;
    const int i = order[t];
    int cnt = 0;
    db_neighbours(pts, stride, h, table, order, i, eps2, [&](int) { return ++cnt < min_points; });
    core[i] = cnt >= min_points ? 1 : 0;
    parent[i] = i;
    (void)cell_of;
}

__device__ __forceinline__ int db_find(int* parent, int i) {
    // path halving; concurrent hooks only ever lower a parent, so a stale read still leads towards a root (reads and
    // writes go to L2: another SM's hook must become visible inside this launch)
    int p = __ldcg(parent + i);
    while (p != i) {
        const int g = __ldcg(parent + p);
        if (g != p) __stcg(parent + i, g);
        i = p;
        p = g;
    }
    return i;
}

// larger root hooks under smaller root: the root of a component ends up as its lowest point index
__device__ __forceinline__ void db_union(int* parent, int a, int b) {
    while (true) {
        a = db_find(parent, a);
        b = db_find(parent, b);
        if (a == b) return;
        if (a < b) { const int t = a; a = b; b = t; }
        const int old = atomicMin(&parent[a], b);          // a was a root (parent[a] == a) unless somebody hooked it meanwhile
        if (old == a) return;
        a = old;
    }
}

__global__ void __launch_bounds__(kDbThreads) db_union_kernel(const float* __restrict__ pts, int stride, int n, const DbHeader* h,
                                                              const int* __restrict__ table, const int* __restrict__ order,
                                                              double eps2, int* __restrict__ parent,
                                                              const unsigned char* __restrict__ core) {
    const int t = blockIdx.x * kDbThreads + threadIdx.x;
    if (t >= n || t >= table[h->ncell]) return;
    const int i = order[t];
    if (!core[i]) return;
    db_neighbours(pts, stride, h, table, order, i, eps2, [&](int j) {
        if (j < i && core[j]) db_union(parent, i, j);
        return true;
    });
}

// roots of the core components; flag[i] = 1 for a root (scanned into the cluster numbers afterwards)
__global__ void __launch_bounds__(kDbThreads) db_roots_kernel(int n, int* __restrict__ parent, const unsigned char* __restrict__ core,
                                                              const int* __restrict__ cell_of, int* __restrict__ flag) {
    const int i = blockIdx.x * kDbThreads + threadIdx.x;
    if (i >= n) return;
    int f = 0;
    if (cell_of[i] >= 0 && core[i]) {
        const int r = db_find(parent, i);
        parent[i] = r;
        f = (r == i) ? 1 : 0;
    }
    flag[i] = f;
}

__global__ void __launch_bounds__(kDbThreads) db_label_kernel(const float* __restrict__ pts, int stride, int n, const DbHeader* h,
                                                              const int* __restrict__ table, const int* __restrict__ order,
                                                              const int* __restrict__ cell_of, double eps2,
                                                              const int* __restrict__ parent, const unsigned char* __restrict__ core,
                                                              const int* __restrict__ cluster_of_root, int* __restrict__ labels) {
    const int i = blockIdx.x * kDbThreads + threadIdx.x;
    if (i >= n) return;
    int lab = -1;
    if (cell_of[i] >= 0) {
        if (core[i]) {
            lab = cluster_of_root[parent[i]];
        } else {
            // border point: the lowest-numbered cluster with a core point within eps
            int best = 0x7fffffff;
            db_neighbours(pts, stride, h, table, order, i, eps2, [&](int j) {
                if (core[j]) best = min(best, cluster_of_root[parent[j]]);
                return true;
            });
            lab = (best == 0x7fffffff) ? -1 : best;
        }
    }
    labels[i] = lab;
}

struct DbWs {
    DbHeader* h;
    int* table;        // [db_table_cells(n) + 2]
    int* cell_of;      // [n]
    int* order;        // [n]
    int* parent;       // [n]
    int* flag;         // [n + 1]
    int* block_sums;   // scan scratch
    unsigned char* core;   // [n]
};

size_t db_align(size_t v) { return (v + 255) / 256 * 256; }

// cells of the table for a scan of n points: 64 per point is far more than a scan occupies, 8 M bounds the workspace
int db_table_cells(int n) {
    const long long want = 64ll * (n > 0 ? n : 0);
    return (int)(want < 4096 ? 4096 : (want > kDbMaxCells ? kDbMaxCells : want));
}

}  // namespace

size_t dbscan_workspace_bytes(int n) {
    const size_t nn = (size_t)(n > 0 ? n : 0);
    const size_t cells = (size_t)db_table_cells(n);
    return db_align(sizeof(DbHeader)) + db_align((cells + 2) * 4) + 4 * db_align(nn * 4 + 4) +
           db_align((cells / kScanBlock + 2 + nn / kScanBlock) * 4) + db_align(nn);
}

int launch_dbscan(const float* points, int stride, int n, double eps, int min_points, int* out_labels, int* out_num_clusters,
                  void* workspace, size_t workspace_bytes, cudaStream_t stream) {
    if (n == 0) {
        if (out_num_clusters) return (int)cudaMemsetAsync(out_num_clusters, 0, sizeof(int), stream);
        return ICPF_OK;
    }
    if (workspace == nullptr || workspace_bytes < dbscan_workspace_bytes(n)) return ICPF_E_WORKSPACE;
    unsigned char* w = static_cast<unsigned char*>(workspace);
    DbWs ws;
    ws.h = reinterpret_cast<DbHeader*>(w); w += db_align(sizeof(DbHeader));
    const int cells = db_table_cells(n);
    ws.table = reinterpret_cast<int*>(w); w += db_align(((size_t)cells + 2) * 4);
    ws.cell_of = reinterpret_cast<int*>(w); w += db_align((size_t)n * 4 + 4);
    ws.order = reinterpret_cast<int*>(w); w += db_align((size_t)n * 4 + 4);
    ws.parent = reinterpret_cast<int*>(w); w += db_align((size_t)n * 4 + 4);
    ws.flag = reinterpret_cast<int*>(w); w += db_align((size_t)n * 4 + 4);
    ws.block_sums = reinterpret_cast<int*>(w); w += db_align(((size_t)cells / kScanBlock + 2 + (size_t)n / kScanBlock) * 4);
    ws.core = w;
    const int nb = (n + kDbThreads - 1) / kDbThreads;
    const double eps2 = eps * eps;
    cudaError_t err = cudaMemsetAsync(ws.table, 0, ((size_t)cells + 2) * 4, stream);
    if (err != cudaSuccess) return (int)err;
    err = cudaMemsetAsync(ws.core, 0, (size_t)n, stream);
    if (err != cudaSuccess) return (int)err;
    ICPF_LAUNCH(db_init_kernel, 1, 1, 0, stream)(ws.h);
    ICPF_LAUNCH(db_bbox_kernel, nb < 1024 ? nb : 1024, kDbThreads, 0, stream)(points, stride, n, ws.h);
    ICPF_LAUNCH(db_setup_kernel, 1, 1, 0, stream)(ws.h, eps, cells);
    ICPF_LAUNCH(db_count_kernel, nb, kDbThreads, 0, stream)(points, stride, n, ws.h, ws.cell_of, ws.table);
    // table[0, ncell + 1): counts at cell + 1 -> exclusive scan -> first sorted position of every cell
    exclusive_scan(ws.table, cells + 1, &ws.h->scan_len, ws.block_sums, nullptr, stream);
    ICPF_LAUNCH(db_scatter_kernel, nb, kDbThreads, 0, stream)(ws.cell_of, n, ws.table, ws.order);
    ICPF_LAUNCH(db_core_kernel, nb, kDbThreads, 0, stream)(points, stride, n, ws.h, ws.table, ws.order, ws.cell_of, eps2,
                                                          min_points, ws.parent, ws.core);
    ICPF_LAUNCH(db_union_kernel, nb, kDbThreads, 0, stream)(points, stride, n, ws.h, ws.table, ws.order, eps2, ws.parent, ws.core);
    ICPF_LAUNCH(db_roots_kernel, nb, kDbThreads, 0, stream)(n, ws.parent, ws.core, ws.cell_of, ws.flag);
    exclusive_scan(ws.flag, n, nullptr, ws.block_sums, out_num_clusters ? out_num_clusters : &ws.h->num_clusters, stream);
    ICPF_LAUNCH(db_label_kernel, nb, kDbThreads, 0, stream)(points, stride, n, ws.h, ws.table, ws.order, ws.cell_of, eps2,
                                                           ws.parent, ws.core, ws.flag, out_labels);
    return (int)cudaGetLastError();
}

}  // namespace icpf
