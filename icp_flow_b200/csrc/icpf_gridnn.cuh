// icpf_gridnn.cuh -- exact UNBOUNDED nearest-neighbour distance through the uniform grid of icpf_pair.cuh.
//
// nearest_neighbor_batch (utils_helper.py:20-30) has no radius: the reference scans every row of the other cloud for
// every query (12 passes in estimate_init_pose, 2 in apply_icp, 2 in match_eval).  The quantity those callers consume is
// min_j d(q, c_j) -- a minimum, so the visiting order is irrelevant and any search that provably saw the minimiser
// returns the SAME bits as the full scan as long as each candidate distance is evaluated with the same arithmetic.
// Search (nn_unbounded_grid_warp): blocks of cells around the query per lane, then a warp-wide scan of all rows for
// the queries the blocks could not settle -- the worst case is the reference's scan, spread over the 32 lanes.
#pragma once

#include "icpf_pair.cuh"

namespace icpf {

// lets build_grid() sort an arbitrary row block: rows -> sorted (cell order) + packed run boundaries
struct GridTiles {
    const float4* rows_p;
    float4* sorted_p;
    uint32_t* cells_p;
    float* red_p;
    __device__ __forceinline__ float4* dst() const { return const_cast<float4*>(rows_p); }
    __device__ __forceinline__ float4* sorted() const { return sorted_p; }
    __device__ __forceinline__ uint32_t* cells() const { return cells_p; }
    __device__ __forceinline__ float* red() const { return red_p; }
};

// float4 units of shared memory one grid over N rows needs (sorted copy + run boundaries)
__host__ __device__ inline int gridnn_units(int N) { return N + up16(kCellWords * 4); }

constexpr float kUnbMaxR = 3.5f;      // widest block (half-width in cells) tried before the full scan

// min over the block of half-width r (cells) around the look-up position l of sqdist(e, c + s), c the grid's rows.
//   box  lower bound (m) on the distance from l to every row NOT inspected (+inf when the block covers the grid)
template <bool SHIFT>
__device__ __forceinline__ void grid_block_min(const GridInfo& g, const float4* __restrict__ sorted,
                                               const unsigned short* __restrict__ a, float lx, float ly, float lz,
                                               float ex, float ey, float ez, float sx, float sy, float sz, float r,
                                               float& dmin, float& box) {
    const float INF = __int_as_float(0x7f800000);
    const float fx = (lx - g.ox) * g.inv_c, fy = (ly - g.oy) * g.inv_c, fz = (lz - g.oz) * g.inv_c;
    const float x0 = floorf(fx - r), x1 = floorf(fx + r);
    const float y0 = floorf(fy - r), y1 = floorf(fy + r);
    const float z0 = floorf(fz - r), z1 = floorf(fz + r);
    const float hx = (float)(g.gx - 1), hy = (float)(g.gy - 1), hz = (float)(g.gz - 1);
    if (!(x1 >= 0.f && y1 >= 0.f && z1 >= 0.f && x0 <= hx && y0 <= hy && z0 <= hz)) {
        // the block misses the grid (or the position is NaN): all rows lie inside [0, g]^3 cell coordinates
        const float gapx = fmaxf(-fx, fx - (hx + 1.f)), gapy = fmaxf(-fy, fy - (hy + 1.f)),
                    gapz = fmaxf(-fz, fz - (hz + 1.f));
        box = fmaxf(fmaxf(gapx, fmaxf(gapy, gapz)) * g.c - g.pad, 0.f);
        return;
    }
    const float bx = fminf(x0 < 0.f ? INF : fx - x0, x1 > hx ? INF : x1 + 1.f - fx);
    const float by = fminf(y0 < 0.f ? INF : fy - y0, y1 > hy ? INF : y1 + 1.f - fy);
    const float bz = fminf(z0 < 0.f ? INF : fz - z0, z1 > hz ? INF : z1 + 1.f - fz);
    box = fmaxf(fminf(bx, fminf(by, bz)) * g.c - g.pad, 0.f);
    const int ix0 = max(0, (int)x0), ix1 = min(g.gx - 1, (int)x1);
    const int iy0 = max(0, (int)y0), iy1 = min(g.gy - 1, (int)y1);
    const int iz0 = max(0, (int)z0), iz1 = min(g.gz - 1, (int)z1);
#ifndef ICPF_BLOCK_NESTED_ONLY
    if (ix1 - ix0 <= 1 && iy1 - iy0 <= 1) {
        // the gate block (r < 0.5 cells: at most 2 x 2 columns): its <= 4 runs walked as ONE flat candidate sequence, like
        // grid_search of the ICP loop -- a lane's trip count is its own number of candidates, not the warp's worst column
        const bool two_x = ix1 > ix0, two_y = iy1 > iy0;
        const int b00 = (ix0 * g.gy + iy0) * g.gz, b01 = b00 + g.gz, b10 = b00 + g.gy * g.gz, b11 = b10 + g.gz;
        const int s0 = a[b00 + iz0], e0 = a[b00 + iz1 + 1];
        int s1 = 0, e1 = 0, s2 = 0, e2 = 0, s3 = 0, e3 = 0;
        if (two_y) { s1 = a[b01 + iz0]; e1 = a[b01 + iz1 + 1]; }
        if (two_x) { s2 = a[b10 + iz0]; e2 = a[b10 + iz1 + 1]; }
        if (two_x && two_y) { s3 = a[b11 + iz0]; e3 = a[b11 + iz1 + 1]; }
        const int c1 = e0 - s0, c2 = c1 + (e1 - s1), c3 = c2 + (e2 - s2), total = c3 + (e3 - s3);
        const int o0 = s0, o1 = s1 - c1, o2 = s2 - c2, o3 = s3 - c3;
        for (int t = 0; t < total; ++t) {
            const int off = t < c2 ? (t < c1 ? o0 : o1) : (t < c3 ? o2 : o3);
            const float4 c = sorted[t + off];
            const float d = SHIFT ? sqdist(ex, ey, ez, __fadd_rn(c.x, sx), __fadd_rn(c.y, sy), __fadd_rn(c.z, sz))
                                  : sqdist(ex, ey, ez, c.x, c.y, c.z);
            dmin = fminf(dmin, d);
        }
        return;
    }
#endif
    for (int ix = ix0; ix <= ix1; ++ix) {
        for (int iy = iy0; iy <= iy1; ++iy) {
            const int base = (ix * g.gy + iy) * g.gz;
            const int s = a[base + iz0], e = a[base + iz1 + 1];
            for (int j = s; j < e; ++j) {
                const float4 c = sorted[j];
                const float d = SHIFT ? sqdist(ex, ey, ez, __fadd_rn(c.x, sx), __fadd_rn(c.y, sy), __fadd_rn(c.z, sz))
                                      : sqdist(ex, ey, ez, c.x, c.y, c.z);
                dmin = fminf(dmin, d);
            }
        }
    }
}

// Exact  min_j sqdist(e, rows_j (+ s))  over ALL n rows of the grid -- the value of the reference's full scan -- for the
// queries of one WARP (every lane calls it; `active` says whether the lane has a query).
//   l = where e sits relative to the un-shifted rows (e - s up to rounding; only used to pick cells, the slack g.pad
//   absorbs its rounding);  SHIFT = candidates are evaluated as fadd(row, s), the arithmetic of the scan it replaces.
// The block levels run per lane: the block of cells that just covers the gate radius first; if the best candidate found
// is farther than the nearest un-inspected face (`box`), one more block sized to contain the ball of that candidate
// (which then proves it).  A query the levels cannot settle (nothing within kUnbMaxR cells: wrong candidate
// translations, unrelated clusters, NaN positions) is then scanned by the whole warp -- lane j takes the
// rows j, j + 32, ... of the sorted copy (conflict-free 16-byte loads, four independent minima in flight) and a butterfly
// folds the 32 partial minima.  A minimum does not depend on the visiting order, every candidate is evaluated with the
// scan's arithmetic: the bits of the full scan, at n / 32 candidates per lane instead of a divergent walk over hundreds
// of mostly empty columns (a far query cost ~70 000 cycles in a slab / column search with lower-bound pruning, ~400 here).
// ICPF_COOP_LEVELS=0 scans every query in full, as the reference does (the variant tests/test_simt_variants.py compares with).
#ifndef ICPF_COOP_LEVELS
#define ICPF_COOP_LEVELS 2          // block levels tried per lane before the warp scan (measured: 1 -> 8.66, 2 -> 8.05, 3 -> 8.41 ms on C3)
#endif
template <bool SHIFT>
__device__ __forceinline__ float nn_unbounded_grid_warp(const GridInfo& g, const float4* __restrict__ sorted,
                                                        const unsigned short* __restrict__ a, int n, bool active,
                                                        float lx, float ly, float lz, float ex, float ey, float ez,
                                                        float sx, float sy, float sz) {
    const float INF = __int_as_float(0x7f800000);
    float dmin = INF;
    bool open = active;
    if (active) {
        float box = 0.f, r = g.r;
        for (int lvl = 0; lvl < ICPF_COOP_LEVELS; ++lvl) {
            grid_block_min<SHIFT>(g, sorted, a, lx, ly, lz, ex, ey, ez, sx, sy, sz, r, dmin, box);
            if (sqrtf(dmin) * 1.0001f + 1e-6f <= box) { open = false; break; }
            const float want = (dmin < INF) ? sqrtf(dmin) * g.inv_c * 1.001f + 0.02f : r + 1.0f;
            if (!(want <= kUnbMaxR) || !(want > r)) break;
            r = want;
        }
    }
    unsigned int pend = __ballot_sync(FULL_MASK, open);
    const int lane = threadIdx.x & 31;
    while (pend != 0u) {
        const int owner = __ffs(pend) - 1;
        pend &= pend - 1u;
        const float qx = __shfl_sync(FULL_MASK, ex, owner), qy = __shfl_sync(FULL_MASK, ey, owner),
                    qz = __shfl_sync(FULL_MASK, ez, owner);
        float tx = 0.f, ty = 0.f, tz = 0.f;
        if (SHIFT) {
            tx = __shfl_sync(FULL_MASK, sx, owner); ty = __shfl_sync(FULL_MASK, sy, owner); tz = __shfl_sync(FULL_MASK, sz, owner);
        }
        float m0 = INF, m1 = INF, m2 = INF, m3 = INF;
        auto dist = [&](int j) -> float {
            const float4 c = sorted[j];
            return SHIFT ? sqdist(qx, qy, qz, __fadd_rn(c.x, tx), __fadd_rn(c.y, ty), __fadd_rn(c.z, tz))
                         : sqdist(qx, qy, qz, c.x, c.y, c.z);
        };
        int j = lane;
        for (; j + 96 < n; j += 128) {
            m0 = fminf(m0, dist(j)); m1 = fminf(m1, dist(j + 32)); m2 = fminf(m2, dist(j + 64)); m3 = fminf(m3, dist(j + 96));
        }
        for (; j < n; j += 32) m0 = fminf(m0, dist(j));
        float m = fminf(fminf(m0, m1), fminf(m2, m3));
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m = fminf(m, __shfl_xor_sync(FULL_MASK, m, o));
        if (lane == owner) dmin = fminf(dmin, m);
    }
    return dmin;
}

}  // namespace icpf
