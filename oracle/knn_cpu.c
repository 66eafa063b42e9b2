/*
 * oracle/knn_cpu.c -- TEST INFRASTRUCTURE (CPU checker), not product code.
 *
 * Restates the K=1 brute-force nearest-neighbour leaf that the reference path
 * reaches through pytorch3d 0.7.4 `knn_points` (third-party, not vendored in
 * /root/reference; pinned by /root/reference/environment.yml:141).  Call sites
 * in the reference: utils_icp_pytorch3d.py:154-156 (with lengths) and
 * utils_helper.py:27 (no lengths).  Published semantics followed here:
 *   - squared L2 distance accumulated in coordinate order  d = dx*dx; d += dy*dy; d += dz*dz
 *   - candidates scanned 0..len2-1 in order, replaced on strict '<'  (ties -> lowest index)
 *   - query rows >= len1 keep dist = 0 and idx = 0
 * Build with -ffp-contract=off so no FMA contraction changes the rounding.
 *
 * Layout: p1 [B, P1, 3] and p2 [B, P2, 3] contiguous fp32; outputs [B, P1].
 */
#include <stdint.h>
#include <stddef.h>

void icpf_oracle_knn1(const float *p1, const float *p2,
                      const int64_t *len1, const int64_t *len2,
                      int64_t B, int64_t P1, int64_t P2,
                      float *out_d2, int64_t *out_idx)
{
    #pragma omp parallel for schedule(dynamic, 1)
    for (int64_t b = 0; b < B; ++b) {
        const float *q = p1 + (size_t)b * P1 * 3;
        const float *c = p2 + (size_t)b * P2 * 3;
        const int64_t n1 = len1 ? len1[b] : P1;
        const int64_t n2 = len2 ? len2[b] : P2;
        float *od = out_d2 + (size_t)b * P1;
        int64_t *oi = out_idx + (size_t)b * P1;
        for (int64_t i = 0; i < P1; ++i) { od[i] = 0.0f; oi[i] = 0; }
        for (int64_t i = 0; i < n1; ++i) {
            const float qx = q[3 * i], qy = q[3 * i + 1], qz = q[3 * i + 2];
            float best = 0.0f; int64_t bi = 0; int have = 0;
            for (int64_t j = 0; j < n2; ++j) {
                float dx = qx - c[3 * j], dy = qy - c[3 * j + 1], dz = qz - c[3 * j + 2];
                float d = dx * dx;
                d = d + dy * dy;
                d = d + dz * dz;
                if (!have || d < best) { best = d; bi = j; have = 1; }
            }
            if (have) { od[i] = best; oi[i] = bi; }
        }
    }
}

/*
 * All-pairs difference histogram, restating /root/reference/hist_cuda/cpp/hist_cuda_core.cuh:35-62
 * (one vote per (i,j) with both flags > 0, half-open range test, floor of
 * (v-min)/(max-min)*len in fp32) and the zero-filled [B,lx,ly,lz] fp32 output of
 * hist_cuda.cu:59.  X,Y are [B,NX,4] / [B,NY,4] rows (x,y,z,flag).
 */
#include <math.h>
void icpf_oracle_hist(const float *X, const float *Y, int64_t B, int64_t NX, int64_t NY,
                      float min_x, float min_y, float min_z,
                      float max_x, float max_y, float max_z,
                      int64_t lx, int64_t ly, int64_t lz, float *bins)
{
    const float rx = max_x - min_x, ry = max_y - min_y, rz = max_z - min_z;
    const float flx = (float)lx, fly = (float)ly, flz = (float)lz;
    #pragma omp parallel for schedule(dynamic, 1)
    for (int64_t b = 0; b < B; ++b) {
        float *h = bins + (size_t)b * lx * ly * lz;
        for (int64_t k = 0; k < lx * ly * lz; ++k) h[k] = 0.0f;
        const float *xb = X + (size_t)b * NX * 4;
        const float *yb = Y + (size_t)b * NY * 4;
        for (int64_t i = 0; i < NX; ++i) {
            if (!(xb[4 * i + 3] > 0.0f)) continue;
            for (int64_t j = 0; j < NY; ++j) {
                if (!(yb[4 * j + 3] > 0.0f)) continue;
                float vx = xb[4 * i] - yb[4 * j];
                float vy = xb[4 * i + 1] - yb[4 * j + 1];
                float vz = xb[4 * i + 2] - yb[4 * j + 2];
                if (vx >= min_x && vx < max_x && vy >= min_y && vy < max_y && vz >= min_z && vz < max_z) {
                    float fx = (vx - min_x) / rx; fx = fx * flx;
                    float fy = (vy - min_y) / ry; fy = fy * fly;
                    float fz = (vz - min_z) / rz; fz = fz * flz;
                    int64_t px = (int64_t)floorf(fx), py = (int64_t)floorf(fy), pz = (int64_t)floorf(fz);
                    /* For v one ulp below max, (v-min) rounds up to (max-min) and the reference kernel computes bin
                     * index == len: an out-of-bounds write (hist_cuda_core.cuh:54-59 has no guard).  There is no
                     * defined behaviour to restate; the oracle and the engine both clamp to the last bin. */
                    if (px > lx - 1) px = lx - 1;
                    if (py > ly - 1) py = ly - 1;
                    if (pz > lz - 1) pz = lz - 1;
                    h[px * ly * lz + py * lz + pz] += 1.0f;
                }
            }
        }
    }
}
