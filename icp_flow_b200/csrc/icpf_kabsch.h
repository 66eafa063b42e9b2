// icpf_kabsch.h -- closed-form 3x3 Kabsch rotation, usable from device code and (for CPU unit tests of the
// arithmetic) from host code.
#pragma once

#include <math.h>

#if defined(__CUDACC__)
#define ICPF_HD __host__ __device__
#define ICPF_INLINE __forceinline__
#else
#define ICPF_HD
#define ICPF_INLINE inline
#endif

namespace icpf {

ICPF_HD ICPF_INLINE float icpf_rsqrt(float x) {
#if defined(__CUDA_ARCH__)
    return rsqrtf(x);
#else
    return 1.0f / sqrtf(x);
#endif
}

ICPF_HD ICPF_INLINE float icpf_fast_div(float a, float b) {
#if defined(__CUDA_ARCH__)
    return __fdividef(a, b);
#else
    return a / b;
#endif
}

// ------------------------------------------------------------------------------------------------ 3x3 Kabsch
// Rotation of /root/reference/utils_icp_pytorch3d.py:339-363 in closed form.
//
// Input  H = Xc^T Yc / W  (row-vector convention, Y ~ X R + T), row-major h[3*i+j].
// Reference: U,S,V = svd(H); R = U diag(1,1,det(U V^T)) V^T.
// Here: one-sided (Hestenes) Jacobi finds an orthogonal V with B = H V column-orthogonal (B = U Sigma).  With a,b
// the two dominant columns,  R = u_a v_a^T + u_b v_b^T + (u_a x u_b)(v_a x v_b)^T  -- algebraically identical to the
// determinant-corrected product (the flipped singular vector is always the one of the smallest singular value) and
// needs neither the third column nor a determinant, so rank-2 (planar cluster) inputs are handled exactly like the
// reference.  H == 0 (no inliers) gives R = I, as torch.svd does for the zero matrix.
//
// Warm start: consecutive ICP iterations solve nearly the same problem, so the V of the previous solve already
// orthogonalises H V to ~1e-3; the Jacobi sweeps then only polish (1-2 sweeps instead of 4-5).  The state is left
// untouched when no rotation was needed, so an unchanged H reproduces the previous R bit for bit (the ICP loop's
// fixed-point test relies on that); a slightly wider threshold for *starting* to rotate absorbs the rounding noise
// of recomputing H V.  The inner arithmetic uses approximate divide / rsqrt: Jacobi is self-correcting and every
// output vector is re-orthonormalised at the end.
struct Rot3 {
    float r[9];  // row-major, row-vector convention: x' = x R
};

struct KabschState {
    float v[9];   // right singular vectors of the previous solve, row-major (column c = v[c], v[3+c], v[6+c])
    bool warm;
    bool changed; // set by every solve: did it replace the stored frame?
};

template <int p, int q>
ICPF_HD ICPF_INLINE void jacobi_pair(float (&b)[9], float (&v)[9], float thr2, bool& rotated, float& max_t) {
    // columns p,q of b (3x3 row-major): b[3*i+p]
    const float bp0 = b[p], bp1 = b[3 + p], bp2 = b[6 + p];
    const float bq0 = b[q], bq1 = b[3 + q], bq2 = b[6 + q];
    const float alpha = fmaf(bp2, bp2, fmaf(bp1, bp1, bp0 * bp0));
    const float beta = fmaf(bq2, bq2, fmaf(bq1, bq1, bq0 * bq0));
    const float gamma = fmaf(bp2, bq2, fmaf(bp1, bq1, bp0 * bq0));
    // converged for this pair when the columns are orthogonal to working precision: gamma^2 <= thr^2 alpha beta
    if (!(gamma * gamma > thr2 * (alpha * beta))) return;
    rotated = true;
    float t, c;
#ifndef ICPF_KABSCH_EXACT_ANGLE
    // warm-started solves rotate by ~1e-3 rad: |zeta| is huge and t = 1 / (2 zeta) (1 - 1 / (4 zeta^2) + ...), c = 1 - t^2 / 2
    // to fp32 precision; the exact formulas (two divisions, a square root, a reciprocal square root) for the rest
    const float diff = beta - alpha;
    if (fabsf(diff) > 64.0f * fabsf(gamma)) {
        t = icpf_fast_div(gamma, diff);
        t = fmaf(-t * t, t, t);                    // t (1 - t^2): the next term of the series
        c = fmaf(-0.5f * t, t, 1.0f);
        max_t = fmaxf(max_t, fabsf(t));
    } else
#endif
    {
        const float zeta = icpf_fast_div(beta - alpha, 2.0f * gamma);
        const float az = fabsf(zeta);
        t = copysignf(1.0f, zeta) * icpf_fast_div(1.0f, az + sqrtf(fmaf(az, az, 1.0f)));
        c = icpf_rsqrt(fmaf(t, t, 1.0f));
        max_t = 1.0f;
    }
    const float s = c * t;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const float x = b[3 * i + p], y = b[3 * i + q];
        b[3 * i + p] = fmaf(c, x, -s * y);
        b[3 * i + q] = fmaf(s, x, c * y);
        const float vx = v[3 * i + p], vy = v[3 * i + q];
        v[3 * i + p] = fmaf(c, vx, -s * vy);
        v[3 * i + q] = fmaf(s, vx, c * vy);
    }
}

template <int p, int q>
ICPF_HD ICPF_INLINE void swap_cols(float (&b)[9], float (&v)[9]) {
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        float t = b[3 * i + p]; b[3 * i + p] = b[3 * i + q]; b[3 * i + q] = t;
        t = v[3 * i + p]; v[3 * i + p] = v[3 * i + q]; v[3 * i + q] = t;
    }
}

ICPF_HD ICPF_INLINE void normalize3(float& x, float& y, float& z, float n2) {
    const float inv = icpf_rsqrt(n2);
    x *= inv;
    y *= inv;
    z *= inv;
}

ICPF_HD inline Rot3 kabsch_rotation(const float (&h)[9], KabschState* st = nullptr) {
    Rot3 out;
    float b[9], v[9];
    // scale to unit magnitude so the thresholds below are relative (also avoids under/overflow of the squares)
    float amax = 0.f;
#pragma unroll
    for (int i = 0; i < 9; ++i) amax = fmaxf(amax, fabsf(h[i]));
    if (!(amax > 0.f) || !(amax < 3.0e38f)) {  // zero, NaN or inf cross-covariance -> identity
#pragma unroll
        for (int i = 0; i < 9; ++i) out.r[i] = (i % 4 == 0) ? 1.f : 0.f;
        if (st != nullptr) st->changed = false;
        return out;
    }
    const float sc = 1.0f / amax;
    const bool warm = (st != nullptr) && st->warm;
    if (warm) {
#pragma unroll
        for (int i = 0; i < 9; ++i) v[i] = st->v[i];
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            const float h0 = h[3 * i] * sc, h1 = h[3 * i + 1] * sc, h2 = h[3 * i + 2] * sc;
#pragma unroll
            for (int c = 0; c < 3; ++c) b[3 * i + c] = fmaf(h2, v[6 + c], fmaf(h1, v[3 + c], h0 * v[c]));
        }
    } else {
#pragma unroll
        for (int i = 0; i < 9; ++i) {
            b[i] = h[i] * sc;
            v[i] = (i % 4 == 0) ? 1.f : 0.f;
        }
    }
    const float kThr2 = 9e-14f;                 // (3e-7)^2
    float thr2 = warm ? 3.6e-13f : kThr2;       // (6e-7)^2 to START rotating from a warm state
    bool any = false;
    for (int sweep = 0; sweep < 12; ++sweep) {
        bool rotated = false;
        float max_t = 0.f;
        jacobi_pair<0, 1>(b, v, thr2, rotated, max_t);
        jacobi_pair<0, 2>(b, v, thr2, rotated, max_t);
        jacobi_pair<1, 2>(b, v, thr2, rotated, max_t);
        if (!rotated) break;
        any = true;
        thr2 = kThr2;
#ifndef ICPF_KABSCH_EXACT_ANGLE
        // cyclic Jacobi converges quadratically: after a sweep whose largest rotation was 1e-4 the columns are orthogonal to
        // 1e-8, below the threshold the next sweep would test against
        if (max_t < 1e-4f) break;
#endif
    }
    // squared column norms = squared singular values; pick the two dominant columns a, b_
    float n0 = fmaf(b[6], b[6], fmaf(b[3], b[3], b[0] * b[0]));
    float n1 = fmaf(b[7], b[7], fmaf(b[4], b[4], b[1] * b[1]));
    float n2 = fmaf(b[8], b[8], fmaf(b[5], b[5], b[2] * b[2]));
    // order the columns by decreasing norm with three register-level compare-swaps (no dynamic indexing)
    if (n1 > n0) { swap_cols<0, 1>(b, v); float f = n0; n0 = n1; n1 = f; any = true; }
    if (n2 > n0) { swap_cols<0, 2>(b, v); float f = n0; n0 = n2; n2 = f; any = true; }
    if (n2 > n1) { swap_cols<1, 2>(b, v); float f = n1; n1 = n2; n2 = f; any = true; }
    float va0 = v[0], va1 = v[3], va2 = v[6];
    float vb0 = v[1], vb1 = v[4], vb2 = v[7];
    if (any || !warm) {
        // re-orthonormalise the right frame (removes the accumulated fp32 drift of the Givens products); a warm
        // frame that needed no rotation is already clean and is used as stored
        normalize3(va0, va1, va2, fmaf(va2, va2, fmaf(va1, va1, va0 * va0)));
        const float e = fmaf(va2, vb2, fmaf(va1, vb1, va0 * vb0));
        vb0 = fmaf(-e, va0, vb0); vb1 = fmaf(-e, va1, vb1); vb2 = fmaf(-e, va2, vb2);
        normalize3(vb0, vb1, vb2, fmaf(vb2, vb2, fmaf(vb1, vb1, vb0 * vb0)));
    }
    // left vectors from fresh products H v (not from the incrementally rotated B): R is then a pure function of
    // (H, stored frame), which is what makes an unchanged H reproduce R bit for bit
    float ua0 = fmaf(h[2] * sc, va2, fmaf(h[1] * sc, va1, h[0] * sc * va0));
    float ua1 = fmaf(h[5] * sc, va2, fmaf(h[4] * sc, va1, h[3] * sc * va0));
    float ua2 = fmaf(h[8] * sc, va2, fmaf(h[7] * sc, va1, h[6] * sc * va0));
    float ub0 = fmaf(h[2] * sc, vb2, fmaf(h[1] * sc, vb1, h[0] * sc * vb0));
    float ub1 = fmaf(h[5] * sc, vb2, fmaf(h[4] * sc, vb1, h[3] * sc * vb0));
    float ub2 = fmaf(h[8] * sc, vb2, fmaf(h[7] * sc, vb1, h[6] * sc * vb0));
    normalize3(ua0, ua1, ua2, fmaf(ua2, ua2, fmaf(ua1, ua1, ua0 * ua0)));
    {
        const float d = fmaf(ua2, ub2, fmaf(ua1, ub1, ua0 * ub0));
        ub0 = fmaf(-d, ua0, ub0); ub1 = fmaf(-d, ua1, ub1); ub2 = fmaf(-d, ua2, ub2);
    }
    float nub = fmaf(ub2, ub2, fmaf(ub1, ub1, ub0 * ub0));
    if (!(nub > 1e-30f)) {
        // rank-1 cross-covariance: the rotation about u_a is not determined by the data (the reference returns
        // whatever LAPACK picks).  Choose the unit vector orthogonal to u_a closest to the image of v_b under the
        // minimal rotation, i.e. v_b made orthogonal to u_a; fall back to a coordinate axis.
        float d = fmaf(ua2, vb2, fmaf(ua1, vb1, ua0 * vb0));
        ub0 = fmaf(-d, ua0, vb0); ub1 = fmaf(-d, ua1, vb1); ub2 = fmaf(-d, ua2, vb2);
        nub = fmaf(ub2, ub2, fmaf(ub1, ub1, ub0 * ub0));
        if (!(nub > 1e-12f)) {
            const float ax = fabsf(ua0), ay = fabsf(ua1), az = fabsf(ua2);
            float e0 = 0.f, e1 = 0.f, e2 = 0.f;
            if (ax <= ay && ax <= az) e0 = 1.f; else if (ay <= az) e1 = 1.f; else e2 = 1.f;
            d = fmaf(ua2, e2, fmaf(ua1, e1, ua0 * e0));
            ub0 = fmaf(-d, ua0, e0); ub1 = fmaf(-d, ua1, e1); ub2 = fmaf(-d, ua2, e2);
            nub = fmaf(ub2, ub2, fmaf(ub1, ub1, ub0 * ub0));
        }
    }
    normalize3(ub0, ub1, ub2, nub);
    const float uc0 = fmaf(ua1, ub2, -ua2 * ub1), uc1 = fmaf(ua2, ub0, -ua0 * ub2), uc2 = fmaf(ua0, ub1, -ua1 * ub0);
    const float vc0 = fmaf(va1, vb2, -va2 * vb1), vc1 = fmaf(va2, vb0, -va0 * vb2), vc2 = fmaf(va0, vb1, -va1 * vb0);
    const float ua[3] = {ua0, ua1, ua2}, ub[3] = {ub0, ub1, ub2}, uc[3] = {uc0, uc1, uc2};
    const float va[3] = {va0, va1, va2}, vb[3] = {vb0, vb1, vb2}, vc[3] = {vc0, vc1, vc2};
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) out.r[3 * i + j] = fmaf(uc[i], vc[j], fmaf(ub[i], vb[j], ua[i] * va[j]));
    if (st != nullptr) {
        bool changed = !warm;
        if (any || !warm) {
            // keep the re-orthonormalised frame for the next solve (only when something changed: see the header
            // comment); `changed` reports whether the stored bits actually differ -- a polish that lands on the same
            // frame again leaves the state, and therefore every later solve of the same H, unchanged
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                changed = changed || (st->v[3 * i] != va[i]) || (st->v[3 * i + 1] != vb[i]) || (st->v[3 * i + 2] != vc[i]);
                st->v[3 * i] = va[i];
                st->v[3 * i + 1] = vb[i];
                st->v[3 * i + 2] = vc[i];
            }
            st->warm = true;
        }
        st->changed = changed;
    }
    return out;
}

}  // namespace icpf
