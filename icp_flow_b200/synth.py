"""Synthetic cluster-pair batches in the reference's padded layout.

Each cluster is a ``[N,4]`` fp32 row block ``(x, y, z, flag)``: valid rows first with flag 1.0, padded rows
``(1e8, 1e8, 1e8, 0.0)`` -- the convention of ``pad_segment`` (/root/reference/utils_helper.py:185-196).

Points are drawn uniformly on the surface of an axis-aligned box (a LiDAR-like shell); ``dst`` is an
independent resampling of the same shell moved by a small yaw about the cluster centre plus a translation,
with Gaussian range noise.  ``wrong_frac`` of the pairs are unrelated boxes (zero-inlier / rollback paths).
numpy only, fully determined by ``seed``.
"""
from __future__ import annotations

import numpy as np

PAD_COORD = np.float32(1e8)


def _shell_points(rng: np.random.Generator, extents: np.ndarray, n: int) -> np.ndarray:
    """n points uniform on the surface of the box [-e/2, e/2]^3 (extents e = (ex, ey, ez))."""
    ex, ey, ez = extents
    areas = np.array([ey * ez, ey * ez, ex * ez, ex * ez, ex * ey, ex * ey], dtype=np.float64)
    face = rng.choice(6, size=n, p=areas / areas.sum())
    u = rng.uniform(-0.5, 0.5, size=(n, 3)) * extents[None, :]
    axis = face // 2
    sign = np.where(face % 2 == 0, -0.5, 0.5)
    u[np.arange(n), axis] = sign * extents[axis]
    return u


def make_pairs(num_pairs: int, max_points: int, seed: int = 1234, *, ragged: bool = False,
               residual_only: bool = False, wrong_frac: float = 0.05, noise: float = 0.01,
               min_points: int | None = None, keep_density: bool = True):
    """Return ``(src, dst, meta)`` with ``src, dst`` float32 ``[P, N, 4]`` and ``meta`` the ground-truth motion.

    residual_only: translation U[-0.05,0.05]^3 m (ICP started from identity has inliers; configs C2/C5);
    otherwise translation (U[-2,2], U[-2,2], U[-0.05,0.05]) m (needs the histogram init; config C3).
    ragged: valid counts n_s, n_d ~ U[N/4, N] independently, else all N rows valid.
    keep_density: box extents are scaled by sqrt(min(N,512)/512) so that small test batches keep the surface density
    of the N=512 benchmark clusters (~15 points/m^2; otherwise few points have a neighbour within thres_dist).
    """
    rng = np.random.default_rng(seed)
    P, N = int(num_pairs), int(max_points)
    src = np.empty((P, N, 4), dtype=np.float32)
    dst = np.empty((P, N, 4), dtype=np.float32)
    src[..., :3] = PAD_COORD
    dst[..., :3] = PAD_COORD
    src[..., 3] = 0.0
    dst[..., 3] = 0.0
    yaw = np.zeros(P)
    trans = np.zeros((P, 3))
    wrong = np.zeros(P, dtype=bool)
    lo = max(4, N // 4) if min_points is None else min_points
    shrink = float(np.sqrt(min(N, 512) / 512.0)) if keep_density else 1.0
    for p in range(P):
        extents = np.array([rng.uniform(1.5, 5.0), rng.uniform(0.8, 2.2), rng.uniform(0.8, 2.0)]) * shrink
        centre = np.array([rng.uniform(-32, 32), rng.uniform(-32, 32), rng.uniform(0, 2)])
        n_s = int(rng.integers(lo, N + 1)) if ragged else N
        n_d = int(rng.integers(lo, N + 1)) if ragged else N
        a = _shell_points(rng, extents, n_s)
        is_wrong = rng.uniform() < wrong_frac
        if is_wrong:
            extents_d = np.array([rng.uniform(1.5, 5.0), rng.uniform(0.8, 2.2), rng.uniform(0.8, 2.0)]) * shrink
            b = _shell_points(rng, extents_d, n_d)
        else:
            b = _shell_points(rng, extents, n_d)
        ang = np.deg2rad(rng.uniform(-3.0, 3.0))
        if residual_only:
            t = rng.uniform(-0.05, 0.05, size=3)
        else:
            t = np.array([rng.uniform(-2, 2), rng.uniform(-2, 2), rng.uniform(-0.05, 0.05)])
        c, s = np.cos(ang), np.sin(ang)
        Rz = np.array([[c, -s, 0.0], [s, c, 0.0], [0.0, 0.0, 1.0]])
        b = b @ Rz.T + t + rng.normal(0.0, noise, size=b.shape)
        src[p, :n_s, :3] = (a + centre).astype(np.float32)
        dst[p, :n_d, :3] = (b + centre).astype(np.float32)
        src[p, :n_s, 3] = 1.0
        dst[p, :n_d, 3] = 1.0
        yaw[p], trans[p], wrong[p] = ang, t, is_wrong
    meta = {"yaw": yaw, "translation": trans, "wrong": wrong}
    return src, dst, meta
