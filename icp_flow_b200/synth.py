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


def make_scene(num_clusters: int = 200, num_points: int = 150_000, seed: int = 4, *, dynamic_frac: float = 0.15,
               median_size: int = 80, sigma: float = 1.3, max_size: int = 12_000, noise: float = 0.01,
               density: float = 40.0):
    """A "Waymo-shape" frame pair (BASELINE config C4): ``num_clusters`` box-shell clusters with a
    log-normal size distribution (median ~80, p90 ~450, a few of several thousand points) on a 100 m x 100 m lattice,
    the rest of the ``num_points`` budget as ground (label -1e8) and unclustered (-1) points, scan order shuffled.

    Static clusters keep their label in both scans and move by a residual (<= 5 cm, <= 3 deg yaw); ``dynamic_frac`` of
    the clusters move by up to 1.5 m and carry a DIFFERENT label in the dst scan, so that only the all-against-all
    dynamic stage of ``match_pcds`` can pair them.  Returns ``(src_points, src_labels, dst_points, dst_labels, meta)``
    with fp32 points ``[n,3]``, fp32 labels ``[n]`` and ``meta`` = per src label the dst label and the 4x4 motion.
    """
    rng = np.random.default_rng(seed)
    K = int(num_clusters)
    sizes = np.clip(np.exp(rng.normal(np.log(median_size), sigma, size=K)).astype(int), 30, max_size)
    side = int(np.ceil(np.sqrt(K)))
    pitch = 100.0 / side
    dynamic = rng.uniform(size=K) < dynamic_frac
    src_pts, src_lab, dst_pts, dst_lab = [], [], [], []
    dst_label_of = np.arange(K)
    dst_label_of[dynamic] = K + np.arange(int(dynamic.sum()))            # moved objects get a fresh label
    motion = np.tile(np.eye(4), (K, 1, 1))
    for k in range(K):
        n = int(sizes[k])
        scale = np.sqrt(n / 512.0 * 15.0 / density)                     # `density` points / m^2 whatever the size
        extents = np.array([rng.uniform(1.5, 5.0), rng.uniform(0.8, 2.2), rng.uniform(0.8, 2.0)]) * scale
        extents = np.minimum(extents, [pitch - 3.5, pitch - 3.5, 4.0])
        centre = np.array([-50.0 + pitch * (k % side + 0.5), -50.0 + pitch * (k // side + 0.5), rng.uniform(0.5, 2.0)])
        a = _shell_points(rng, extents, n)
        b = _shell_points(rng, extents, max(30, int(n * rng.uniform(0.85, 1.15))))
        ang = np.deg2rad(rng.uniform(-3.0, 3.0))
        t = rng.uniform(-0.05, 0.05, size=3)
        if dynamic[k]:
            t[:2] = rng.uniform(-1.0, 1.0, size=2)
        c, s = np.cos(ang), np.sin(ang)
        Rz = np.array([[c, -s, 0.0], [s, c, 0.0], [0.0, 0.0, 1.0]])
        b = b @ Rz.T + t + rng.normal(0.0, noise, size=b.shape)
        src_pts.append(a + centre)
        dst_pts.append(b + centre)
        src_lab.append(np.full(len(a), k, np.float32))
        dst_lab.append(np.full(len(b), dst_label_of[k], np.float32))
        motion[k, :3, :3] = Rz
        motion[k, :3, 3] = centre + t - Rz @ centre                     # p' = Rz (p - c) + c + t
    scans = []
    for pts, lab in ((src_pts, src_lab), (dst_pts, dst_lab)):
        pts, lab = np.concatenate(pts), np.concatenate(lab)
        rest = max(0, int(num_points) - len(pts))
        n_ground = int(rest * 0.9)
        ground = np.stack([rng.uniform(-55, 55, n_ground), rng.uniform(-55, 55, n_ground),
                           rng.normal(0.0, 0.03, n_ground)], 1)
        loose = np.stack([rng.uniform(-55, 55, rest - n_ground), rng.uniform(-55, 55, rest - n_ground),
                          rng.uniform(0.2, 3.0, rest - n_ground)], 1)
        pts = np.concatenate([pts, ground, loose])
        lab = np.concatenate([lab, np.full(n_ground, -1e8, np.float32), np.full(rest - n_ground, -1.0, np.float32)])
        perm = rng.permutation(len(pts))
        scans += [pts[perm].astype(np.float32), lab[perm].astype(np.float32)]
    meta = {"dst_label": dst_label_of, "motion": motion, "dynamic": dynamic, "sizes": sizes}
    return scans[0], scans[1], scans[2], scans[3], meta
