"""oracle/icp_oracle.py -- TEST INFRASTRUCTURE: CPU restatement of ICP-Flow's per-cluster-pair registration path.

This file is the *checker* for the CUDA engine.  It is imported only by ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py``; the
product package ``icp_flow_b200`` never imports it and has no CPU fallback.

Parity status: **the reference ships no test or golden vector for this path** (SURVEY.md section 4), so
the restatement is pinned the other way round -- ``oracle/gen_golden.py`` runs the reference's own
Python files verbatim (``oracle/ref_loader.py``) on seeded inputs in the build container and commits the
outputs under ``tests/golden/``; ``tests/test_oracle_golden.py`` requires this restatement to reproduce
them bit for bit (fp32, torch CPU), and the analytic known answer of ``hist_cuda/test.py`` (arg-max bin
``(50,130,7)``).

Everything is written with torch CPU ops in the same order as the reference so that the fp32 rounding
is identical; the arithmetic-carrying third-party leaves (pytorch3d ``knn_points``, the CUDA-only
``hist`` kernel) are the C restatements in ``oracle/knn_cpu.c``.

Reference map (file:line under /root/reference):
    hist_icp                      utils_match.py:138-157
    estimate_init_pose[_batch]    utils_hist.py:33-124        topk_nms  utils_hist.py:21-29
    hist (vote kernel)            hist_cuda/cpp/hist_cuda_core.cuh:35-62, hist_cuda.cu:59-85
    apply_icp / pytorch3d_icp     utils_icp.py:20-73
    iterative_closest_point       utils_icp_pytorch3d.py:37-225
    corresponding_points_alignment utils_icp_pytorch3d.py:233-382
    _apply_similarity_transform   utils_icp_pytorch3d.py:385-396
    nearest_neighbor_batch        utils_helper.py:20-30
    transform_points_batch        utils_helper.py:76-87
    pad_segment                   utils_helper.py:185-196
    flow_estimation_torch         utils_flow.py:57-69
"""
from __future__ import annotations

import dataclasses
from typing import List, NamedTuple, Optional

import torch

from . import leaves


@dataclasses.dataclass
class PathParams:
    """The values of ``args`` (and the hard-coded constants) that the path reads.

    thres_dist / translation_frame / chunk_size: main.py:45-132 (argparse), translation_frame is rewritten
    per frame pair (main.py:200, demo.py:205).  max_iterations / relative_rmse_thr: utils_icp.py:54-55.
    topk / nms_kernel: utils_hist.py:21.
    """

    thres_dist: float = 0.1
    translation_frame: float = 3.333
    chunk_size: int = 50
    max_iterations: int = 100
    relative_rmse_thr: float = 1e-6
    topk: int = 5
    nms_kernel: int = 11
    # NOT a reference parameter: which of several EQUAL vote counts `torch.topk` returns first is implementation-defined
    # (CPU partial sort vs CUDA radix select, utils_hist.py:27).  "torch" = whatever this torch build does (the pinned
    # goldens); "low" / "high" = lowest / highest flat bin index first -- the other outcomes the reference admits, used by
    # oracle/adjudicate.py only.
    topk_ties: str = "torch"


class IcpTrace(NamedTuple):
    R: torch.Tensor          # [P,3,3] row-vector convention  x' = x R + T
    T: torch.Tensor          # [P,3]
    rmse: torch.Tensor       # [P]
    iterations: int          # batch iterations executed
    converged: bool
    conv_flags: torch.Tensor  # [iterations,P] bool: relative_rmse <= thr at that iteration
    R_hist: List[torch.Tensor]
    T_hist: List[torch.Tensor]
    # conditioning diagnostics (not part of the reference; used by the parity tests to explain discrete flips)
    min_gate_margin: Optional[torch.Tensor] = None   # [iterations,P] min over points of | |x - nn| - thres |  (m)
    min_inliers: Optional[torch.Tensor] = None       # [P] min over iterations of the gated correspondence count
    min_sigma_ratio: Optional[torch.Tensor] = None   # [P] min over iterations of sigma_2 / sigma_1 of the cross-covariance
    coord_scale: Optional[torch.Tensor] = None       # [P] largest |coordinate| of the pair's valid rows (m): sets the fp32 ulp


# ------------------------------------------------------------------------------------------ leaves
def pad_cluster(points: torch.Tensor, max_points: int) -> torch.Tensor:
    """utils_helper.py:185-196 without the random subsampling branch (inputs must have <= max_points rows).

    Row layout (x, y, z, flag); valid rows first with flag 1, padded rows (1e8,1e8,1e8,0).
    """
    n = points.shape[0]
    if n > max_points:
        raise ValueError("subsampling uses the global torch RNG in the reference; compare on padded inputs")
    out = points.new_full((max_points, 4), 1e8)
    out[:n, 0:3] = points[:, 0:3]
    out[:n, 3] = 1.0
    out[n:, 3] = 0.0
    return out


def nearest_neighbor_batch(src: torch.Tensor, dst: torch.Tensor):
    """utils_helper.py:20-30 -- unbounded K=1 NN over ALL rows (padding included), returns sqrt distances."""
    assert src.dim() == 3 and dst.dim() == 3 and len(src) == len(dst)
    assert src.shape[2] >= 3 and dst.shape[2] >= 3
    d2, idx = leaves.knn1(src[:, :, 0:3], dst[:, :, 0:3])
    return idx, d2.to(src.dtype).sqrt()


def transform_points_batch(xyz: torch.Tensor, pose: torch.Tensor) -> torch.Tensor:
    """utils_helper.py:76-87 -- [x y z 1] @ pose^T, flag column carried through."""
    assert xyz.dim() == 3 and pose.dim() == 3 and xyz.shape[2] == 4 and pose.shape[1:] == (4, 4)
    b, n, _ = xyz.shape
    homo = torch.cat([xyz[:, :, 0:3], xyz.new_ones((b, n, 1))], dim=-1)
    moved = torch.bmm(homo, pose.permute(0, 2, 1))
    return torch.cat([moved[:, :, 0:3], xyz[:, :, -1:]], dim=-1)


def bin_edges(p: PathParams):
    """utils_hist.py:60-65 -- bin *starts*; xy cover [-F, F], z covers {-tau, 0, tau}."""
    eps = 1e-8
    f, tau = p.translation_frame, p.thres_dist
    bx = torch.arange(-f, f + tau - eps, tau)
    by = torch.arange(-f, f + tau - eps, tau)
    bz = torch.arange(-tau, tau + tau - eps, tau)
    return bx, by, bz


def vote_histogram(src: torch.Tensor, dst: torch.Tensor, p: PathParams):
    """utils_hist.py:69-72: hist(dst, src, min=bins.min(), max=bins.max(), len=len(bins))."""
    bx, by, bz = bin_edges(p)
    mins = (float(bx.min()), float(by.min()), float(bz.min()))
    maxs = (float(bx.max()), float(by.max()), float(bz.max()))
    lens = (len(bx), len(by), len(bz))
    h = leaves.hist_votes(dst, src, mins, maxs, lens).to(src.dtype)
    return h, (bx, by, bz)


def topk_nms(x: torch.Tensor, k: int = 5, kernel_size: int = 11, ties: str = "torch"):
    """utils_hist.py:21-29 -- keep bins equal to their 11^3 window max, then top-k of the flattened volume.
    `ties` (see PathParams.topk_ties): the order among equal vote counts; "torch" is the reference call itself."""
    b = x.shape[0]
    x5 = x.unsqueeze(1)
    pooled = torch.nn.functional.max_pool3d(x5, kernel_size=kernel_size, stride=1, padding=(kernel_size - 1) // 2)
    keep = (x5 == pooled).float().clamp(min=0.0)
    flat = (x5 * keep).view(b, -1)
    if ties == "torch":
        votes, idxs = torch.topk(flat, dim=1, k=k)
        return votes, idxs.long()
    if ties == "high":
        flat = flat.flip(1)
    votes, idxs = torch.sort(flat, dim=1, descending=True, stable=True)      # stable: equal counts keep index order
    votes, idxs = votes[:, :k], idxs[:, :k]
    if ties == "high":
        idxs = flat.shape[1] - 1 - idxs
    return votes, idxs.long()


def ambiguous_topk_rows(src: torch.Tensor, dst: torch.Tensor, p: PathParams) -> torch.Tensor:
    """Pairs whose top-k peak SET is not determined by the reference: the k-th and (k+1)-th largest surviving vote
    counts are equal, so which of the tied peaks ``torch.topk`` keeps is implementation-defined (CPU partial sort vs
    CUDA radix select give different answers).  Diagnostics for the parity tests, not part of the reference."""
    hist, _ = vote_histogram(src, dst, p)
    b = hist.shape[0]
    x5 = hist.unsqueeze(1)
    pooled = torch.nn.functional.max_pool3d(x5, kernel_size=p.nms_kernel, stride=1, padding=(p.nms_kernel - 1) // 2)
    kept = (x5 * (x5 == pooled).float()).view(b, -1)
    top = torch.topk(kept, k=p.topk + 1, dim=1)[0]
    return (top[:, p.topk] > 0) & (top[:, p.topk - 1] == top[:, p.topk])


# ------------------------------------------------------------------------------------ histogram init
def estimate_init_pose(src: torch.Tensor, dst: torch.Tensor, p: PathParams, return_debug: bool = False):
    """utils_hist.py:33-44 -- chunked driver around the per-chunk estimate."""
    assert len(src) == len(dst)
    outs, dbg = [], []
    for lo in range(0, len(src), p.chunk_size):
        t, d = _estimate_init_pose_chunk(src[lo:lo + p.chunk_size], dst[lo:lo + p.chunk_size], p)
        outs.append(t)
        dbg.append(d)
    poses = torch.vstack(outs)
    if return_debug:
        keys = dbg[0].keys()
        return poses, {k: torch.cat([d[k] for d in dbg], dim=0) for k in keys}
    return poses


def _estimate_init_pose_chunk(src, dst, p: PathParams):
    """utils_hist.py:46-124."""
    xyz_s, xyz_d = src[:, :, 0:3], dst[:, :, 0:3]
    valid_s, valid_d = src[:, :, -1] > 0.0, dst[:, :, -1] > 0.0
    hist, (bx, by, bz) = vote_histogram(src, dst, p)
    b, h, w, d = hist.shape
    votes, flat = topk_nms(hist, k=p.topk, kernel_size=p.nms_kernel, ties=p.topk_ties)
    cand = torch.stack([bx[flat // d // w % h], by[flat // d % w], bz[flat % d]], dim=-1) + p.thres_dist // 2
    n = xyz_s.shape[1]
    cand = torch.cat([cand, cand.new_zeros(b, 1, 3)], dim=1)         # + the zero translation, last
    k = cand.shape[1]
    moved = xyz_s[:, None, :, :] + cand[:, :, None, :]
    fixed = xyz_d[:, None, :, :].expand(-1, k, -1, -1)
    _, e_fwd = nearest_neighbor_batch(moved.reshape(b * k, n, 3), fixed.reshape(b * k, n, 3))
    _, e_bwd = nearest_neighbor_batch(fixed.reshape(b * k, n, 3), moved.reshape(b * k, n, 3))
    e_fwd = (e_fwd.view(b, k, n) * valid_s[:, None, :]).sum(dim=-1) / valid_s[:, None, :].sum(dim=-1)
    e_bwd = (e_bwd.view(b, k, n) * valid_d[:, None, :]).sum(dim=-1) / valid_d[:, None, :].sum(dim=-1)
    score = torch.minimum(e_fwd, e_bwd)
    best, which = score.min(dim=-1)
    t_best = cand[torch.arange(0, b), which, :]
    pose = torch.eye(4)[None].repeat(b, 1, 1)
    pose[:, 0:3, -1] = t_best
    debug = {"votes": votes, "flat_idx": flat, "candidates": cand, "scores": score, "which": which}
    return pose, debug


# ------------------------------------------------------------------------------------------- Kabsch
def kabsch_weighted(X: torch.Tensor, Y: torch.Tensor, w: torch.Tensor, eps: float = 1e-9):
    """utils_icp_pytorch3d.py:303-377 in the configuration the path uses
    (bool weights, estimate_scale=False, allow_reflection=False).  Row-vector convention: Y ~ X R + T."""
    b = X.shape[0]
    wsum = w[..., None].sum(dim=-2, keepdim=True).clamp(eps)
    mu_x = (X * w[..., None]).sum(dim=-2, keepdim=True) / wsum
    mu_y = (Y * w[..., None]).sum(dim=-2, keepdim=True) / wsum
    Xc = X - mu_x
    Yc = Y - mu_y
    Xc *= w[:, :, None]
    Yc *= w[:, :, None]
    total = torch.clamp(w.sum(1), eps)
    H = torch.bmm(Xc.transpose(2, 1), Yc)
    H = H / total[:, None, None]
    U, S, V = torch.svd(H)
    E = torch.eye(3, dtype=H.dtype)[None].repeat(b, 1, 1)
    E[:, -1, -1] = torch.det(torch.bmm(U, V.transpose(2, 1)))
    R = torch.bmm(torch.bmm(U, E), V.transpose(2, 1))
    T = mu_y[:, 0, :] - torch.bmm(mu_x, R)[:, 0, :]
    return R, T, S


# ---------------------------------------------------------------------------------------------- ICP
def icp_loop(X: torch.Tensor, Y: torch.Tensor, thres: float = 0.1, max_iterations: int = 100,
             relative_rmse_thr: float = 1e-6, keep_history: bool = False, diagnostics: bool = False) -> IcpTrace:
    """utils_icp_pytorch3d.py:100-225 with init_transform=None.

    Absolute transform re-estimated from the *initial* cloud every iteration; NN on the current cloud among
    the first ``len_d`` rows of Y for the first ``len_s`` rows of X; gate d^2 <= thres^2; batch-coupled stop.
    """
    Xt = X[:, :, 0:3]
    Yt = Y[:, :, 0:3]
    b = Xt.shape[0]
    if Xt.shape[0] != Yt.shape[0]:
        raise ValueError("Point sets X and Y have to have the same number of batches and data dimensions.")
    valid_x0 = X[:, :, -1] > 0.0
    len_x = valid_x0.sum(dim=-1)
    len_y = (Y[:, :, -1] > 0.0).sum(dim=-1)
    X0 = Xt.clone()
    R = torch.eye(3, dtype=Xt.dtype)[None].repeat(b, 1, 1)
    T = Xt.new_zeros((b, 3))
    prev = None
    rmse = None
    converged = False
    flags, Rh, Th = [], [], []
    margin = []
    inliers = torch.full((b,), 2 ** 31 - 1, dtype=torch.int64)
    sig = torch.full((b,), float("inf"), dtype=torch.float64)
    in_len = torch.arange(Xt.shape[1])[None, :] < len_x[:, None]
    it = -1
    for it in range(max_iterations):
        d2, idx = leaves.knn1(Xt, Yt, len_x, len_y)
        d2 = d2.to(Xt.dtype)
        nn_pts = torch.gather(Yt, 1, idx[:, :, None].expand(-1, -1, 3))
        m = torch.logical_and(valid_x0, d2 <= thres ** 2)
        R, T, S = kabsch_weighted(X0 * m[:, :, None], nn_pts * m[:, :, None], m)
        if diagnostics:
            gap = (d2.double().sqrt() - thres).abs().masked_fill(~(in_len & valid_x0), float("inf"))
            margin.append(gap.amin(dim=1))
            inliers = torch.minimum(inliers, m.sum(dim=1))
            sig = torch.minimum(sig, (S[:, 1] / S[:, 0].clamp(min=1e-30)).double())
        Xt = torch.bmm(X0, R) + T[:, None, :]          # s == 1: ones * bmm is exact
        sq = ((Xt - nn_pts) ** 2).sum(2)
        rmse = ((sq[:, :, None] * m[..., None]).sum(dim=-2, keepdim=True)
                / m[..., None].sum(dim=-2, keepdim=True).clamp(1e-9)).sqrt()[:, 0, 0]
        rel = rmse.new_ones(b) if prev is None else (prev - rmse) / prev
        ok = rel <= relative_rmse_thr
        flags.append(ok)
        if keep_history:
            Rh.append(R.clone())
            Th.append(T.clone())
        if ok.all():
            converged = True
            break
        prev = rmse
    return IcpTrace(R, T, rmse, it + 1, converged, torch.stack(flags) if flags else torch.zeros(0, b, dtype=torch.bool),
                    Rh, Th, torch.stack(margin) if diagnostics else None, inliers if diagnostics else None,
                    sig if diagnostics else None, _coord_scale(X, Y) if diagnostics else None)


def _coord_scale(X: torch.Tensor, Y: torch.Tensor) -> torch.Tensor:
    sx = (X[:, :, 0:3].abs() * (X[:, :, 3:4] > 0)).amax(dim=(1, 2))
    sy = (Y[:, :, 0:3].abs() * (Y[:, :, 3:4] > 0)).amax(dim=(1, 2))
    return torch.maximum(sx, sy).double()


def unstable_pairs(trace: IcpTrace, margin_m: float = 3e-5, min_inliers: int = 6, sigma_ratio: float = 1e-3,
                   last: int = 3, early_margin_m: float = 1e-6, early_ulps: float = 0.0):
    """Pairs whose reference result is not numerically determined to 1e-4: during the last `last` iterations a
    correspondence sat within `margin_m` of the gate (a differently-rounded but equally valid fp32 evaluation of
    x R + T -- 1 ulp at 50 m is 4e-6 m -- flips it and moves the fixed point), or at ANY iteration one sat within
    `early_margin_m` of it (which side it falls on is decided by the last bits of R, T and the order of the three
    products, and a registration need not find its way back to the same fixed point within the iterations the batch
    grants it), or the Kabsch system was (nearly) rank deficient (fewer than `min_inliers` correspondences / second
    singular value below `sigma_ratio` of the first: the reference returns whatever LAPACK picks, SURVEY.md section 7
    "3x3 SVD").

    `early_ulps` > 0 widens the any-iteration margin to that many fp32 ulps of the pair's largest coordinate (2.5 ulps at
    50 m are 1e-5 m).  Most flips that early are harmless -- the registration finds the same fixed point anyway -- so
    the wide criterion marks 15-40 % of ordinary pairs and is NOT a reason to skip a pair: tests/test_random_parity.py
    uses it to decide which of the rare pairs outside the tolerance may be held to the quality of the registration
    instead (the engine's residual must not exceed the reference's)."""
    assert trace.min_gate_margin is not None, "run icp_loop(..., diagnostics=True)"
    tail = trace.min_gate_margin[-last:].amin(dim=0)
    anywhere = trace.min_gate_margin.amin(dim=0)
    early = torch.full_like(anywhere, early_margin_m)
    if early_ulps > 0 and trace.coord_scale is not None:
        ulp = torch.exp2(torch.floor(torch.log2(trace.coord_scale.clamp(min=1e-3))) - 23.0)     # fp32 spacing at that range
        early = torch.maximum(early, early_ulps * ulp.to(anywhere.dtype))
    return ((tail < margin_m) | (anywhere < early) | (trace.min_inliers < min_inliers) |
            (trace.min_sigma_ratio < sigma_ratio))


def pack_rt(R: torch.Tensor, T: torch.Tensor) -> torch.Tensor:
    """utils_icp.py:60-65 -- row-convention (R,T) to column-convention 4x4 [[R^T, T],[0,1]]."""
    top = torch.cat([R, T[:, None, :]], dim=1)
    M = torch.cat([top.permute(0, 2, 1), top.new_zeros(len(T), 1, 4)], dim=1)
    M[:, 3, 3] = 1.0
    return M


def apply_icp(src: torch.Tensor, dst: torch.Tensor, init_poses: torch.Tensor, p: PathParams,
              return_debug: bool = False):
    """utils_icp.py:20-48 -- ICP from the initialised cloud, compose, roll back where the mean NN error did not drop."""
    moved = transform_points_batch(src, init_poses)
    trace = icp_loop(moved, dst, thres=p.thres_dist, max_iterations=p.max_iterations,
                     relative_rmse_thr=p.relative_rmse_thr)
    poses = torch.bmm(pack_rt(trace.R, trace.T), init_poses)
    valid_s = src[:, :, -1] > 0.0
    _, e0 = nearest_neighbor_batch(moved, dst)
    e0 = (e0 * valid_s).sum(dim=1) / valid_s.sum(dim=1)
    _, e1 = nearest_neighbor_batch(transform_points_batch(src, poses), dst)
    e1 = (e1 * valid_s).sum(dim=1) / valid_s.sum(dim=1)
    worse = e1 >= e0
    poses[worse] = init_poses[worse]
    if return_debug:
        return poses, {"error_init": e0, "error_icp": e1, "rolled_back": worse,
                       "iterations": torch.tensor(trace.iterations), "icp_R": trace.R, "icp_T": trace.T}
    return poses


def hist_icp(src: torch.Tensor, dst: torch.Tensor, p: PathParams, return_debug: bool = False):
    """utils_match.py:138-157 -- always register the smaller cloud onto the larger one, undo the swap by inversion."""
    n_s = (src[:, :, -1] > 0.0).sum(dim=1)
    n_d = (dst[:, :, -1] > 0.0).sum(dim=1)
    swap = n_s > n_d
    a, c = src.clone(), dst.clone()
    a[swap] = dst[swap]
    c[swap] = src[swap]
    with torch.no_grad():
        init = estimate_init_pose(a, c, p)
        poses_, dbg = apply_icp(a, c, init, p, return_debug=True)
    if int(swap.sum()) > 0:
        poses = poses_.clone()
        poses[swap] = torch.linalg.inv(poses_[swap])
    else:
        poses = poses_
    if return_debug:
        dbg = dict(dbg)
        dbg.update({"init": init, "swapped": swap, "poses_before_unswap": poses_})
        return poses, dbg
    return poses


# --------------------------------------------------------------------------------------------- flow
def flow_from_transforms(src_points: torch.Tensor, src_labels: torch.Tensor, pair_src_labels: torch.Tensor,
                         transforms: torch.Tensor, pose: Optional[torch.Tensor] = None) -> torch.Tensor:
    """utils_flow.py:57-69 -- per-point rigid flow  T_cluster(label) * pose * p - p  (identity for unmatched labels)."""
    n = len(src_points)
    pose = torch.eye(4) if pose is None else pose
    per_point = torch.eye(4)[None].repeat(n, 1, 1)
    hit_pt, hit_pair = torch.nonzero((src_labels[:, None] - pair_src_labels[None, :]) == 0, as_tuple=True)
    per_point[hit_pt] = transforms[hit_pair]
    per_point = torch.bmm(per_point, pose[None].expand(n, 4, 4))
    homo = torch.cat([src_points, src_points.new_ones(n, 1)], dim=-1)
    return torch.bmm(per_point, homo[:, :, None])[:, 0:3, 0] - src_points


# ------------------------------------------------------------------------------------- match_eval (row f1)
def euler_zyx_degrees(R: torch.Tensor) -> torch.Tensor:
    """pytorch3d 0.7.4 ``matrix_to_euler_angles(R, "ZYX")`` in degrees, restated in closed form
    (call site utils_match.py:184):  (atan2(m10, m00), asin(-m20), atan2(m21, m22))."""
    z = torch.atan2(R[..., 1, 0], R[..., 0, 0])
    y = torch.asin(-R[..., 2, 0])
    x = torch.atan2(R[..., 2, 1], R[..., 2, 2])
    return torch.stack([z, y, x], dim=-1) * 180.0 / 3.141592653589793


def match_eval(pcd1: torch.Tensor, pcd2: torch.Tensor, transformations: torch.Tensor, p: PathParams):
    """utils_match.py:159-213 -- quality metrics of a batch of registrations: mean NN errors in both directions,
    inlier counts (strict ``< thres_dist``), inlier ratios, IoUs, mean translation of the moved cloud, ZYX Euler angles."""
    moved = transform_points_batch(pcd1, transformations)
    m1 = pcd1[:, :, -1] > 0.0
    m2 = pcd2[:, :, -1] > 0.0
    _, e12 = nearest_neighbor_batch(moved, pcd2)
    _, e21 = nearest_neighbor_batch(pcd2, moved)
    in1 = torch.logical_and(e12 < p.thres_dist, m1).float()
    in2 = torch.logical_and(e21 < p.thres_dist, m2).float()
    r1 = in1.sum(dim=1) / m1.sum(dim=1)
    r2 = in2.sum(dim=1) / m2.sum(dim=1)
    iou1 = in1.sum(dim=1) / (m1.sum(dim=1) + m2.sum(dim=1) - in2.sum(dim=1))
    iou2 = in2.sum(dim=1) / (m1.sum(dim=1) + m2.sum(dim=1) - in1.sum(dim=1))
    err1 = (e12 * m1).sum(1) / m1.sum(1)
    err2 = (e21 * m2).sum(1) / m2.sum(1)
    mean_moved = (moved[:, :, 0:3] * m1[:, :, None]).sum(dim=1) / m1.sum(dim=1, keepdim=True)
    mean_orig = (pcd1[:, :, 0:3] * m1[:, :, None]).sum(dim=1) / m1.sum(dim=1, keepdim=True)
    return (torch.stack([err1, err2], dim=1), torch.stack([in1.sum(1), in2.sum(1)], dim=1),
            torch.stack([r1, r2], dim=1), torch.stack([iou1, iou2], dim=1), mean_moved - mean_orig,
            euler_zyx_degrees(transformations[:, 0:3, 0:3]))


@dataclasses.dataclass
class MatchGates:
    """The association thresholds ``match_pairs`` reads from ``args`` (main.py:97-110 defaults; demo.sh overrides
    thres_error / thres_iou to 0.2)."""

    max_points: int = 10000
    thres_error: float = 0.1
    thres_iou: float = 0.1
    thres_rot: float = 0.1
    min_cluster_size: int = 30      # main.py:79 (demo.sh: 20)
    thres_box: float = 0.1          # main.py:101


def check_transformation(translation: torch.Tensor, rotation: torch.Tensor, iou: torch.Tensor, p: PathParams,
                         g: MatchGates) -> bool:
    """utils_check.py:51-66 -- reject a registration that moves further than the frame budget, overlaps too little, or
    pitches / rolls more than ``thres_rot * 90`` degrees."""
    if torch.linalg.norm(translation) > p.translation_frame:
        return False
    if iou < g.thres_iou:
        return False
    if torch.abs(rotation[1:3]).max() > g.thres_rot * 90.0:
        return False
    return True


def match_select(pairs, src_labels_unq, dst_labels_unq, evals, transformations, p: PathParams, g: MatchGates):
    """utils_match.py:70-75,94-135 -- scatter the accepted registrations into [n_src, n_dst] matrices, keep for every
    src cluster the dst cluster of least ``min(error)`` (``match_segments_descend``, utils_helper.py:108-115) if that
    error is below ``thres_error``.  Returns (rows [K,10], transformations [K,4,4])."""
    errors, inliers, ratios, ious, translations, rotations = evals
    ns, nd = len(src_labels_unq), len(dst_labels_unq)
    m_err = torch.zeros((ns, nd, 2)) + 1e8
    m_inl, m_rat, m_iou = torch.zeros((ns, nd, 2)), torch.zeros((ns, nd, 2)), torch.zeros((ns, nd, 2))
    m_T = torch.zeros((ns, nd, 4, 4))
    matches = 0
    for k in range(len(pairs)):
        if not check_transformation(translations[k], rotations[k], min(ious[k]), p, g):
            continue
        si = torch.nonzero(src_labels_unq == pairs[k][0])
        di = torch.nonzero(dst_labels_unq == pairs[k][1])
        m_err[si, di, :], m_inl[si, di, :], m_rat[si, di, :], m_iou[si, di, :] = errors[k], inliers[k], ratios[k], ious[k]
        m_T[si, di] = transformations[k]
        matches += 1
    if matches == 0:
        return torch.zeros(0, 10), torch.zeros(0, 4, 4)
    e_min = m_err.min(-1)[0]
    si = torch.arange(0, ns)
    di = torch.argmin(e_min, dim=1)
    ok = e_min[si, di] < g.thres_error
    si, di = si[ok], di[ok]
    rows = torch.cat([src_labels_unq[si][:, None], dst_labels_unq[di][:, None], m_err[si, di], m_inl[si, di],
                      m_rat[si, di], m_iou[si, di]], dim=1)
    return rows, m_T[si, di]


def match_pairs(src_points, dst_points, src_labels, dst_labels, pairs, p: PathParams, g: MatchGates,
                return_debug: bool = False):
    """utils_match.py:69-135 for clusters of at most ``max_points`` rows (``pad_segment`` subsamples larger ones with the
    global torch RNG, which is outside what a restatement can pin)."""
    assert len(pairs) > 0
    segs_src = torch.stack([pad_cluster(src_points[src_labels == pr[0], 0:3], g.max_points) for pr in pairs])
    segs_dst = torch.stack([pad_cluster(dst_points[dst_labels == pr[1], 0:3], g.max_points) for pr in pairs])
    T = hist_icp(segs_src, segs_dst, p)
    evals = match_eval(segs_src, segs_dst, T, p)
    rows, T_sel = match_select(pairs, torch.unique(src_labels), torch.unique(dst_labels), evals, T, p, g)
    if return_debug:
        return rows, T_sel, {"segs_src": segs_src, "segs_dst": segs_dst, "T": T, "evals": evals}
    return rows, T_sel


# ------------------------------------------------------------------------------- batch construction (row f2)
def cluster_bbox(points: torch.Tensor):
    """utils_helper.py:166-170 (get_bbox_tensor): the three axis extents |max - min|, sorted ascending."""
    x = torch.abs(points[:, 0].max() - points[:, 0].min())
    y = torch.abs(points[:, 1].max() - points[:, 1].min())
    z = torch.abs(points[:, 2].max() - points[:, 2].min())
    return sorted([x, y, z])


def sanity_check(src_points, dst_points, src_labels, dst_labels, pairs, p: PathParams, g: MatchGates):
    """utils_check.py:21-49 -- keep a candidate pair when both clusters have >= min_cluster_size points, no label is
    negative, the xy centroid offset is within translation_frame and each sorted bbox extent pair satisfies
    min >= thres_box * max.  Returns the kept rows of ``pairs`` (``[K,2]``, ``zeros((0,2))`` when none)."""
    keep = []
    for pair in pairs:
        src = src_points[src_labels == pair[0]]
        dst = dst_points[dst_labels == pair[1]]
        if min(len(src), len(dst)) < g.min_cluster_size:
            continue
        if min(pair[0], pair[1]) < 0:
            continue
        if torch.linalg.norm((dst.mean(0) - src.mean(0))[0:2]) > p.translation_frame:
            continue
        sb, db = cluster_bbox(src), cluster_bbox(dst)
        if any(min(sb[k], db[k]) < g.thres_box * max(sb[k], db[k]) for k in range(3)):
            continue
        keep.append(pair)
    return torch.vstack(keep) if len(keep) > 0 else torch.zeros((0, 2))


def pad_cluster_sampled(points: torch.Tensor, max_points: int) -> torch.Tensor:
    """utils_helper.py:185-201 (pad_segment + random_choice) including the subsampling branch: a cluster with more than
    ``max_points`` rows keeps ``torch.randperm(len)[:max_points]`` -- the global torch RNG, so a comparison has to seed
    it and make the calls in the reference's order (pair by pair, src before dst, utils_match.py:84-88)."""
    if len(points) > max_points:
        points = points[torch.randperm(len(points))[0:max_points], :]
    return pad_cluster(points, max_points)


def match_pairs_sampled(src_points, dst_points, src_labels, dst_labels, pairs, p: PathParams, g: MatchGates,
                        return_debug: bool = False):
    """``match_pairs`` (utils_match.py:69-135) with the reference's subsampling of oversized clusters."""
    assert len(pairs) > 0
    segs_src, segs_dst = [], []
    for pr in pairs:                                    # one loop: the RNG is consumed src, dst, src, dst, ...
        segs_src.append(pad_cluster_sampled(src_points[src_labels == pr[0], 0:3], g.max_points))
        segs_dst.append(pad_cluster_sampled(dst_points[dst_labels == pr[1], 0:3], g.max_points))
    segs_src, segs_dst = torch.stack(segs_src), torch.stack(segs_dst)
    T = hist_icp(segs_src, segs_dst, p)
    evals = match_eval(segs_src, segs_dst, T, p)
    rows, T_sel = match_select(pairs, torch.unique(src_labels), torch.unique(dst_labels), evals, T, p, g)
    if return_debug:
        return rows, T_sel, {"segs_src": segs_src, "segs_dst": segs_dst, "T": T, "evals": evals, "pairs": pairs}
    return rows, T_sel


def setdiff1d(t1: torch.Tensor, t2: torch.Tensor) -> torch.Tensor:
    """utils_helper.py:172-183: elements of t1 that are not in t2 (t2 assumed a subset of t1)."""
    t12, counts = torch.cat([torch.unique(t1), torch.unique(t2)]).unique(return_counts=True)
    return t12[torch.where(counts.eq(1))]


def match_pcds(src_points, dst_points, src_labels, dst_labels, p: PathParams, g: MatchGates, return_debug: bool = False):
    """utils_match.py:26-66 -- static stage (label l against label l), then every unmatched src cluster against every
    unmatched dst cluster; both filtered by ``sanity_check`` and resolved by ``match_pairs``."""
    src_unq = torch.unique(src_labels).long()
    dst_unq = torch.unique(dst_labels).long()
    labels_unq = torch.unique(torch.cat([src_unq, dst_unq], dim=0))
    dbg = {}
    pairs = torch.stack([labels_unq, labels_unq], dim=1)
    pairs = pairs[pairs.min(dim=1)[0] >= 0]
    pairs_true = sanity_check(src_points, dst_points, src_labels, dst_labels, pairs, p, g)
    dbg["static_candidates"], dbg["static_kept"] = pairs, pairs_true
    if len(pairs_true) > 0:
        rows_sta, T_sta, dbg["static"] = match_pairs_sampled(src_points, dst_points, src_labels, dst_labels, pairs_true,
                                                             p, g, return_debug=True)
    else:
        rows_sta, T_sta = torch.zeros(0, 10), torch.zeros(0, 4, 4)
    if len(rows_sta) < len(labels_unq):
        if len(rows_sta) > 0:
            src_unq = setdiff1d(src_unq, rows_sta[:, 0])
            dst_unq = setdiff1d(dst_unq, rows_sta[:, 1])
        pairs = torch.stack([src_unq.repeat_interleave(len(dst_unq)), dst_unq.repeat(len(src_unq))], dim=1)
        pairs_true = sanity_check(src_points, dst_points, src_labels, dst_labels, pairs, p, g)
    else:
        pairs, pairs_true = torch.zeros(0, 2), torch.zeros(0, 2)
    dbg["dynamic_candidates"], dbg["dynamic_kept"] = pairs, pairs_true
    if len(pairs_true) > 0:
        rows_dyn, T_dyn, dbg["dynamic"] = match_pairs_sampled(src_points, dst_points, src_labels, dst_labels,
                                                              pairs_true, p, g, return_debug=True)
    else:
        rows_dyn, T_dyn = torch.zeros(0, 10), torch.zeros(0, 4, 4)
    rows, T = torch.cat([rows_sta, rows_dyn], dim=0), torch.cat([T_sta, T_dyn], dim=0)
    if return_debug:
        return rows, T, dbg
    return rows, T


def undetermined_pairs(src: torch.Tensor, dst: torch.Tensor, p: PathParams, early_ulps: float = 0.0) -> torch.Tensor:
    """Pairs whose ``hist_icp`` result the reference does not determine numerically -- a tied top-k peak set, an ICP
    run that passes a discrete flip (``unstable_pairs``) or a roll-back decision inside fp32 noise.  Diagnostics for
    the parity tests (the same three exclusions tests/test_gpu_path.py applies), not part of the reference."""
    n_s = (src[:, :, -1] > 0.0).sum(dim=1)
    n_d = (dst[:, :, -1] > 0.0).sum(dim=1)
    swap = n_s > n_d
    a, c = src.clone(), dst.clone()
    a[swap] = dst[swap]
    c[swap] = src[swap]
    amb = ambiguous_topk_rows(a, c, p)
    init = estimate_init_pose(a, c, p)
    trace = icp_loop(transform_points_batch(a, init), c, p.thres_dist, p.max_iterations, p.relative_rmse_thr,
                     diagnostics=True)
    _, dbg = apply_icp(a, c, init, p, return_debug=True)
    e0, e1 = dbg["error_init"], dbg["error_icp"]
    tie = (e1 - e0).abs() <= 1e-5 * e0.clamp(min=1e-6)
    return amb | unstable_pairs(trace, early_ulps=early_ulps) | tie
