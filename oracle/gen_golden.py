"""oracle/gen_golden.py -- TEST INFRASTRUCTURE.  Regenerates ``tests/golden/*.npz`` from the REAL reference.

Run in the build container only (needs /root/reference):

    python -m oracle.gen_golden            # writes tests/golden/*.npz

Every output array below is produced by calling the reference's own, unmodified functions
(``utils_match.hist_icp``, ``utils_hist.estimate_init_pose``, ``utils_icp.apply_icp``,
``utils_icp_pytorch3d.iterative_closest_point``, ``utils_helper.nearest_neighbor_batch``,
``utils_match.match_eval``, ``utils_match.match_pairs``, ``utils_flow.flow_estimation_torch``) on torch CPU fp32 through ``oracle/ref_loader.py``; the inputs are
stored next to them so the fixtures are self-contained on the GPU box.

Fixtures
    hist_test_vector.npz  inputs of hist_cuda/test.py:19-50 (torch.manual_seed(2022)) + histogram arg-max
                          (analytic known answer (50,130,7)) and peak / total vote counts
    c1_demo.npz           BASELINE config C1: demo.npz -> sklearn DBSCAN(eps .25, min 20) stand-in labels ->
                          first 32 static pairs with both clouds <= 256 points, N=256, F=2.0 (demo.sh)
    synth_hist.npz        24 ragged synthetic pairs, N=128, full hist_icp with F=3.333 (argparse default)
    synth_match_dyn.npz   6 x 6 dynamic-stage candidate pairs through match_pairs (hist_icp + match_eval + gates + selection)
    frame_demo.npz        a whole frame pair through match_pcds (sanity_check, gather + pad_segment incl. the randperm
                          subsampling, both stages) and flow_estimation_torch: the demo.npz scene with the DBSCAN
                          stand-in labels, its largest cluster relabelled as ground (-1e8) and thinned, max_points=512
    synth_icp20.npz       32 synthetic residual-only pairs, N=256: ICP only, 20 forced iterations
                          (relative_rmse_thr=-1) and the reference stopping rule (1e-6, max 100)
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import ref_loader  # noqa: E402
from icp_flow_b200 import synth  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")


def _args(**kw):
    base = dict(thres_dist=0.1, translation_frame=3.333, chunk_size=50, max_points=256, min_cluster_size=20,
                thres_box=0.1, thres_error=0.2, thres_iou=0.2, thres_rot=0.1)       # gates as in demo.sh
    base.update(kw)
    return types.SimpleNamespace(**base)


def _np(t):
    return t.detach().cpu().numpy()


def _run_path(ref, args, src, dst):
    """All reference outputs for one padded batch (torch CPU fp32)."""
    src_t, dst_t = torch.from_numpy(src), torch.from_numpy(dst)
    out = {}
    T = ref.utils_match.hist_icp(args, src_t, dst_t)
    out["T_hist_icp"] = _np(T)
    # inner seams on the swapped copies, exactly as hist_icp builds them (utils_match.py:139-146)
    n_s = (src_t[:, :, -1] > 0).sum(1)
    n_d = (dst_t[:, :, -1] > 0).sum(1)
    swap = n_s > n_d
    a, c = src_t.clone(), dst_t.clone()
    a[swap] = dst_t[swap]
    c[swap] = src_t[swap]
    init = ref.utils_hist.estimate_init_pose(args, a, c)
    out["swapped"] = _np(swap)
    out["init_pose"] = _np(init)
    out["T_apply_icp"] = _np(ref.utils_icp.apply_icp(args, a, c, init))
    moved = ref.utils_helper.transform_points_batch(a, init)
    sol = ref.utils_icp_pytorch3d.iterative_closest_point(moved, c, thres=args.thres_dist, max_iterations=100,
                                                          relative_rmse_thr=1e-6)
    out["icp_R"], out["icp_T"], out["icp_rmse"] = _np(sol.RTs.R), _np(sol.RTs.T), _np(sol.rmse)
    out["icp_iterations"] = np.int64(len(sol.t_history))
    out["icp_converged"] = np.bool_(sol.converged)
    idx, dist = ref.utils_helper.nearest_neighbor_batch(a, c)
    out["nn_idx"], out["nn_dist"] = _np(idx), _np(dist)
    # match_eval on the ORIGINAL (unswapped) clouds with the final transforms, as match_pairs calls it (utils_match.py:93)
    ev = ref.utils_match.match_eval(args, src_t, dst_t, T)
    for name, v in zip(("errors", "inliers", "ratios", "ious", "translations", "rotations"), ev):
        out["eval_" + name] = _np(v)
    return out


def clouds_from_batches(src_b, dst_b, pairs):
    """Scan-level inputs of match_pairs (points + per-point labels) rebuilt from padded pair batches: the valid rows of
    the FIRST batch entry that carries a label, in order -- so that ``points[labels == l]`` is that entry's cluster."""
    out = []
    for batch, col in ((src_b, 0), (dst_b, 1)):
        pts, lab, seen = [], [], set()
        for k, pr in enumerate(pairs):
            l = float(pr[col])
            if l in seen:
                continue
            seen.add(l)
            rows = batch[k][batch[k][:, 3] > 0, 0:3]
            pts.append(rows)
            lab.append(np.full(len(rows), l, np.float32))
        out += [np.concatenate(pts).astype(np.float32), np.concatenate(lab)]
    return out


def _run_match_pairs(ref, args, src_pts, src_lab, dst_pts, dst_lab, pairs):
    rows, T = ref.utils_match.match_pairs(args, torch.from_numpy(src_pts), torch.from_numpy(dst_pts),
                                          torch.from_numpy(src_lab), torch.from_numpy(dst_lab), torch.from_numpy(pairs))
    return {"mp_rows": _np(rows), "mp_T": _np(T), "thres_error": np.float64(args.thres_error),
            "thres_iou": np.float64(args.thres_iou), "thres_rot": np.float64(args.thres_rot),
            "max_points": np.int64(args.max_points)}


def gen_match_dyn(ref):
    """Dynamic-stage shape of match_pairs (utils_match.py:46-57): every src cluster against every dst cluster.  Six
    clusters re-centred onto a 1.2 m lattice so that wrong pairings are inside the histogram range and compete."""
    n = 6
    src, dst, meta = synth.make_pairs(n, 128, seed=99, ragged=True, residual_only=False)
    for k in range(n):
        vs, vd = src[k, :, 3] > 0, dst[k, :, 3] > 0
        off = np.array([10.0 + 1.2 * k, -5.0 + 0.6 * (k % 2), 0.5], np.float32) - src[k, vs, :3].mean(0)
        src[k, vs, :3] += off
        dst[k, vd, :3] += off
    ident = np.stack([np.arange(n), np.arange(n) + 100], 1).astype(np.float32)
    src_pts, src_lab, dst_pts, dst_lab = clouds_from_batches(src, dst, ident)
    pairs = np.stack([np.repeat(np.arange(n), n), np.tile(np.arange(n) + 100, n)], 1).astype(np.float32)
    args = _args(translation_frame=3.333, max_points=128)
    out = _run_match_pairs(ref, args, src_pts, src_lab, dst_pts, dst_lab, pairs)
    np.savez_compressed(os.path.join(GOLDEN, "synth_match_dyn.npz"), src_points=src_pts, src_labels=src_lab,
                        dst_points=dst_pts, dst_labels=dst_lab, pairs=pairs, thres_dist=np.float64(args.thres_dist),
                        translation_frame=np.float64(args.translation_frame), chunk_size=np.int64(args.chunk_size), **out)
    print("synth_match_dyn: candidate pairs", len(pairs), "selected", out["mp_rows"][:, :2].tolist())


def gen_hist_test_vector(ref):
    torch.manual_seed(2022)
    pts = torch.randn(3, 1000, 3)
    flags = torch.randint(0, 2, size=(3, 1000, 1))
    a = torch.cat([pts, flags], dim=-1)
    b = a.clone()
    b[:, :, 0] += 5.0
    b[:, :, 1] += -3.0
    b[:, :, 2] += -0.2
    rng = (10.0, 10.0, 0.5)
    thres = 0.1
    lens = tuple(len(torch.arange(-r, r + thres, thres)) for r in rng)
    from hist_cuda.hist import hist  # the stub == oracle leaf (the real kernel is CUDA-only)

    h = hist(a, b, -rng[0], -rng[1], -rng[2], rng[0], rng[1], rng[2], *lens)
    _, hh, ww, dd = h.shape
    flat = h.reshape(3, -1).argmax(dim=1)
    arg = torch.stack([flat // dd // ww % hh, flat // dd % ww, flat % dd], dim=1)
    np.savez_compressed(os.path.join(GOLDEN, "hist_test_vector.npz"), X=_np(a), Y=_np(b),
                        mins=np.array([-rng[0], -rng[1], -rng[2]], np.float32),
                        maxs=np.array(rng, np.float32), lens=np.array(lens, np.int64),
                        argmax=_np(arg), peak=_np(h.reshape(3, -1).max(dim=1)[0]), total=_np(h.sum(dim=(1, 2, 3))))
    print("hist_test_vector: argmax", arg.tolist(), "peak", h.reshape(3, -1).max(dim=1)[0].tolist())


def gen_c1_demo(ref):
    from sklearn.cluster import DBSCAN

    d = np.load(os.path.join(ref_loader.REFERENCE_ROOT, "demo.npz"))
    pc_src = d["pc1"][d["pc1_flows_valid_idx"]].astype(np.float32)
    pc_dst = d["pc2"][d["pc2_flows_valid_idx"]].astype(np.float32)
    fused = np.concatenate([pc_dst, pc_src], axis=0)
    labels = DBSCAN(eps=0.25, min_samples=20).fit_predict(fused[:, :3]).astype(np.int64)
    lab_dst, lab_src = labels[: len(pc_dst)], labels[len(pc_dst):]
    args = _args(translation_frame=2.0, max_points=256)
    src_t, dst_t = torch.from_numpy(pc_src), torch.from_numpy(pc_dst)
    ls_t, ld_t = torch.from_numpy(lab_src), torch.from_numpy(lab_dst)
    uniq = torch.unique(torch.cat([ls_t, ld_t]))
    uniq = uniq[uniq >= 0]
    pairs = torch.stack([uniq, uniq], dim=1)
    ok = ref.utils_check.sanity_check(args, src_t, dst_t, ls_t, ld_t, pairs).long()
    chosen = []
    for pr in ok:
        if int((ls_t == pr[0]).sum()) <= 256 and int((ld_t == pr[1]).sum()) <= 256:
            chosen.append(pr)
        if len(chosen) == 32:
            break
    chosen = torch.stack(chosen)
    src_b = torch.stack([ref.utils_helper.pad_segment(src_t[ls_t == pr[0], 0:3], 256) for pr in chosen])
    dst_b = torch.stack([ref.utils_helper.pad_segment(dst_t[ld_t == pr[1], 0:3], 256) for pr in chosen])
    out = _run_path(ref, args, _np(src_b), _np(dst_b))
    # flow over the points of the chosen source clusters (utils_flow.py:57-69), identity ego pose
    sel = torch.isin(ls_t, chosen[:, 0])
    pts, lab = src_t[sel], ls_t[sel]
    flow = ref.utils_flow.flow_estimation_torch(args, pts, None, lab, None, chosen.float(),
                                                torch.from_numpy(out["T_hist_icp"]), torch.eye(4))
    # match_pairs on the same 32 pairs (scan rebuilt from the padded batches so the fixture stays small)
    mp_in = clouds_from_batches(_np(src_b), _np(dst_b), _np(chosen))
    out.update(_run_match_pairs(ref, args, *mp_in, _np(chosen).astype(np.float32)))
    print("c1_demo: match_pairs kept", len(out["mp_rows"]), "of", len(chosen))
    np.savez_compressed(os.path.join(GOLDEN, "c1_demo.npz"), src=_np(src_b), dst=_np(dst_b),
                        thres_dist=np.float64(args.thres_dist), translation_frame=np.float64(args.translation_frame),
                        chunk_size=np.int64(args.chunk_size), pair_labels=_np(chosen),
                        flow_points=_np(pts), flow_labels=_np(lab), flow=_np(flow), **out)
    print("c1_demo: pairs", len(chosen), "icp iterations", int(out["icp_iterations"]),
          "nonzero init", int((np.abs(out["init_pose"][:, :3, 3]).sum(1) > 0).sum()))


def gen_synth_hist(ref):
    src, dst, meta = synth.make_pairs(24, 128, seed=77, ragged=True, residual_only=False, wrong_frac=0.1)
    args = _args(translation_frame=3.333)
    out = _run_path(ref, args, src, dst)
    np.savez_compressed(os.path.join(GOLDEN, "synth_hist.npz"), src=src, dst=dst,
                        thres_dist=np.float64(args.thres_dist), translation_frame=np.float64(args.translation_frame),
                        chunk_size=np.int64(args.chunk_size), gt_translation=meta["translation"], gt_yaw=meta["yaw"],
                        wrong=meta["wrong"], **out)
    print("synth_hist: icp iterations", int(out["icp_iterations"]), "swapped", int(out["swapped"].sum()))


def gen_synth_hist_default(ref):
    """The reference's argparse defaults (main.py: translation_frame 6.666 -> 135 x 135 x 3 bins, the setting of BASELINE
    config C3) on 512-row clusters moved by up to 2 m."""
    src, dst, meta = synth.make_pairs(12, 512, seed=78, ragged=True, residual_only=False, wrong_frac=0.1)
    args = _args(translation_frame=6.666)
    out = _run_path(ref, args, src, dst)
    np.savez_compressed(os.path.join(GOLDEN, "synth_hist_default.npz"), src=src, dst=dst,
                        thres_dist=np.float64(args.thres_dist), translation_frame=np.float64(args.translation_frame),
                        chunk_size=np.int64(args.chunk_size), gt_translation=meta["translation"], gt_yaw=meta["yaw"],
                        wrong=meta["wrong"], **out)
    print("synth_hist_default: icp iterations", int(out["icp_iterations"]), "swapped", int(out["swapped"].sum()),
          "nonzero init", int((np.abs(out["init_pose"][:, :3, 3]).sum(1) > 0).sum()))


def gen_synth_icp20(ref):
    src, dst, meta = synth.make_pairs(32, 256, seed=1234, ragged=False, residual_only=True)
    src_r, dst_r, _ = synth.make_pairs(16, 256, seed=4321, ragged=True, residual_only=True)
    icp = ref.utils_icp_pytorch3d.iterative_closest_point
    res = {}
    for tag, (a, c) in {"full": (src, dst), "ragged": (src_r, dst_r)}.items():
        a_t, c_t = torch.from_numpy(a), torch.from_numpy(c)
        fixed = icp(a_t, c_t, thres=0.1, max_iterations=20, relative_rmse_thr=-1.0)
        stop = icp(a_t, c_t, thres=0.1, max_iterations=100, relative_rmse_thr=1e-6)
        res.update({
            f"{tag}_src": a, f"{tag}_dst": c,
            f"{tag}_fixed20_R": _np(fixed.RTs.R), f"{tag}_fixed20_T": _np(fixed.RTs.T),
            f"{tag}_fixed20_rmse": _np(fixed.rmse), f"{tag}_fixed20_iterations": np.int64(len(fixed.t_history)),
            f"{tag}_stop_R": _np(stop.RTs.R), f"{tag}_stop_T": _np(stop.RTs.T), f"{tag}_stop_rmse": _np(stop.rmse),
            f"{tag}_stop_iterations": np.int64(len(stop.t_history)), f"{tag}_stop_converged": np.bool_(stop.converged),
        })
        print(f"synth_icp20[{tag}]: fixed its", len(fixed.t_history), "stop its", len(stop.t_history), stop.converged)
    np.savez_compressed(os.path.join(GOLDEN, "synth_icp20.npz"), thres_dist=np.float64(0.1), **res)


def gen_frame_demo(ref):
    """Rows f2 / f3: the reference's own match_pcds + flow_estimation_torch on a whole frame pair (torch CPU fp32,
    torch.manual_seed(0) right before match_pcds so that the randperm subsampling of oversized clusters is reproducible
    by anyone making the same calls in the same order)."""
    from sklearn.cluster import DBSCAN

    d = np.load(os.path.join(ref_loader.REFERENCE_ROOT, "demo.npz"))
    pc_src = d["pc1"][d["pc1_flows_valid_idx"]].astype(np.float32)
    pc_dst = d["pc2"][d["pc2_flows_valid_idx"]].astype(np.float32)
    fused = np.concatenate([pc_dst, pc_src], axis=0)
    labels = DBSCAN(eps=0.25, min_samples=20).fit_predict(fused[:, :3]).astype(np.int64)
    big = np.bincount(labels[labels >= 0]).argmax()
    rng = np.random.default_rng(7)
    scans = []
    for pts, lab in ((pc_src, labels[len(pc_dst):]), (pc_dst, labels[: len(pc_dst)])):
        lab = lab.astype(np.float32)
        keep = np.ones(len(pts), bool)
        for value, relabel, quota in ((big, -1e8, 3000), (-1, -1.0, 1500)):     # thin what is not a cluster
            idx = np.flatnonzero(lab == value)
            lab[idx] = relabel
            keep[rng.permutation(idx)[quota:]] = False
        scans.append((pts[keep], lab[keep]))
    (src_pts, src_lab), (dst_pts, dst_lab) = scans
    args = _args(translation_frame=2.0, max_points=512)                        # demo.sh gates
    tens = [torch.from_numpy(x) for x in (src_pts, dst_pts, src_lab, dst_lab)]
    calls = []
    real_sanity = ref.utils_match.sanity_check

    def spy(a, sp, dp, sl, dl, pairs):
        out = real_sanity(a, sp, dp, sl, dl, pairs)
        calls.append((_np(pairs).astype(np.int64).reshape(-1, 2), _np(out).astype(np.int64).reshape(-1, 2)))
        return out

    ref.utils_match.sanity_check = spy
    try:
        torch.manual_seed(0)
        rows, T = ref.utils_match.match_pcds(args, *tens)
    finally:
        ref.utils_match.sanity_check = real_sanity
    yaw = np.deg2rad(1.0)
    pose = np.eye(4, dtype=np.float32)
    pose[:2, :2] = [[np.cos(yaw), -np.sin(yaw)], [np.sin(yaw), np.cos(yaw)]]
    pose[:3, 3] = [0.3, -0.1, 0.02]
    flow = ref.utils_flow.flow_estimation_torch(args, tens[0], tens[1], tens[2], tens[3], rows, T, torch.from_numpy(pose))
    (cand_sta, kept_sta), (cand_dyn, kept_dyn) = calls
    np.savez_compressed(os.path.join(GOLDEN, "frame_demo.npz"), src_points=src_pts, src_labels=src_lab,
                        dst_points=dst_pts, dst_labels=dst_lab, thres_dist=np.float64(args.thres_dist),
                        translation_frame=np.float64(args.translation_frame), chunk_size=np.int64(args.chunk_size),
                        max_points=np.int64(args.max_points), min_cluster_size=np.int64(args.min_cluster_size),
                        thres_box=np.float64(args.thres_box), thres_error=np.float64(args.thres_error),
                        thres_iou=np.float64(args.thres_iou), thres_rot=np.float64(args.thres_rot),
                        static_candidates=cand_sta, static_kept=kept_sta,
                        dynamic_candidates_shape=np.array(cand_dyn.shape), dynamic_kept=kept_dyn,
                        rows=_np(rows), T=_np(T), pose=pose, flow_stride=np.int64(4), flow=_np(flow)[::4])
    sizes = np.bincount(src_lab[src_lab >= 0].astype(np.int64))
    print("frame_demo: points", len(src_pts), len(dst_pts), "static kept", len(kept_sta), "of", len(cand_sta),
          "dynamic kept", len(kept_dyn), "of", len(cand_dyn), "matched", len(rows), "oversized src clusters",
          int((sizes > args.max_points).sum()))


def main():
    torch.set_num_threads(os.cpu_count() or 1)
    os.makedirs(GOLDEN, exist_ok=True)
    ref = ref_loader.load_reference()
    gen_hist_test_vector(ref)
    gen_synth_icp20(ref)
    gen_synth_hist(ref)
    gen_synth_hist_default(ref)
    gen_match_dyn(ref)
    gen_c1_demo(ref)
    gen_frame_demo(ref)


if __name__ == "__main__":
    main()
