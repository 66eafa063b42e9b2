"""Stage timings of a whole frame pair (rows f2 + path + f1 + f3): wall clock around torch.cuda.synchronize(), best of 5.

    python tools/time_frame.py            # the frame_demo fixture (demo.npz scene, max_points 512, demo.sh gates)
    python tools/time_frame.py c4 [N]     # BASELINE config C4: synthetic Waymo-shape frame pair, 150k points, 200 clusters,
                                          # max_points N (default 10000 = the reference default), F = 3.34
"""
import os, sys, time, types
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import icp_flow_b200 as E
from icp_flow_b200 import scan

if len(sys.argv) > 1 and sys.argv[1] == "c4":
    from icp_flow_b200 import synth
    a_sp, a_sl, a_dp, a_dl, _ = synth.make_scene()
    g = {"src_points": a_sp, "src_labels": a_sl, "dst_points": a_dp, "dst_labels": a_dl, "pose": np.eye(4, dtype=np.float32)}
    args = types.SimpleNamespace(thres_dist=0.1, translation_frame=3.34, chunk_size=50, min_cluster_size=30, thres_box=0.1,
                                 max_points=int(sys.argv[2]) if len(sys.argv) > 2 else 10000, thres_error=0.2,
                                 thres_iou=0.2, thres_rot=0.1)                # main.py defaults, demo.sh error / iou gates
    sl_np, dl_np = a_sl.astype(np.int64), a_dl.astype(np.int64)
    both = np.union1d(sl_np[sl_np >= 0], dl_np[dl_np >= 0])
    g["static_candidates"] = np.stack([both, both], 1)
else:
    g = dict(np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "frame_demo.npz")))
    args = types.SimpleNamespace(**{k: (int(g[k]) if k in ("chunk_size", "max_points", "min_cluster_size") else float(g[k]))
                                    for k in ("thres_dist", "translation_frame", "chunk_size", "max_points",
                                              "min_cluster_size", "thres_box", "thres_error", "thres_iou", "thres_rot")})
sp, dp, sl, dl = (torch.from_numpy(g[k]).cuda() for k in ("src_points", "dst_points", "src_labels", "dst_labels"))
pose = torch.from_numpy(g["pose"]).cuda()


def timed(fn, reps=5):
    best, out = 1e9, None
    for _ in range(reps):
        torch.cuda.synchronize(); t = time.perf_counter(); out = fn(); torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t)
    return best * 1e3, out


def frame():
    scan.clear_cache()
    torch.manual_seed(0)
    rows, T = E.match_pcds(args, sp, dp, sl, dl)
    return rows, T, E.flow_estimation_torch(args, sp, dp, sl, dl, rows, T, pose)


ms, (rows, T, flow) = timed(frame)
print(f"whole frame pair (index + match_pcds + flow): {ms:.2f} ms, {len(rows)} matched pairs, {len(sp)} + {len(dp)} points")
ms, (si, di) = timed(lambda: (E.ScanIndex(sp, sl), E.ScanIndex(dp, dl)))
print(f"  cluster index x2: {ms:.3f} ms")
cand = torch.from_numpy(g["static_candidates"]).cuda()
ms, _ = timed(lambda: scan.sanity_check_indexed(args, si, di, cand))
print(f"  sanity_check static ({len(cand)} candidates): {ms:.3f} ms")
sl_u, dl_u = si.present_labels(), di.present_labels()
cross = torch.stack([sl_u.repeat_interleave(len(dl_u)), dl_u.repeat(len(sl_u))], 1)
ms, _ = timed(lambda: scan.sanity_check_indexed(args, si, di, cross))
print(f"  sanity_check all x all ({len(cross)} candidates): {ms:.3f} ms")
kept, _ = scan.sanity_check_indexed(args, si, di, cand)
ms, (a, b) = timed(lambda: scan.pad_pairs(si, di, kept, args.max_points))
print(f"  gather + pad ({len(kept)} pairs x {args.max_points}, {int((si.counts_host > args.max_points).sum())} oversized src clusters): {ms:.3f} ms")
ms, Tp = timed(lambda: E.hist_icp(args, a, b))
print(f"  hist_icp: {ms:.3f} ms")
ms, _ = timed(lambda: E.match_eval(args, a, b, Tp, return_accept=True))
print(f"  match_eval: {ms:.3f} ms")
ms, _ = timed(lambda: E.flow_estimation_torch(args, sp, dp, sl, dl, rows, T, pose))
print(f"  flow: {ms:.3f} ms")
