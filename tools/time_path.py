#!/usr/bin/env python
"""Time the stages of the full hist_icp path on a synthetic batch (GPU): python tools/time_path.py P N F"""
import sys, os, time, types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from icp_flow_b200 import ops, synth

P, N, F = int(sys.argv[1]), int(sys.argv[2]), float(sys.argv[3])
t0 = time.time()
src, dst, meta = synth.make_pairs(P, N, seed=99, ragged=False, residual_only=False)
print(f"generated {P}x{N} in {time.time() - t0:.1f}s")
dev = torch.device("cuda:0")
s, d = torch.from_numpy(src).to(dev), torch.from_numpy(dst).to(dev)
args = types.SimpleNamespace(thres_dist=0.1, translation_frame=F, chunk_size=50)

def timed(fn, reps=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        out = fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, out

hb = ops._hist_bins(args.thres_dist, args.translation_frame, dev)
t_votes, _ = timed(lambda: ops.hist(d[:256], s[:256], *hb.c.min, *hb.c.max, *hb.lens))
t_init, init = timed(lambda: ops.estimate_init_pose(args, s, d, auto_swap=True))
t_apply, _ = timed(lambda: ops.apply_icp(args, s, d, init, auto_swap=True))
t_all, (T, dbg) = timed(lambda: ops.hist_icp(args, s, d, return_debug=True))
print(f"bins {hb.lens}; votes(256 pairs) {t_votes:.2f} ms; estimate_init_pose {t_init:.2f} ms; apply_icp {t_apply:.2f} ms; "
      f"hist_icp {t_all:.2f} ms -> {P / t_all * 1e3:.0f} pairs/s; batch iterations {dbg['batch'].tolist()}")
T = T.cpu().numpy()
R = T[:, :3, :3].astype(np.float64)
print("finite", np.isfinite(T).all(), "orth err", np.abs(R @ R.transpose(0, 2, 1) - np.eye(3)).max(), "det min", np.linalg.det(R).min())
ok = ~meta["wrong"]
t_err = np.abs(dbg["init"].cpu().numpy()[ok, :3, 3] - meta["translation"][ok]).max(axis=1)
print("init translation within 0.15 m of ground truth on", float((t_err < 0.15).mean()), "of the true-match pairs")
# row f1: match_eval (+ fused check_transformation) on the same batch
eargs = types.SimpleNamespace(thres_dist=0.1, translation_frame=F, thres_iou=0.2, thres_rot=0.1)
Tdev = torch.from_numpy(T).to(dev)
t_eval, ev = timed(lambda: ops.match_eval(eargs, s, d, Tdev, return_accept=True))
n_valid = float((s[:, :, 3] > 0).sum() + (d[:, :, 3] > 0).sum())
print(f"match_eval {t_eval:.2f} ms -> {P / t_eval * 1e3:.0f} pairs/s; accepted {int(ev[6].sum())}/{P}; "
      f"mean inlier ratio {float(ev[2].nanmean()):.3f}; brute-force candidates/s {float(((s[:, :, 3] > 0).sum(1).double() * (d[:, :, 3] > 0).sum(1).double()).sum()) * 2 / t_eval * 1e3:.3e}")
