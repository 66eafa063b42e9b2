#!/usr/bin/env python
"""Stage breakdown of hist_icp on the static-stage batch of the C4 scene (GPU): python tools/time_c4_stages.py [N ...]"""
import os, sys, types
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import icp_flow_b200 as E
from icp_flow_b200 import ops, scan, synth

sp, sl, dp, dl, meta = synth.make_scene()
dev = torch.device("cuda:0")
t = [torch.from_numpy(x).to(dev) for x in (sp, dp, sl, dl)]
args = types.SimpleNamespace(thres_dist=0.1, translation_frame=3.34, chunk_size=50, min_cluster_size=30, thres_box=0.1,
                             max_points=10000, thres_error=0.2, thres_iou=0.2, thres_rot=0.1)
si, di = scan.scan_index(t[0], t[2]), scan.scan_index(t[1], t[3])
both = torch.unique(torch.cat([si.present_labels(), di.present_labels()]))
kept, _ = scan.sanity_check_indexed(args, si, di, torch.stack([both, both], 1))
cs, cd = si.counts_host[kept[:, 0].cpu().numpy()], di.counts_host[kept[:, 1].cpu().numpy()]
print("pairs", len(kept), "max valid rows", int(max(cs.max(), cd.max())), "median", int(np.median(np.maximum(cs, cd))))


def timed(fn, reps=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        out = fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, out


for N in [int(a) for a in sys.argv[1:]] or [10000, 4096, 2048]:
    a, b = scan.pad_pairs(si, di, kept, N)
    t_init, init = timed(lambda: ops.estimate_init_pose(args, a, b, auto_swap=True))
    t_apply, _ = timed(lambda: ops.apply_icp(args, a, b, init, auto_swap=True))
    t_all, (T, dbg) = timed(lambda: ops.hist_icp(args, a, b, return_debug=True))
    line = f"N={N}: estimate_init_pose {t_init:.2f} ms; apply_icp {t_apply:.2f} ms; hist_icp {t_all:.2f} ms; batch {dbg['batch'].tolist()}"
    for its in (32, 100):
        prm = ops.make_params(max_iterations=its, early_exit=True, batch_stop=False)
        t_icp, r = timed(lambda: ops.icp_batch(a, b, prm))
        it = r.iterations.cpu().numpy()
        line += f"; icp(max {its}, no batch stop) {t_icp:.2f} ms (iterations median {int(np.median(it))}, at cap {int((it == its).sum())})"
    print(line)
    # which pairs are the long poles: time the 8 largest and the rest separately
    big = np.argsort(-np.maximum(cs, cd))[:8]
    rest = np.setdiff1d(np.arange(len(kept)), big)
    for name, sel in (("8 largest", big), ("the rest", rest)):
        idx = torch.from_numpy(sel).to(dev)
        aa, bb = a[idx].contiguous(), b[idx].contiguous()
        t_h, _ = timed(lambda: ops.hist_icp(args, aa, bb))
        print(f"    {name}: hist_icp {t_h:.2f} ms")
