#!/usr/bin/env python
"""Randomized whole-frame parity (match_pcds, both stages) against the CPU oracle on the GPU:
    python tools/frame_fuzz.py [n_scenes] [first_seed]
Same matched label pairs in the same order; every selected transform within 1e-4 m of the oracle's or adjudicated."""
import os, sys, types
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import icp_flow_b200 as E
from icp_flow_b200 import synth
from oracle import icp_oracle as O
from parity import selected_pairs_parity

n_scenes = int(sys.argv[1]) if len(sys.argv) > 1 else 10
seed0 = int(sys.argv[2]) if len(sys.argv) > 2 else 300
rows_total = flagged_total = order_diff = 0
for i in range(n_scenes):
    rng = np.random.default_rng(seed0 + i)
    sp, sl, dp, dl, meta = synth.make_scene(num_clusters=int(rng.integers(12, 40)), num_points=int(rng.integers(6000, 14000)),
                                            seed=seed0 + i, median_size=int(rng.integers(50, 110)), sigma=1.0, max_size=900,
                                            dynamic_frac=float(rng.uniform(0.1, 0.3)))
    F = float(rng.choice([2.0, 3.34]))
    p = O.PathParams(thres_dist=0.1, translation_frame=F, chunk_size=50)
    gates = O.MatchGates(max_points=512, thres_error=0.2, thres_iou=0.2, thres_rot=0.1, min_cluster_size=30, thres_box=0.1)
    args = types.SimpleNamespace(thres_dist=0.1, translation_frame=F, chunk_size=50, max_points=512, thres_error=0.2,
                                 thres_iou=0.2, thres_rot=0.1, min_cluster_size=30, thres_box=0.1)
    scan_t = [torch.from_numpy(x) for x in (sp, dp, sl, dl)]
    torch.manual_seed(0)
    ref_rows, ref_T, dbg = O.match_pcds(*scan_t, p, gates, return_debug=True)
    torch.manual_seed(0)
    rows, T = E.match_pcds(args, *[t.cuda() for t in scan_t])
    rows, T, ref_rows, ref_T = rows.cpu().numpy(), T.cpu().numpy(), ref_rows.numpy(), ref_T.numpy()
    if rows.shape != ref_rows.shape or not np.array_equal(rows[:, :2], ref_rows[:, :2]):
        order_diff += 1
        print(f"scene {seed0 + i}: matched pairs differ: engine {rows[:, :2].tolist()} oracle {ref_rows[:, :2].tolist()}")
        continue
    stages = [(dbg[st]["segs_src"], dbg[st]["segs_dst"], dbg[st]["pairs"]) for st in ("static", "dynamic") if st in dbg and len(dbg[st]["pairs"])]
    its = lambda a, b: E.hist_icp(args, a.cuda(), b.cuda(), return_debug=True)[1]["batch"].tolist()[0]
    flagged = selected_pairs_parity(stages, p, ref_rows, T, ref_T, max_explained=0.5, what=f"scene {seed0 + i}", engine_its=its)
    rows_total += len(rows)
    flagged_total += int(flagged.sum())
print(f"frame fuzz (seeds {seed0}..{seed0 + n_scenes - 1}): {n_scenes} frame pairs, {rows_total} matched cluster pairs compared, "
      f"{flagged_total} adjudicated, {order_diff} scenes with a different set of matched pairs")
