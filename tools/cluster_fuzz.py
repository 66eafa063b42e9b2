#!/usr/bin/env python
"""Randomized DBSCAN / HDBSCAN parity against scikit-learn on the GPU: python tools/cluster_fuzz.py [n_cases] [first_seed]"""
import os, sys, warnings
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from icp_flow_b200 import cluster, synth
from oracle import cluster_oracle as CO
warnings.filterwarnings("ignore")
n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 30
seed0 = int(sys.argv[2]) if len(sys.argv) > 2 else 100
bad_db = bad_hd = 0
pts_total = 0
for i in range(n_cases):
    rng = np.random.default_rng(seed0 + i)
    n = int(rng.choice([800, 2000, 4000]))
    if i % 3 == 0:          # blobs + clutter, some duplicated points, a quantised axis (many exactly equal distances)
        pts = np.concatenate([rng.normal(rng.uniform(-8, 8, 3), rng.uniform(0.05, 0.4), (int(rng.integers(20, 200)), 3)) for _ in range(12)]
                             + [rng.uniform(-9, 9, (300, 3))])[:n]
        pts[:, 2] = np.round(pts[:, 2], 1)
        pts = np.concatenate([pts, pts[:40]])
    else:
        sp, sl, _, _, _ = synth.make_scene(num_clusters=int(rng.integers(6, 30)), num_points=3 * n, seed=seed0 + i, max_size=int(rng.choice([200, 600])))
        pts = sp[sl > -1e7][:n, :3]
    pts = np.ascontiguousarray(pts, dtype=np.float32)
    pts_total += len(pts)
    eps, mp = float(rng.choice([0.15, 0.25, 0.4])), int(rng.choice([5, 10, 30]))
    t = torch.from_numpy(pts).cuda()
    if not np.array_equal(cluster.dbscan_labels(t, eps, mp).cpu().numpy(), CO.dbscan_labels(pts, eps, mp)):
        bad_db += 1
        print("DBSCAN mismatch", seed0 + i, eps, mp)
    mcs = int(rng.choice([5, 10, 20, 30]))
    if not CO.same_partition(cluster.hdbscan_labels(t, mcs), CO.hdbscan_labels(pts, mcs)):
        bad_hd += 1
        print("HDBSCAN mismatch", seed0 + i, mcs, len(pts))
print(f"cluster fuzz (seeds {seed0}..{seed0 + n_cases - 1}): {n_cases} scans, {pts_total} points; DBSCAN label mismatches {bad_db}, "
      f"HDBSCAN partition mismatches {bad_hd}")
