#!/usr/bin/env python
"""HDBSCAN of a synthetic scan on the GPU against scikit-learn on the host: python tools/time_hdbscan.py [n_points] [min_cluster_size]"""
import os, sys, time, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from icp_flow_b200 import cluster, synth
warnings.filterwarnings("ignore")
n_want = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
m = int(sys.argv[2]) if len(sys.argv) > 2 else 30
sp, sl, dp, dl, _ = synth.make_scene(num_clusters=200, num_points=max(3 * n_want, 30000), seed=4)
X = sp[sl > -1e7][:, :3][:n_want].astype(np.float32)
pts = torch.from_numpy(X).cuda()
cluster.hdbscan_labels(pts[:2000], m)
torch.cuda.synchronize()
t = time.perf_counter(); got = cluster.hdbscan_labels(pts, m); torch.cuda.synchronize(); t_gpu = time.perf_counter() - t
print(f"n = {len(X)}, min_cluster_size = {m}: engine {t_gpu * 1e3:.1f} ms, {got.max() + 1} clusters, {(got < 0).sum()} noise points")
if "--no-oracle" not in sys.argv:
    from sklearn.cluster import HDBSCAN
    t = time.perf_counter(); ref = HDBSCAN(min_cluster_size=m, algorithm="kd_tree", leaf_size=100).fit(X.astype(np.float64)).labels_; t_cpu = time.perf_counter() - t
    same_noise = np.array_equal(ref < 0, got < 0)
    pairs = set(zip(ref[ref >= 0].tolist(), got[got >= 0].tolist()))
    same = same_noise and len(pairs) == len({p[0] for p in pairs}) == len({p[1] for p in pairs})
    print(f"sklearn.cluster.HDBSCAN (kd_tree, one core): {t_cpu:.2f} s, {ref.max() + 1} clusters; identical partition: {same}; speed-up {t_cpu / t_gpu:.0f}x")
