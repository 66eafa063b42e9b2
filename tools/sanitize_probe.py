"""Small end-to-end workload for compute-sanitizer (memcheck / racecheck / synccheck)."""
import sys, os, types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from icp_flow_b200 import ops, synth
dev = torch.device("cuda:0")
src, dst, _ = synth.make_pairs(6, 160, seed=3, ragged=True, residual_only=False, wrong_frac=0.2)
s, d = torch.from_numpy(src).to(dev), torch.from_numpy(dst).to(dev)
args = types.SimpleNamespace(thres_dist=0.1, translation_frame=2.0, chunk_size=50)
T = ops.hist_icp(args, s, d)
for mode in (1, 2, 3):
    r = ops.icp_batch(s, d, ops.make_params(max_iterations=12, nn_mode=mode))
idx, dist = ops.nearest_neighbor_batch(s, d)
big = torch.full((2, 5000, 4), 1e8, device=dev); big[:, :, 3] = 0
big[:, :160] = s[:2]
big2 = big.clone(); big2[:, :160] = d[:2]
T2 = ops.hist_icp(args, big, big2)
# unrelated clusters with more rows than the scoring budget: deferred (pair, candidate) items + the warp scans of far queries
src3, dst3, _ = synth.make_pairs(8, 700, seed=8, ragged=False, residual_only=False, wrong_frac=0.5)
args3 = types.SimpleNamespace(thres_dist=0.1, translation_frame=6.666, chunk_size=50)
T3, dbg3 = ops.hist_icp(args3, torch.from_numpy(src3).to(dev), torch.from_numpy(dst3).to(dev), return_debug=True)
ev3 = ops.match_eval(args3, torch.from_numpy(src3).to(dev), torch.from_numpy(dst3).to(dev), T3)
torch.cuda.synchronize()
print("ok", float(T.abs().sum()), float(T2.abs().sum()), float(T3.abs().sum()))
# rows f1-f3: a small scene through the cluster index, sanity_check, gather/pad (with an oversized cluster), match_eval
# (NN grids), selection and flow
import icp_flow_b200 as E
sp, sl, dp, dl, meta = synth.make_scene(num_clusters=12, num_points=6000, seed=2, max_size=700)
t = [torch.from_numpy(x).to(dev) for x in (sp, dp, sl, dl)]
fargs = types.SimpleNamespace(thres_dist=0.1, translation_frame=3.34, chunk_size=50, min_cluster_size=30, thres_box=0.1,
                              max_points=256, thres_error=0.2, thres_iou=0.2, thres_rot=0.1)
torch.manual_seed(0)          # (the oversized cluster is subsampled with torch.randperm, like the reference)
rows, Tm = E.match_pcds(fargs, *t)
flow = E.flow_estimation_torch(fargs, t[0], t[1], t[2], t[3], rows, Tm, torch.eye(4, device=dev))
torch.cuda.synchronize()
print("frame ok", len(rows), float(flow.abs().sum()))
# row f4: DBSCAN of the scene's non-ground points
from icp_flow_b200 import cluster
lab = cluster.dbscan_labels(t[0][t[2] > -1e7][:, :3].contiguous(), 0.25, 20)
torch.cuda.synchronize()
print("dbscan ok", int(lab.max()) + 1)
hl = cluster.hdbscan_labels(t[0][t[2] > -1e7][:1500, :3].contiguous(), 20)
torch.cuda.synchronize()
print("hdbscan ok", int(hl.max()) + 1, int((hl < 0).sum()))
hf = cluster.hdbscan_labels(t[0][t[2] > -1e7][:1500, :3].contiguous(), 20, exact_order=False)
torch.cuda.synchronize()
print("hdbscan any-order ok", int(hf.max()) + 1, int((hf < 0).sum()))
