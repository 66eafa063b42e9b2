#!/usr/bin/env python
"""Host-side profile of one C4 frame pair (match_pcds + flow): cProfile over 20 frames, top cumulative entries, and the
wall clock per frame next to the GPU-busy time (CUDA events around the same call do not separate the two, a frame is
many small launches).   python tools/profile_frame_host.py"""
import cProfile, io, os, pstats, sys, time, types
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import icp_flow_b200 as E
from icp_flow_b200 import scan, synth

a_sp, a_sl, a_dp, a_dl, _ = synth.make_scene()
args = types.SimpleNamespace(thres_dist=0.1, translation_frame=3.34, chunk_size=50, min_cluster_size=30, thres_box=0.1,
                             max_points=10000, thres_error=0.2, thres_iou=0.2, thres_rot=0.1)
sp, dp, sl, dl = (torch.from_numpy(x).cuda() for x in (a_sp, a_dp, a_sl, a_dl))
pose = torch.eye(4, device="cuda")


def frame():
    scan.clear_cache()
    torch.manual_seed(0)
    rows, T = E.match_pcds(args, sp, dp, sl, dl)
    return rows, T, E.flow_estimation_torch(args, sp, dp, sl, dl, rows, T, pose)


for _ in range(5):
    frame()
torch.cuda.synchronize()
t = time.perf_counter()
for _ in range(20):
    frame()
torch.cuda.synchronize()
print(f"wall per frame: {(time.perf_counter() - t) / 20 * 1e3:.2f} ms")
pr = cProfile.Profile()
pr.enable()
for _ in range(20):
    frame()
torch.cuda.synchronize()
pr.disable()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(45)
print(s.getvalue()[:9000])
