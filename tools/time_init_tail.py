#!/usr/bin/env python
"""Where the time of estimate_init_pose goes on C3-shaped data: all-true pairs vs the generator's unrelated pairs.
Run under `ncu --metrics gpu__time_duration.sum -k regex:hist_` for per-kernel times, or plain for CUDA-event totals."""
import os, sys, types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from icp_flow_b200 import ops, synth
dev = torch.device("cuda:0")
args = types.SimpleNamespace(thres_dist=0.1, translation_frame=6.666, chunk_size=50)
src, dst, meta = synth.make_pairs(1024, 1024, seed=99, ragged=False, residual_only=False)
wrong = meta["wrong"]
sets = {"mixed(1024)": np.arange(1024), "true only": np.nonzero(~wrong)[0][:512], "wrong only": np.nonzero(wrong)[0]}
for name, idx in sets.items():
    s, d = torch.from_numpy(src[idx]).to(dev), torch.from_numpy(dst[idx]).to(dev)
    ops.estimate_init_pose(args, s, d, auto_swap=True); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); ops.estimate_init_pose(args, s, d, auto_swap=True); e1.record(); torch.cuda.synchronize()
    print(f"{name:14s} pairs {len(idx):5d}  estimate_init_pose {e0.elapsed_time(e1):7.3f} ms", flush=True)
