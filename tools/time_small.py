#!/usr/bin/env python
"""Latency of small calls (C1: 32 pairs x 256 points, full hist_icp; and the C2 step) -- where launch counts matter."""
import os, sys, types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import icp_flow_b200 as E
from icp_flow_b200 import ops, synth
dev = torch.device("cuda:0")
g = dict(np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "c1_demo.npz")))
s1, d1 = torch.from_numpy(g["src"]).to(dev), torch.from_numpy(g["dst"]).to(dev)
a1 = types.SimpleNamespace(thres_dist=0.1, translation_frame=2.0, chunk_size=50)
def ev(fn, reps):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(5):
        e0.record()
        for _ in range(reps): fn()
        e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / reps)
    return best
print("lib", os.path.basename(os.environ.get("ICPF_LIB_PATH", "product")), "C1 hist_icp ms", round(ev(lambda: E.hist_icp(a1, s1, d1), 50), 4), flush=True)
