#!/usr/bin/env python
"""Per-pair iteration statistics of the engine on the C1 golden (needs a GPU): how early pairs reach their fixed point."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from icp_flow_b200 import ops, synth
g = dict(np.load("tests/golden/c1_demo.npz"))
dev = torch.device("cuda:0")
for name, (s, d) in {"c1": (g["src"], g["dst"]), "synth512": synth.make_pairs(256, 512, seed=7, residual_only=True)[:2]}.items():
    r = ops.icp_batch(torch.from_numpy(s).to(dev), torch.from_numpy(d).to(dev), ops.make_params(batch_stop=False))
    it = r.iterations.cpu().numpy()
    print(name, "iterations: median", np.median(it), "p90", np.percentile(it, 90), "max", it.max(), "n==100:", int((it == 100).sum()), "of", len(it))
