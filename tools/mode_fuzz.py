#!/usr/bin/env python
"""The three NN modes of the ICP kernel (1 brute force, 2 grid, 3 grid + correspondence cache) must agree bit for bit:
    python tools/mode_fuzz.py [n_batches] [first_seed]        (GPU)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from icp_flow_b200 import ops, synth
n_batches = int(sys.argv[1]) if len(sys.argv) > 1 else 100
seed0 = int(sys.argv[2]) if len(sys.argv) > 2 else 5000
bad = pairs = 0
for i in range(n_batches):
    rng = np.random.default_rng(seed0 + i)
    P, N = int(rng.integers(4, 65)), int(rng.choice([32, 64, 160, 300, 512, 700, 1024, 2048]))
    src, dst, _ = synth.make_pairs(P, N, seed=seed0 + i, ragged=bool(rng.integers(0, 2)), residual_only=bool(rng.integers(0, 2)),
                                   wrong_frac=float(rng.choice([0.0, 0.1, 0.3])))
    s, d = torch.from_numpy(src).cuda(), torch.from_numpy(dst).cuda()
    kw = dict(thres=float(rng.choice([0.05, 0.1, 0.2])), max_iterations=int(rng.choice([5, 20, 100])),
              relative_rmse_thr=float(rng.choice([-1.0, 1e-6])), early_exit=bool(rng.integers(0, 2)))
    outs = []
    for mode in (1, 2, 3):
        r = ops.icp_batch(s, d, ops.make_params(nn_mode=mode, **kw))
        outs.append([x.clone() for x in (r.R, r.T, r.rmse, r.iterations, r.batch, r.conv_mask, r.pose)])
    same = all(torch.equal(a, b) or (a.is_floating_point() and torch.equal(a.nan_to_num(-7.0), b.nan_to_num(-7.0)))
               for o in outs[1:] for a, b in zip(outs[0], o))
    pairs += P
    if not same:
        bad += 1
        print("modes differ:", seed0 + i, P, N, kw)
print(f"mode fuzz (seeds {seed0}..{seed0 + n_batches - 1}): {n_batches} batches, {pairs} pairs; batches where the NN modes differ in any bit: {bad}")
