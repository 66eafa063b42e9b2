#!/usr/bin/env python
"""Extended randomized parity of hist_icp against the CPU oracle on the GPU, beyond the test-suite's seeds:
    python tools/parity_fuzz.py [n_batches] [first_seed]
Every pair must be within 1e-4 m of the oracle or adjudicated (oracle/adjudicate.py); prints one summary line."""
import os, sys, types
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from icp_flow_b200 import ops, synth
from oracle import icp_oracle as O
from parity import assert_path_parity

n_batches = int(sys.argv[1]) if len(sys.argv) > 1 else 24
seed0 = int(sys.argv[2]) if len(sys.argv) > 2 else 9000
rng = np.random.default_rng(seed0)
total = explained = 0
worst = 0.0
kinds = []
for i in range(n_batches):
    P, N = 8, int(rng.choice([64, 160, 300, 512, 700]))
    F = float(rng.choice([2.0, 3.333, 6.666]))
    src, dst, _ = synth.make_pairs(P, N, seed=seed0 + i, ragged=bool(i % 2), residual_only=(F == 2.0), wrong_frac=0.15)
    args = types.SimpleNamespace(thres_dist=0.1, translation_frame=F, chunk_size=50)
    p = O.PathParams(thres_dist=0.1, translation_frame=F)
    s_t, d_t = torch.from_numpy(src), torch.from_numpy(dst)
    want, odbg = O.hist_icp(s_t, d_t, p, return_debug=True)
    got, dbg = ops.hist_icp(args, s_t.cuda(), d_t.cuda(), return_debug=True)
    sw = odbg["swapped"]
    a_, c_ = s_t.clone(), d_t.clone()
    a_[sw] = d_t[sw]
    c_[sw] = s_t[sw]
    trace = O.icp_loop(O.transform_points_batch(a_, odbg["init"]), c_, 0.1, 100, 1e-6, diagnostics=True)
    v = assert_path_parity(src, dst, got.cpu(), want, p, trace.iterations, dbg["batch"].tolist()[0], max_explained=0.5,
                           what=f"fuzz batch {i} (N={N}, F={F})", trace=trace)
    total += P
    explained += int(v.explained.sum())
    kinds += [x for x in v.verdict if x != "ok"]
    ok = np.array([x == "ok" for x in v.verdict])
    worst = max(worst, float(v.err[ok].max()) if ok.any() else 0.0)
print(f"parity fuzz (seeds {seed0}..{seed0 + n_batches - 1}): {total} pairs, {explained} adjudicated {kinds}, 0 unexplained, "
      f"worst error of the others {worst:.2e} m")
