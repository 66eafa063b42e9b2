#!/usr/bin/env python
"""C4-like "Waymo-shape" frame: ~200 cluster pairs with a log-normal size distribution, padded to max_points=10000
(the reference default).  Times hist_icp and checks a sample against the oracle.  GPU."""
import sys, os, time, types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from icp_flow_b200 import ops, synth
from oracle import icp_oracle as O

rng = np.random.default_rng(4)
P, N = 200, int(sys.argv[1]) if len(sys.argv) > 1 else 10000
sizes = np.clip(np.exp(rng.normal(np.log(80), 1.3, size=P)).astype(int), 20, N)
src = np.full((P, N, 4), 1e8, np.float32); src[..., 3] = 0
dst = src.copy()
for p, n in enumerate(sizes):
    s, d, _ = synth.make_pairs(1, int(n), seed=1000 + p, ragged=False, residual_only=False, wrong_frac=0.0, keep_density=False)
    scale = np.sqrt(n / 512.0)                      # keep ~15 points / m^2
    c = s[0, :, :3].mean(0)
    src[p, :n, :3] = (s[0, :, :3] - c) * scale + c; src[p, :n, 3] = 1
    dst[p, :n, :3] = (d[0, :, :3] - c) * scale + c; dst[p, :n, 3] = 1
print("sizes: median", int(np.median(sizes)), "p90", int(np.percentile(sizes, 90)), "max", sizes.max(), "N", N)
dev = torch.device("cuda:0")
sd, dd = torch.from_numpy(src).to(dev), torch.from_numpy(dst).to(dev)
args = types.SimpleNamespace(thres_dist=0.1, translation_frame=3.34, chunk_size=50)
for _ in range(2):
    T = ops.hist_icp(args, sd, dd)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
    T, dbg = ops.hist_icp(args, sd, dd, return_debug=True)
e1.record(); torch.cuda.synchronize()
print(f"hist_icp({P} pairs, N={N}): {e0.elapsed_time(e1) / 3:.2f} ms; batch iterations {dbg['batch'].tolist()}")
# oracle sample: the 6 smallest pairs re-padded to 256 rows (the oracle is O(N^2) in the padding)
small = np.argsort(sizes)[:6]
m = 256
s6 = np.full((6, m, 4), 1e8, np.float32); s6[..., 3] = 0; d6 = s6.copy()
for k, p in enumerate(small):
    n = sizes[p]; s6[k, :n] = src[p, :n]; d6[k, :n] = dst[p, :n]
want = O.hist_icp(torch.from_numpy(s6), torch.from_numpy(d6), O.PathParams(thres_dist=0.1, translation_frame=3.34))
got = ops.hist_icp(args, torch.from_numpy(s6).to(dev), torch.from_numpy(d6).to(dev)).cpu()
print("oracle sample (alone, N=256) max |dT|:", float((got - want).abs().max()))
