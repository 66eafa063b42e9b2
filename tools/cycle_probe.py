#!/usr/bin/env python
"""Debug (GPU): show how the state of the C1 pairs that never reach a bitwise fixed point evolves."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from icp_flow_b200 import ops
g = dict(np.load("tests/golden/c1_demo.npz"))
dev = torch.device("cuda:0")
s, d = torch.from_numpy(g["src"]).to(dev), torch.from_numpy(g["dst"]).to(dev)
r = ops.icp_batch(s, d, ops.make_params(batch_stop=False))
bad = (r.iterations.cpu().numpy() == 100).nonzero()[0]
print("never fixed:", bad)
hist = []
for k in range(40, 52):
    rk = ops.icp_batch(s, d, ops.make_params(max_iterations=k, early_exit=False, batch_stop=False))
    hist.append(torch.cat([rk.R.reshape(-1, 9), rk.T], dim=1).cpu().numpy().view(np.uint32))
hist = np.stack(hist)            # [k, P, 12]
for p in bad[:4]:
    h = hist[:, p]
    same_prev = [(h[i] == h[i - 1]).all() for i in range(1, len(h))]
    same_prev2 = [(h[i] == h[i - 2]).all() for i in range(2, len(h))]
    diffbits = [int(np.abs(h[i].astype(np.int64) - h[i - 1].astype(np.int64)).max()) for i in range(1, len(h))]
    print("pair", p, "n_s", int((g["src"][p, :, 3] > 0).sum()), "n_d", int((g["dst"][p, :, 3] > 0).sum()))
    print("  == previous:", same_prev)
    print("  == two back:", same_prev2)
    print("  max ulp diff to previous:", diffbits)
