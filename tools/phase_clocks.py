import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from icp_flow_b200 import ops, synth
dev = torch.device("cuda:0")
P, N = 1024, 512
s, d, _ = synth.make_pairs(P, N, seed=1234, ragged=False, residual_only=True)
s, d = torch.from_numpy(s).to(dev), torch.from_numpy(d).to(dev)
prm = ops.make_params(thres=0.1, max_iterations=20, relative_rmse_thr=-1.0, early_exit=False, batch_stop=True)
for k in range(3):
    out = ops.icp_batch(s, d, prm)
    torch.cuda.synchronize()
    print("---- run", k, flush=True)
