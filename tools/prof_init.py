import sys, types, torch
sys.path.insert(0, "/root/repo")
from icp_flow_b200 import ops, synth
dev = torch.device("cuda:0")
s, d, _ = synth.make_pairs(1024, 1024, seed=99, ragged=False, residual_only=False)
s, d = torch.from_numpy(s).to(dev), torch.from_numpy(d).to(dev)
args = types.SimpleNamespace(thres_dist=0.1, translation_frame=6.666, chunk_size=50)
for _ in range(3):
    init = ops.estimate_init_pose(args, s, d, auto_swap=True)
torch.cuda.synchronize()
print("ok")
