"""torchrun, >= 2 GPUs: the fused peer-store gather must equal an NCCL all_gather of the same transforms."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from icp_flow_b200 import ops, synth, shard
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank); dev = torch.device("cuda", rank)
dist.init_process_group("nccl", device_id=dev)
P = 96
src, dst, _ = synth.make_pairs(P, 256, seed=100 + rank, ragged=True, residual_only=True)
s, d = torch.from_numpy(src).to(dev), torch.from_numpy(dst).to(dev)
peer = shard.PeerGather(P, dev, slots=1)
prm = ops.make_params()
peer.arm(0)
r = ops.icp_batch(s, d, prm)
full = peer.finish(0)
want = shard.gather_transforms(r.pose, world * P)
torch.cuda.synchronize()
ok = torch.equal(full, want) and torch.equal(full[rank * P:(rank + 1) * P], r.pose)
print(f"rank {rank}: fused peer gather == NCCL all_gather: {ok}", flush=True)
assert ok
dist.barrier(); dist.destroy_process_group()
