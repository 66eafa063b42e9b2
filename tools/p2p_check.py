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
r = ops.icp_batch(s, d, prm, ext=peer.ext(0))
full = peer.finish(0).clone()
want = shard.gather_transforms(r.pose, world * P)
torch.cuda.synchronize()
ok = torch.equal(full, want) and torch.equal(full[rank * P:(rank + 1) * P], r.pose)
print(f"rank {rank}: fused peer gather == NCCL all_gather: {ok}", flush=True)
assert ok
# the un-fused variant: one small launch pushes the rank's block to every peer
peer.slots[0][0].zero_()
dist.barrier()
peer.push(r.pose, 0)
full = peer.finish(0)
torch.cuda.synchronize()
ok = torch.equal(full, want)
print(f"rank {rank}: peer push == NCCL all_gather: {ok}", flush=True)
assert ok
# the sharded hist_icp with the exchanged batch stop == the unsharded call on the whole batch, bit for bit
import types
args = types.SimpleNamespace(thres_dist=0.1, translation_frame=2.0, chunk_size=50)
gs, gd, _ = synth.make_pairs(11, 160, seed=14, ragged=True, residual_only=False, wrong_frac=0.2)
whole = ops.hist_icp(args, torch.from_numpy(gs).to(dev), torch.from_numpy(gd).to(dev))
lo, hi = shard.shard_range(len(gs), rank, world)
got = shard.hist_icp_sharded(args, torch.from_numpy(gs[lo:hi]).to(dev), torch.from_numpy(gd[lo:hi]).to(dev))
torch.cuda.synchronize()
ok2 = torch.equal(got, whole)
print(f"rank {rank}: sharded hist_icp == unsharded: {ok2}", flush=True)
assert ok2
dist.barrier(); dist.destroy_process_group()
