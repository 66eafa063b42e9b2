"""Probe: does torch symmetric memory work on this box? (torchrun, 2 ranks)"""
import os, torch, torch.distributed as dist
import torch.distributed._symmetric_memory as symm
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
t = symm.empty(world * 8, 16, dtype=torch.float32, device=torch.device("cuda", rank))
t.zero_()
h = symm.rendezvous(t, dist.group.WORLD)
print(rank, "ptrs", [hex(p) for p in h.buffer_ptrs], "dev", hex(h.buffer_ptrs_dev), "multicast", h.has_multicast_support("cuda", rank) if hasattr(h, "has_multicast_support") else None, flush=True)
# write my block into every peer through get_buffer
for r in range(world):
    peer = h.get_buffer(r, (world * 8, 16), torch.float32)
    peer[rank * 8:(rank + 1) * 8] = float(rank + 1)
h.barrier()
torch.cuda.synchronize()
print(rank, "sum per block", [float(t[r * 8:(r + 1) * 8].mean()) for r in range(world)], flush=True)
dist.destroy_process_group()
