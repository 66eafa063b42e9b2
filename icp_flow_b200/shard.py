"""Multi-GPU sharding of the per-cluster-pair path: pairs are independent units, so each rank registers a contiguous
block of pairs and the 4x4 transforms (64 B per pair) are all-gathered -- the only collective of the path
(SURVEY.md section 8e).  One process per GPU, ``torch.distributed`` (NCCL on GPUs; gloo works for the host logic)."""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.distributed as dist


def shard_range(num_pairs: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous block [lo, hi) of pairs owned by ``rank``; the first ``num_pairs % world_size`` ranks get one more."""
    if world_size <= 0 or not (0 <= rank < world_size):
        raise ValueError("invalid rank / world_size")
    base, extra = divmod(int(num_pairs), world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def gather_transforms(local: torch.Tensor, num_pairs: int, group: Optional[dist.ProcessGroup] = None) -> torch.Tensor:
    """All-gather the per-rank ``[p_local, 4, 4]`` transforms into ``[num_pairs, 4, 4]`` on every rank.

    Shards may differ by one pair; every rank contributes a block padded to the largest shard (one fixed-size
    ``all_gather_into_tensor`` over NVLink / NVSwitch, 64 B per pair) and the padding is dropped on arrival."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    lo, hi = shard_range(num_pairs, rank, world)
    if local.shape[0] != hi - lo:
        raise ValueError(f"rank {rank} owns pairs [{lo}, {hi}) but holds {local.shape[0]} transforms")
    biggest = -(-int(num_pairs) // world)
    send = local.reshape(hi - lo, 16)
    if hi - lo < biggest:
        send = torch.cat([send, send.new_zeros(biggest - (hi - lo), 16)], dim=0)
    recv = send.new_empty(world * biggest, 16)
    dist.all_gather_into_tensor(recv, send.contiguous(), group=group)
    if num_pairs % world == 0:
        return recv.view(num_pairs, 4, 4)
    parts = []
    for r in range(world):
        rlo, rhi = shard_range(num_pairs, r, world)
        parts.append(recv[r * biggest: r * biggest + (rhi - rlo)])
    return torch.cat(parts, dim=0).view(num_pairs, 4, 4)


def batch_stop_from_masks(conv_mask: torch.Tensor, max_iterations: int,
                          group: Optional[dist.ProcessGroup] = None) -> Tuple[int, bool]:
    """The reference's batch stop (utils_icp_pytorch3d.py:209) over the pairs of ALL ranks, from the per-pair convergence
    masks ``icp_batch`` returns (``[p_local, 4]`` int32, bit k <=> relative rmse <= thr at iteration k): the first
    iteration at which every pair of every rank passes the test.  Returns ``(iterations, converged)`` like
    ``IcpBatchResult.batch`` -- ``(k* + 1, True)`` or ``(max_iterations, False)``.

    One 16-byte exchange per rank (all-gather of the local AND; NCCL has no bitwise reduction).  This is the building
    block for a sharded ``hist_icp`` that stops where the unsharded batch would."""
    words = conv_mask.reshape(-1, 4).to(torch.int32)
    local = torch.full((4,), -1, dtype=torch.int32, device=words.device)
    if words.shape[0] > 0:
        local = _and_rows(words).clone()                            # AND-reduce over the pairs
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world > 1:
        parts = [torch.empty_like(local) for _ in range(world)]
        dist.all_gather(parts, local, group=group)
        for other in parts:
            local &= other
    bits = local.cpu().numpy().astype("uint32")
    for k in range(min(int(max_iterations), 128)):
        if (int(bits[k >> 5]) >> (k & 31)) & 1:
            return k + 1, True
    return int(max_iterations), False


def _and_rows(rows: torch.Tensor) -> torch.Tensor:
    """Bitwise AND over dim 0 of an int32 ``[n, 4]`` tensor (log-depth halving; torch has no AND reduction)."""
    while rows.shape[0] > 1:
        half = rows.shape[0] // 2
        head = rows[:half] & rows[half:2 * half]
        rows = torch.cat([head, rows[2 * half:]], dim=0) if rows.shape[0] % 2 else head
    return rows[0]


def _stop_from_words(words: torch.Tensor, limit: int, group) -> Tuple[int, bool]:
    """Lowest bit below ``limit`` that is set in the AND of every rank's four mask words -> (k* + 1, True)."""
    return batch_stop_from_masks(words.reshape(1, 4), limit, group)


def hist_icp_sharded(args, src: torch.Tensor, dst: torch.Tensor, group: Optional[dist.ProcessGroup] = None,
                     exact_stop: bool = True):
    """``hist_icp`` over the pairs of ALL ranks: ``src`` / ``dst`` are this rank's shard (``shard_range``) of the
    global padded batch; returns the transforms of every pair, ``[num_pairs, 4, 4]``, on every rank.

    ``exact_stop`` (default): the reference's batch stop looks at all pairs of the call, so the shards exchange the AND
    of their convergence masks (16 bytes, once or twice) and stop where the unsharded batch would -- the result is the
    unsharded ``hist_icp`` bit for bit.  ``exact_stop=False`` applies the stop per shard (no exchange, no host sync):
    pairs still moving when their shard's stop fires may end a few iterations earlier or later."""
    from . import ops

    counts = torch.tensor([src.shape[0]], device=src.device, dtype=torch.int64)
    dist.all_reduce(counts, group=group)
    if not exact_stop:
        local = ops.hist_icp(args, src, dst)
        return gather_transforms(local, int(counts.item()), group)
    # the histogram initialisation is per pair; utils_match.hist_icp registers the smaller cloud onto the larger one
    init = ops.estimate_init_pose(args, src, dst, auto_swap=True) if src.shape[0] > 0 else src.new_empty(0, 4, 4)
    ph = ops.ApplyIcpPhases(args, src, dst, init, auto_swap=True)
    iterations, converged = _stop_from_words(ph.first_pass(), ph.cap, group)
    if not converged and ph.cap < ph.max_iterations:
        iterations, converged = _stop_from_words(ph.full_pass(), ph.max_iterations, group)
    if not converged:
        iterations = ph.max_iterations
    local = ph.finish(iterations, converged)
    return gather_transforms(local, int(counts.item()), group)


class PeerGather:
    """Fused all-gather of the ICP transforms over NVLink peer memory (no NCCL collective on the data path).

    Every rank owns a ``[world * pairs_per_rank, 4, 4]`` buffer allocated as torch symmetric memory; the ICP kernel's
    epilogue stores each pair's 4x4 into row ``rank * pairs_per_rank + p`` of ALL ranks' buffers through the
    peer-mapped pointers, and ``finish()`` runs the cross-rank barrier that makes the rows visible.
    """

    def __init__(self, pairs_per_rank: int, device, group=None, slots: int = 2):
        import torch.distributed._symmetric_memory as symm

        self.group = dist.group.WORLD if group is None else group
        self.world = dist.get_world_size(self.group)
        self.rank = dist.get_rank(self.group)
        self.pairs = int(pairs_per_rank)
        self.slots = []
        for _ in range(slots):          # double-buffered so that a gather may overlap the next batch
            buf = symm.empty(self.world * self.pairs, 16, dtype=torch.float32, device=device)
            buf.zero_()
            hdl = symm.rendezvous(buf, self.group)
            self.slots.append((buf, hdl))

    def ext(self, slot: int = 0, start_event: int = 0, stop_event: int = 0):
        """The ``icp_batch(..., ext=...)`` argument that routes that call's transforms into slot ``slot`` of every rank
        (fused: stored by the ICP kernel's epilogue)."""
        from . import ops

        _buf, hdl = self.slots[slot]
        return ops.icp_ext(hdl.buffer_ptrs_dev, self.world, self.rank * self.pairs, start_event, stop_event)

    def push(self, local_pose: torch.Tensor, slot: int = 0):
        """Un-fused variant for transforms that are final only after a later kernel (``hist_icp``): one small launch
        copies this rank's ``[pairs,4,4]`` block into slot ``slot`` of every rank with 16-byte peer stores."""
        import ctypes
        from . import _lib, ops

        _buf, hdl = self.slots[slot]
        lp = local_pose.contiguous()
        with torch.cuda.device(lp.device):
            code = _lib.lib().icpf_peer_push_f32(ctypes.c_void_p(lp.data_ptr()), ctypes.c_void_p(hdl.buffer_ptrs_dev),
                                                 self.world, self.rank * self.pairs, lp.shape[0], ops._stream_ptr())
        _lib.check(code, "icpf_peer_push_f32")

    def finish(self, slot: int = 0) -> torch.Tensor:
        """Stream-ordered cross-rank barrier; afterwards the slot holds the transforms of every rank's pairs."""
        buf, hdl = self.slots[slot]
        hdl.barrier()
        return buf.view(self.world * self.pairs, 4, 4)
