"""Multi-GPU plumbing for the VMM path: object sharding and gradient averaging.

The reference is data-parallel through a single-process `torch.nn.DataParallel` that
replicates the modules on EVERY one of the T `nem(...)` calls of a step and pads the batch by
repeating trailing samples (tools/Trainer.py:26-32, 142-154).  Here: one process per GPU
(torchrun), objects are independent in the forward pass so each rank runs the fused EM on its
contiguous shard with no data-path collective, and the only exchange is ONE all-reduce of the
flat parameter-gradient bucket per optimizer step (NCCL over NVLink; gloo in the CPU tests).

Every per-batch mean in the loss (B*M*D, B*K, B: tools/NEM_utilities.py:88,119; CE) is a mean over
objects, so with equal shards the full-batch gradient is the average of the rank gradients; with
ragged shards each rank weights its loss by B_local/B_global first (`loss_weight`).
"""
from __future__ import annotations

from typing import Iterable, List, Tuple

import torch
import torch.distributed as dist


def shard_range(n_objects: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [begin, end) of rank's objects; the first n % world ranks get one extra (no padding:
    the reference's repeat-the-last-samples padding changes the loss, Trainer.py:142-154)."""
    base, rem = divmod(n_objects, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def loss_weight(n_local: int, n_global: int, world: int) -> float:
    """Factor for a rank's (mean-over-local-objects) loss so that AVERAGING gradients over ranks
    reproduces the full-batch mean: world * n_local / n_global (1.0 for equal shards)."""
    return world * n_local / n_global


class GradBucket:
    """One flat fp32 buffer holding every live parameter gradient, all-reduced in one call."""

    def __init__(self, params: Iterable[torch.nn.Parameter]):
        self.params: List[torch.nn.Parameter] = [p for p in params]
        self.flat = None

    def _live(self):
        # parameters whose grad is None (after_rnn.*: never used by the reference) stay None
        return [p for p in self.params if p.grad is not None]

    def allreduce_mean(self, group=None) -> int:
        """Average gradients over the group; returns the bucket size in bytes.  One fused copy packs the gradients
        into the flat buffer, one collective averages it (NCCL's AVG; SUM + scale on backends without it), and the
        parameters' .grad are re-pointed at views of the flat buffer (no copy back)."""
        live = self._live()
        if not live:
            return 0
        n = sum(p.grad.numel() for p in live)
        dev = live[0].grad.device
        if self.flat is None or self.flat.numel() != n or self.flat.device != dev:
            self.flat = torch.empty(n, dtype=torch.float32, device=dev)
        views, off = [], 0
        for p in live:
            k = p.grad.numel()
            views.append(self.flat[off:off + k].view_as(p.grad))
            off += k
        torch._foreach_copy_(views, [p.grad for p in live])
        world = dist.get_world_size(group) if dist.is_initialized() else 1
        if world > 1:
            if dist.get_backend(group) == "nccl":
                dist.all_reduce(self.flat, op=dist.ReduceOp.AVG, group=group)
            else:
                dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)
                self.flat.div_(world)
        for p, v in zip(live, views):
            p.grad = v
        return n * 4
