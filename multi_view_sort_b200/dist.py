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
    """Gradient averaging over the ranks.

    Generic path: one flat fp32 buffer holding every live parameter gradient, packed with one fused copy and all-reduced
    in one call.  `attach(nem, head)` removes the packing: each module's hand-written backward then writes its gradients
    straight into the module's own flat arena (functional.GradArena), the parameters' .grad are views of it, and the
    arenas are all-reduced in place - the head's (118 MB) on a side stream as soon as the head's backward is enqueued,
    i.e. under the whole EM backward pass; the NEM arena (516 MB) when allreduce_mean() is called."""

    def __init__(self, params: Iterable[torch.nn.Parameter]):
        self.params: List[torch.nn.Parameter] = [p for p in params]
        self.flat = None
        self.arenas = []            # (module, GradArena)
        self._side = None
        self._pending = []          # (arena, event recorded on the side stream after its collective)

    # ---- arena path -------------------------------------------------------------------------------------------
    def attach(self, nem, head=None, overlap_head=True):
        from .functional import GradArena, em_param_tensors, head_param_tensors
        for module, plist in ((nem, em_param_tensors(nem)), (head, head_param_tensors(head) if head is not None else None)):
            if module is None:
                continue
            arena = GradArena(plist)
            module.__dict__["_vmm_grad_arena"] = arena
            self.arenas.append((module, arena))
        if head is not None and overlap_head:
            head.__dict__["_vmm_after_backward"] = lambda: self._early_allreduce(head)
        return self

    def _world(self, group=None):
        return dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1

    def _reduce(self, flat, group=None):
        world = self._world(group)
        if world > 1:
            if dist.get_backend(group) == "nccl":
                dist.all_reduce(flat, op=dist.ReduceOp.AVG, group=group)
            else:
                dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
                flat.div_(world)

    def _early_allreduce(self, module):
        """Called when `module`'s backward has been enqueued: all-reduce its arena on a side stream."""
        arena = module.__dict__.get("_vmm_grad_arena")
        if arena is None or arena.flat is None or self._world() <= 1 or not arena.flat.is_cuda:
            return
        if not all(p.grad is not None and p.grad.data_ptr() == v.data_ptr() for p, v in zip(arena.params, arena.views_nozero())):
            return                                                  # the arena was not used for this backward
        if self._side is None:
            self._side = torch.cuda.Stream(device=arena.flat.device)
        self._side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(self._side):
            self._reduce(arena.flat)
            ev = torch.cuda.Event()
            ev.record(self._side)
        self._pending.append((arena, ev))

    def _arena_of(self, p):
        for _, arena in self.arenas:
            if arena.flat is not None and p.grad is not None and p.grad.is_cuda == arena.flat.is_cuda:
                lo = arena.flat.data_ptr()
                if lo <= p.grad.data_ptr() < lo + arena.flat.numel() * 4:
                    return arena
        return None

    def _live(self):
        # parameters whose grad is None (after_rnn.*: never used by the reference) stay None
        return [p for p in self.params if p.grad is not None]

    def allreduce_mean(self, group=None) -> int:
        """Average gradients over the group; returns the number of bytes exchanged.  Gradients living in an attached
        arena are reduced in place; everything else is packed into one flat buffer (one fused copy), reduced, and the
        parameters' .grad are re-pointed at views of it (no copy back)."""
        live = self._live()
        if not live:
            return 0
        done = {id(a) for a, _ in self._pending}
        for _, ev in self._pending:
            torch.cuda.current_stream().wait_event(ev)
        self._pending = []
        nbytes = 0
        in_arena, rest = {}, []
        for p in live:
            a = self._arena_of(p)
            if a is None:
                rest.append(p)
            else:
                in_arena[id(a)] = a
        for a in in_arena.values():
            nbytes += a.flat.numel() * 4
            if id(a) not in done:
                self._reduce(a.flat, group)
        if rest:
            n = sum(p.grad.numel() for p in rest)
            dev = rest[0].grad.device
            if self.flat is None or self.flat.numel() != n or self.flat.device != dev:
                self.flat = torch.empty(n, dtype=torch.float32, device=dev)
            views, off = [], 0
            for p in rest:
                k = p.grad.numel()
                views.append(self.flat[off:off + k].view_as(p.grad))
                off += k
            torch._foreach_copy_(views, [p.grad for p in rest])
            self._reduce(self.flat, group)
            for p, v in zip(rest, views):
                p.grad = v
            nbytes += n * 4
        return nbytes
