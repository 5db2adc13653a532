"""Data-parallel plumbing for the hot path: one point cloud (or an equal slice of the batch) per
GPU, full weight replica per GPU, and ONE all-reduce of a single flat gradient bucket per step.

The reference has no distributed code at all (no NCCL / torch.distributed call
sites); every op on the path is per-sample (the batch index is just ``indices[:, 0]``,
``spconv/pytorch/core.py:148``), so there is no feature collective -- only the weight (and bias)
gradients cross NVLink.  For a SECOND-style encoder the bucket is O(1 MB): latency-bound, hence a
single flat bucket rather than per-tensor collectives.
"""
from __future__ import annotations

from typing import Iterable, List, Optional, Tuple

import torch
import torch.distributed as dist


def shard_batch(indices: torch.Tensor, features: torch.Tensor, batch_size: int, rank: int,
                world_size: int) -> Tuple[torch.Tensor, torch.Tensor, int]:
    """Samples ``b`` with ``b % world_size == rank`` go to ``rank``; batch ids are renumbered
    ``0..local_bs-1``.  Returns ``(indices, features, local_batch_size)``."""
    assert batch_size % world_size == 0, "batch must divide evenly across ranks"
    b = indices[:, 0].long()
    keep = (b % world_size) == rank
    local = indices[keep].clone()
    local[:, 0] = (b[keep] // world_size).to(indices.dtype)
    return local, features[keep], batch_size // world_size


class GradBucket:
    """Flat view over the gradients of ``params`` so a step needs exactly one collective."""

    def __init__(self, params: Iterable[torch.nn.Parameter]):
        self.params: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]
        assert self.params, "no trainable parameters"
        dev, dt = self.params[0].device, self.params[0].dtype
        self.numel = sum(p.numel() for p in self.params)
        self.flat = torch.zeros(self.numel, device=dev, dtype=dt)
        self._views: List[torch.Tensor] = []
        off = 0
        for p in self.params:
            assert p.dtype == dt, "one bucket per dtype"
            self._views.append(self.flat[off:off + p.numel()].view_as(p))
            off += p.numel()
        self.attach()

    def attach(self) -> int:
        """(Re-)alias every ``p.grad`` to its slice of the bucket.  ``optimizer.zero_grad()`` and
        ``module.zero_grad()`` default to ``set_to_none=True``, which drops the aliases; the next
        backward then allocates fresh ``.grad`` tensors OUTSIDE the bucket.  A gradient found
        outside is copied in before it is re-aliased, so nothing is lost.  Returns how many
        parameters had to be re-attached."""
        fixed = 0
        for p, view in zip(self.params, self._views):
            g = p.grad
            if g is not None and g.data_ptr() == view.data_ptr() and g.shape == view.shape:
                continue
            if g is not None:
                view.copy_(g)
            p.grad = view
            fixed += 1
        return fixed

    def zero(self) -> None:
        """Zero the gradients in place (use instead of ``zero_grad(set_to_none=True)``)."""
        self.attach()
        self.flat.zero_()

    def all_reduce(self, group: Optional[dist.ProcessGroup] = None, average: bool = True,
                   async_op: bool = False):
        """Sum (or mean) the bucket over the data-parallel group: one NCCL launch per step."""
        self.attach()          # gradients produced after a zero_grad(set_to_none=True) are pulled in
        if not dist.is_initialized() or dist.get_world_size(group) == 1:
            return None
        if average:
            self.flat.div_(dist.get_world_size(group))
        return dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group, async_op=async_op)


def allreduce_gradients(module: torch.nn.Module, group: Optional[dist.ProcessGroup] = None,
                        average: bool = True) -> None:
    """Stateless variant: flatten existing ``.grad`` tensors, one all-reduce, copy back."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    grads = [p.grad for p in module.parameters() if p.grad is not None]
    if not grads:
        return
    flat = torch.cat([g.reshape(-1) for g in grads])
    if average:
        flat.div_(dist.get_world_size(group))
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    off = 0
    for g in grads:
        g.copy_(flat[off:off + g.numel()].view_as(g))
        off += g.numel()
