"""Data-parallel plumbing for CiM training.

The hot path shards by batch: every rank runs the conv kernels on its own images; the only exchange is
the gradient all-reduce of the replicated parameters (weights + step sizes, ~1.2 MB for ResNet-20).  The
reference wraps the model in ``DistributedDataParallel`` (examples/__init__.py:711-712); here the same
arithmetic -- sum over ranks, divide by world size -- is one flat fp32 buffer and ONE ``all_reduce`` (NCCL
over NVLink on the GPU box, gloo in the CPU tests), which is what matters for a latency-bound payload.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


class FlatGradAllReducer:
    """Flattens the gradients of ``params`` into one preallocated buffer, all-reduces it once and scatters the
    averaged values back into ``p.grad`` (views, no extra copies on the way back)."""

    def __init__(self, params, group=None):
        self.params = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("no trainable parameters")
        dev, dt = self.params[0].device, self.params[0].dtype
        self.sizes = [p.numel() for p in self.params]
        self.flat = torch.zeros(sum(self.sizes), device=dev, dtype=dt)
        self.views = [v.view_as(p) for v, p in zip(self.flat.split(self.sizes), self.params)]
        self.group = group

    @property
    def nbytes(self) -> int:
        return self.flat.numel() * self.flat.element_size()

    def all_reduce_(self, average: bool = True):
        """In place: p.grad <- mean over ranks of p.grad (DDP semantics).  Parameters without a gradient
        contribute zeros."""
        # one multi-tensor copy instead of a launch per parameter (about 80 for ResNet-20)
        src = [p.grad if p.grad is not None else torch.zeros_like(v) for v, p in zip(self.views, self.params)]
        todo = [(v, g) for v, g in zip(self.views, src) if g.data_ptr() != v.data_ptr()]  # already a view: in place
        if todo:
            torch._foreach_copy_([v for v, _ in todo], [g for _, g in todo])
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(self.group) > 1:
            dist.all_reduce(self.flat, group=self.group)
            if average:
                self.flat.div_(dist.get_world_size(self.group))
        for v, p in zip(self.views, self.params):
            p.grad = v
        return self.flat


def broadcast_parameters(module: torch.nn.Module, src: int = 0, group=None):
    """Rank ``src``'s parameters and buffers to every rank (DDP's constructor does the same), so that the
    data-dependent step-size initialisation of rank 0's first batch is shared."""
    if not (dist.is_available() and dist.is_initialized()):
        return
    for t in list(module.parameters()) + list(module.buffers()):
        dist.broadcast(t.data, src, group=group)


def shard_batch(x: torch.Tensor, rank: int, world: int) -> torch.Tensor:
    """Contiguous, equal split of a global batch (the reference uses DistributedSampler)."""
    if x.shape[0] % world:
        raise ValueError(f"global batch {x.shape[0]} is not divisible by world size {world}")
    n = x.shape[0] // world
    return x[rank * n:(rank + 1) * n]


# ---- one-off all-reduce of the data-dependent initialisation statistics -----------------
# The reference initialises step sizes from each rank's own first batch *after* DDP has broadcast the
# parameters, so replicas start from different step sizes (H11).  With more than one rank the modules reduce
# the three statistics below over all ranks, so every replica starts from the value a single process would
# compute on the global batch (equal shards).  Single-process behaviour is untouched.
_sync_lazy_init = True


def set_sync_lazy_init(enabled: bool) -> None:
    """False restores the reference's per-rank initialisation."""
    global _sync_lazy_init
    _sync_lazy_init = bool(enabled)


def _world(group=None) -> int:
    if _sync_lazy_init and dist.is_available() and dist.is_initialized():
        return dist.get_world_size(group)
    return 1


def global_mean_(t: torch.Tensor, group=None) -> torch.Tensor:
    """In place: mean over ranks of a per-rank mean (equal shard sizes)."""
    w = _world(group)
    if w > 1:
        dist.all_reduce(t, group=group)
        t.div_(w)
    return t


def global_min_(t: torch.Tensor, group=None) -> torch.Tensor:
    if _world(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MIN, group=group)
    return t


def global_sum_(t: torch.Tensor, group=None):
    """In place sum over ranks; returns (tensor, number of ranks that contributed)."""
    w = _world(group)
    if w > 1:
        dist.all_reduce(t, group=group)
    return t, w
