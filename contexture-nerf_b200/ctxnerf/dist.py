"""Ray-sharded data parallelism: one process per GPU, weights replicated and
resident, ONE all-reduce of a flat fp32 gradient bucket per step.

Replaces the reference's only parallel construct, nn.DataParallel around the
MLP (/root/reference/src/training/trainer.py:134-135: per-forward weight
broadcast + input scatter + output gather + reduce-add onto GPU 0).  The path
shards by rays with no data-path collective; the gradient all-reduce goes
through torch.distributed (NCCL over NVLink on the GPU box, gloo in CPU tests).
"""
from __future__ import annotations

from typing import List, Sequence

import torch
import torch.distributed as dist


class FlatBucket:
    """Re-homes the parameters of ``modules`` into one flat fp32 tensor and gives
    every parameter a ``.grad`` that is a view into one flat gradient tensor
    (module/parameter order preserved), so that the wgrad kernels write the
    bucket in place and one collective covers both networks."""

    def __init__(self, modules: Sequence[torch.nn.Module]):
        self.params: List[torch.nn.Parameter] = [p for m in modules for p in m.parameters()]
        dev = self.params[0].device
        n = sum(p.numel() for p in self.params)
        self.flat = torch.empty(n, device=dev, dtype=torch.float32)
        self.grad = torch.zeros(n, device=dev, dtype=torch.float32)
        off = 0
        with torch.no_grad():
            for p in self.params:
                k = p.numel()
                self.flat[off:off + k].copy_(p.detach().reshape(-1))
                p.data = self.flat[off:off + k].view_as(p)
                p.grad = self.grad[off:off + k].view_as(p)
                off += k
        self.numel = n

    def sinks_for(self, module) -> List[torch.Tensor]:
        """Gradient views in the order of ``module._param_list()``."""
        return [p.grad for p in module._param_list()]

    def zero_grad(self):
        self.grad.zero_()

    def all_reduce(self, group=None, async_op=False):
        """Sum-reduce the bucket over ranks (no-op for a single process)."""
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            return dist.all_reduce(self.grad, op=dist.ReduceOp.SUM, group=group, async_op=async_op)
        return None


def shard_rays(n_total: int, rank: int, world_size: int):
    """Contiguous ray range [lo, hi) of this rank (inference: row blocks of the
    image; any remainder goes to the first ranks)."""
    base, rem = divmod(n_total, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def rank_generator(seed: int, rank: int, device="cpu") -> torch.Generator:
    """Per-rank RNG stream (seed + rank) for ray selection."""
    g = torch.Generator(device=device)
    g.manual_seed(seed + rank)
    return g


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1
