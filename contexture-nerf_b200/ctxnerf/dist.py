"""Ray-sharded data parallelism: one process per GPU, weights replicated and
resident, ONE all-reduce of a flat fp32 gradient bucket per step.

Replaces the reference's only parallel construct, nn.DataParallel around the
MLP (/root/reference/src/training/trainer.py:134-135: per-forward weight
broadcast + input scatter + output gather + reduce-add onto GPU 0).  The path
shards by rays with no data-path collective.  The gradient all-reduce is NCCL
over NVLink: on the GPU through the library's own binding (``BucketComm`` ->
ctx_allreduce of include/ctxnerf.h, enqueued on the step's stream so that the
whole step -- all-reduce included -- is ONE captured CUDA graph), else through
torch.distributed (gloo in the CPU tests, or CTXNERF_NCCL=0).
"""
from __future__ import annotations

from typing import List, Sequence

import ctypes
import os

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


def _loaded_nccl_path():
    """Path of the NCCL shared object already mapped into this process (torch's bundled copy), so that the C-ABI
    binding and torch.distributed use the same library; None -> the binding falls back to the soname."""
    try:
        with open("/proc/self/maps") as fh:
            for line in fh:
                if "libnccl.so" in line:
                    return line.split()[-1]
    except OSError:
        pass
    return None


class BucketComm:
    """NCCL communicator of the C-ABI (ctx_comm_* / ctx_allreduce, include/ctxnerf.h) over the ranks of the
    initialised torch.distributed group: rank 0 makes the rendezvous token, torch.distributed carries its 128 bytes
    to the other ranks, every rank then joins on its own device.  ``all_reduce`` enqueues on the CURRENT stream."""

    def __init__(self, device, group=None):
        from . import _lib
        self._lib = _lib
        rank, world_size = world()
        if world_size < 2:
            raise RuntimeError("BucketComm needs an initialised torch.distributed group with more than one rank")
        path = os.environ.get("CTXNERF_NCCL_LIB") or _loaded_nccl_path()
        on_gpu = "nccl" in str(dist.get_backend(group))      # ("nccl", or the combined "cpu:gloo,cuda:nccl" default)
        token = torch.zeros(128, dtype=torch.uint8)
        local_error = None
        try:                                   # local part: resolve NCCL, make the token
            _lib.call("ctx_comm_load", path.encode() if path else None)
            if rank == 0:
                _lib.call("ctx_comm_unique_id", ctypes.c_void_p(token.data_ptr()))
        except Exception as e:
            local_error = e
        # every rank learns whether every rank got this far BEFORE anyone enters the broadcast / the NCCL rendezvous
        ok = torch.tensor([0 if local_error else 1], dtype=torch.int32, device=device if on_gpu else "cpu")
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
        if int(ok.item()) == 0:
            raise _lib.CtxNerfError(f"BucketComm: NCCL binding unavailable on at least one rank ({local_error})")
        carrier = token.to(device) if on_gpu else token
        dist.broadcast(carrier, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        token = carrier.cpu().contiguous()
        handle = ctypes.c_void_p()
        with torch.cuda.device(device):
            _lib.call("ctx_comm_init", ctypes.c_void_p(ctypes.addressof(handle)), world_size, ctypes.c_void_p(token.data_ptr()), rank)
        self.handle, self.device, self.world_size = handle, torch.device(device), world_size
        self.version = _lib.lib().ctx_comm_version()

    def all_reduce(self, t: torch.Tensor):
        """In-place sum of a contiguous fp32 CUDA tensor over the ranks, on the current stream of its device."""
        if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous() and t.device == self.device):
            raise self._lib.CtxNerfError("BucketComm.all_reduce: contiguous fp32 tensor on the communicator's device")
        self._lib.call("ctx_allreduce", self.handle, self._lib.ptr(t), t.numel(), self._lib.stream_ptr(self.device))

    def close(self):
        if self.handle is not None and self.handle.value:
            torch.cuda.synchronize(self.device)
            self._lib.call("ctx_comm_destroy", self.handle)
        self.handle = None


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


def dist_backend():
    """Backend name of the default process group ("" when there is none)."""
    if dist.is_available() and dist.is_initialized():
        return str(dist.get_backend())
    return ""


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


# ---------------------------------------------------------------------------------------------------------
# Multi-GPU inference: rays are independent, so inference shards with NO data-path collective --
# BASELINE config 2 by contiguous row blocks of one image, config 4 by whole views -- and an optional gather of the
# finished maps to rank 0.  One process per GPU (torchrun); with a single process both functions degenerate to the
# plain render.

def _gather_rows(t: torch.Tensor, counts, dst: int = 0, group=None):
    """Gather row blocks of different lengths to ``dst`` (rows padded to the longest block for the collective)."""
    rank, world_size = world()
    n_max = max(counts)
    pad = torch.zeros((n_max,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    pad[:t.shape[0]] = t
    bufs = [torch.empty_like(pad) for _ in range(world_size)] if rank == dst else None
    dist.gather(pad, bufs, dst=dst, group=group)
    if rank != dst:
        return None
    return torch.cat([b[:c] for b, c in zip(bufs, counts)], 0)


def render_image_sharded(trainer, gather: bool = True, group=None):
    """BASELINE config 2 on N GPUs: this rank renders the contiguous block of image rows ``shard_rays`` gives it
    (coarse + fine, deterministic) -- ``trainer.render`` on its pixel range.  Returns the dict of maps of the local
    rows plus ``"rows": (lo, hi)`` (pixel range); with ``gather`` rank 0 instead gets the full [H, W, ...] maps (other
    ranks get their local block)."""
    rank, world_size = world()
    n = trainer.H * trainer.W
    lo, hi = shard_rays(n, rank, world_size)
    idx = torch.arange(lo, hi, device=trainer.device, dtype=torch.int64)
    out = trainer.render(idx) if hi > lo else None
    if out is None:       # more ranks than pixels
        dev = trainer.device
        out = dict(rgb_map=torch.empty(0, 3, device=dev), disp_map=torch.empty(0, device=dev),
                   acc_map=torch.empty(0, device=dev), depth_map=torch.empty(0, device=dev),
                   rgb0=torch.empty(0, 3, device=dev))
    out["rows"] = (lo, hi)
    if not gather or world_size == 1:
        if world_size == 1:
            out = {k: (v.reshape(trainer.H, trainer.W, *v.shape[1:]) if torch.is_tensor(v) else v) for k, v in out.items()}
        return out
    counts = [b - a for a, b in (shard_rays(n, r, world_size) for r in range(world_size))]
    full = {}
    for k in ("rgb_map", "disp_map", "acc_map", "depth_map", "rgb0"):
        g = _gather_rows(out[k], counts, 0, group)
        if rank == 0:
            full[k] = g.reshape(trainer.H, trainer.W, *g.shape[1:])
    if rank == 0:
        full["rows"] = (0, n)
        return full
    return out


def render_views_sharded(trainer, cameras, H: int, W: int, n_samples: int = 192, sphere=None, gather: bool = True,
                         group=None):
    """BASELINE config 4 on N GPUs: view v is rendered by rank v % N (``trainer.render_view``: single pass of the fine
    network, per-ray near/far from the bounding sphere).  ``cameras`` = [(K, c2w)].  Returns {view index: maps} for the
    views of this rank; with ``gather`` rank 0 gets every view's rgb / acc / depth maps (one gather per round of N
    views, [H,W,5] per rank)."""
    rank, world_size = world()
    mine = {}
    for v, (K, c2w) in enumerate(cameras):
        if v % world_size == rank:
            mine[v] = trainer.render_view(H, W, K, c2w, n_samples=n_samples, sphere=sphere)
    if not gather or world_size == 1:
        return mine
    dev = trainer.device
    out = {}
    for base in range(0, len(cameras), world_size):
        v = base + rank
        if v in mine:
            m = mine[v]
            pack = torch.cat([m["rgb_map"], m["acc_map"][..., None], m["depth_map"][..., None]], -1).contiguous()
        else:
            pack = torch.zeros(H, W, 5, device=dev)
        bufs = [torch.empty_like(pack) for _ in range(world_size)] if rank == 0 else None
        dist.gather(pack, bufs, dst=0, group=group)
        if rank == 0:
            for r, b in enumerate(bufs):
                if base + r < len(cameras):
                    out[base + r] = dict(rgb_map=b[..., :3], acc_map=b[..., 3], depth_map=b[..., 4])
    return out if rank == 0 else mine
