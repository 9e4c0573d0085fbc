"""Host side of the MLP backward: allocates the gradient tensors and the dZ
scratch records and calls ctx_mlp_bwd (dgrad + wgrad kernels)."""
from __future__ import annotations

import ctypes

import torch

from ._lib import call, ptr, stream_ptr


def mlp_backward(module, packed, acts, P, g_out):
    desc = module._desc
    w, wt, f = packed
    params = module._param_list()
    dev = g_out.device
    g = g_out.reshape(P, desc.out_ch).float().contiguous()
    # one flat zeroed bucket; the per-parameter gradients are views into it
    sizes = [p.numel() for p in params]
    flat = torch.zeros(sum(sizes), device=dev, dtype=torch.float32)
    grads, off = [], 0
    for p, n in zip(params, sizes):
        grads.append(flat[off:off + n].view_as(p))
        off += n
    dacts = torch.empty_like(acts)
    arr = (ctypes.c_void_p * len(grads))(*[t.data_ptr() for t in grads])
    with torch.cuda.device(dev):
        call("ctx_mlp_bwd", desc.p, ptr(wt), ptr(f), ptr(g), ptr(acts), ptr(dacts), P,
             ctypes.cast(arr, ctypes.c_void_p), len(grads), stream_ptr(dev))
    return [gr if p.requires_grad else None for gr, p in zip(grads, params)]
