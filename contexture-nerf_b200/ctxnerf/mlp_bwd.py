"""Host side of the MLP backward: allocates the gradient tensors and the dZ
scratch records and calls ctx_mlp_bwd (dgrad + wgrad kernels)."""
from __future__ import annotations

import ctypes

import torch

from ._lib import call, ptr, stream_ptr


def mlp_backward(module, packed, acts, P, g_out, sinks=None, timed=None):
    """Gradients w.r.t. the module's parameters.

    ``sinks`` (list of fp32 tensors, one per parameter, e.g. views into a flat
    all-reduce bucket): the wgrad kernel accumulates straight into them and
    ``None`` is returned; otherwise fresh zeroed tensors are returned.
    ``timed(name, fn)`` (optional) wraps the two kernel launches ("dgrad", "wgrad")."""
    desc = module._desc
    w, wt, f = packed
    params = module._param_list()
    dev = g_out.device
    g = g_out.reshape(P, desc.out_ch).float().contiguous()
    if sinks is None:
        sizes = [p.numel() for p in params]
        flat = torch.zeros(sum(sizes), device=dev, dtype=torch.float32)
        grads, off = [], 0
        for p, n in zip(params, sizes):
            grads.append(flat[off:off + n].view_as(p))
            off += n
    else:
        grads = sinks
    dacts = torch.empty_like(acts)
    arr = (ctypes.c_void_p * len(grads))(*[t.data_ptr() for t in grads])
    run = timed if timed is not None else (lambda name, fn: fn())
    with torch.cuda.device(dev):
        run("dgrad", lambda: call("ctx_mlp_dgrad", desc.p, ptr(wt), ptr(f), ptr(g), ptr(acts), ptr(dacts), P,
                                  stream_ptr(dev)))
        run("wgrad", lambda: call("ctx_mlp_wgrad", desc.p, ptr(acts), ptr(dacts), P,
                                  ctypes.cast(arr, ctypes.c_void_p), len(grads), stream_ptr(dev)))
    if sinks is not None:
        return None
    return [gr if p.requires_grad else None for gr, p in zip(grads, params)]
