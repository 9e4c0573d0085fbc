"""View-weight masks of the texture trainer: drop-ins for ``ConTEXTure.create_face_view_map`` and
``ConTEXTure.compare_face_normals_between_views`` (/root/reference/src/training/trainer.py:155-249), which the
reference builds on torch_scatter's CUDA ``scatter_max`` (:227, absent from this image).

Use inside the reference::

    from ctxnerf import view_weights as vw
    ConTEXTure.create_face_view_map = lambda self, face_idx: vw.create_face_view_map(face_idx)
    ConTEXTure.compare_face_normals_between_views = \\
        lambda self, fvm, face_normals, face_idx: vw.compare_face_normals_between_views(fvm, face_normals, face_idx)

``view_weight_masks(face_normals, face_idx)`` is the fused form: no [N,4] row table is materialised (three
HBM-bound passes over face_idx, csrc/viewmask.cu).  No CPU fallback: CUDA tensors only.
"""
from __future__ import annotations

import torch

from . import _lib
from ._lib import call, ptr, stream_ptr


def _check(face_idx: torch.Tensor):
    if not (torch.is_tensor(face_idx) and face_idx.is_cuda):
        raise _lib.CtxNerfError("ctxnerf view-weight masks run on CUDA tensors only (no CPU fallback)")
    if face_idx.dim() != 4 or face_idx.shape[1] != 1:
        raise _lib.CtxNerfError("face_idx must be [num_views, 1, H, W] (the rasteriser's face index image)")
    return face_idx.to(torch.int64).contiguous()


def create_face_view_map(face_idx: torch.Tensor) -> torch.Tensor:
    """[V,1,H,W] face ids (< 0 = background) -> int64 rows (face, view, i, j) of the covered pixels, in (view, pixel)
    order -- the tensor the reference builds with meshgrid + stack + boolean filtering (trainer.py:155-216)."""
    fi = _check(face_idx)
    V, _, H, W = fi.shape
    dev = fi.device
    nblk = int(_lib.lib().ctx_face_view_map_blocks(V * H * W))
    counts = torch.empty(nblk + 1, dtype=torch.int64, device=dev)
    with torch.cuda.device(dev):
        call("ctx_face_view_map", ptr(fi), V, H, W, ptr(counts), None, 0, stream_ptr(dev))
        n = int(counts[nblk].item())                     # the row count decides the size of the result
        rows = torch.empty(n, 4, dtype=torch.int64, device=dev)
        if n:
            call("ctx_face_view_map", ptr(fi), V, H, W, ptr(counts), ptr(rows), 1, stream_ptr(dev))
    return rows


def view_weight_masks(face_normals: torch.Tensor, face_idx: torch.Tensor) -> torch.Tensor:
    """face_normals [V,3,F], face_idx [V,1,H,W] -> bool weight masks [V,1,H,W]: a covered pixel keeps True iff its
    face's z-normal in this view is not below the face's maximum over every view that shows it."""
    fi = _check(face_idx)
    if not face_normals.is_cuda:
        raise _lib.CtxNerfError("ctxnerf view-weight masks run on CUDA tensors only (no CPU fallback)")
    V, _, H, W = fi.shape
    fn = face_normals.to(torch.float32).contiguous()
    if fn.dim() != 3 or fn.shape[0] != V or fn.shape[1] != 3:
        raise _lib.CtxNerfError("face_normals must be [num_views, 3, num_faces]")
    F = fn.shape[2]
    dev = fi.device
    visible = torch.empty(max(V * F, 1), dtype=torch.uint8, device=dev)
    maxz = torch.empty(max(F, 1), dtype=torch.float32, device=dev)
    mask = torch.empty(V, 1, H, W, dtype=torch.bool, device=dev)
    with torch.cuda.device(dev):
        call("ctx_view_weight_masks", ptr(fn), ptr(fi), V, F, H, W, ptr(visible), ptr(maxz), ptr(mask),
             stream_ptr(dev))
    return mask


def compare_face_normals_between_views(face_view_map, face_normals, face_idx):
    """Signature of the reference method (trainer.py:218).  The row table carries nothing face_idx does not, so the
    fused kernels work from face_idx directly; ``face_view_map`` is accepted for call compatibility."""
    return view_weight_masks(face_normals, face_idx)
