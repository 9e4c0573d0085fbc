"""Golden vectors for the view-weight masks from the UNMODIFIED reference methods.

Run in the build container only (needs /root/reference):  python tests/golden/make_golden_view_weights.py

`ConTEXTure.create_face_view_map` / `compare_face_normals_between_views`
(/root/reference/src/training/trainer.py:155-249) live in a module that imports kaolin, diffusers, ... (absent here),
so the two method definitions are taken out of the source text as they stand (ast, no edits) and executed with
`torch` and ONE injected name: `scatter_max`, the third-party torch_scatter function (:227, not installed, no version
pinned).  The stand-in follows its documented contract -- per-index maximum of the source rows, output length
max(index)+1 -- via Tensor.scatter_reduce_('amax'); that one call is therefore not pinned by the reference, the rest of
both methods is.  Output: tests/golden/ref_view_weights.npz.
"""
import ast
import os

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/src/training/trainer.py"


def scatter_max(src, index, dim=0):
    n = int(index.max().item()) + 1
    out = torch.full((n,), float("-inf"), dtype=src.dtype).scatter_reduce_(0, index, src, "amax", include_self=False)
    return out, None


def load_methods():
    tree = ast.parse(open(SRC).read())
    cls = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "ConTEXTure")
    want = {"create_face_view_map", "compare_face_normals_between_views"}
    fns = [n for n in cls.body if isinstance(n, ast.FunctionDef) and n.name in want]
    mod = ast.Module(body=fns, type_ignores=[])
    ns = {"torch": torch, "scatter_max": scatter_max}
    exec(compile(mod, SRC, "exec"), ns)
    return ns["create_face_view_map"], ns["compare_face_normals_between_views"]


def main():
    create, compare = load_methods()
    g = torch.Generator().manual_seed(77)
    V, H, W, F = 4, 24, 20, 60
    ids = torch.randint(0, F, (V, 1, H, W // 4), generator=g).repeat_interleave(4, dim=3)
    face_idx = torch.where(torch.rand(V, 1, H, W, generator=g) < 0.35, torch.full_like(ids, -1), ids)
    normals = torch.randn(V, 3, F, generator=g)
    normals[:, 2, 5] = 0.5                       # a tie between all views
    rows = create(None, face_idx)
    masks = compare(None, rows, normals, face_idx)
    np.savez_compressed(os.path.join(HERE, "ref_view_weights.npz"), face_idx=face_idx.numpy(), normals=normals.numpy(),
                        rows=rows.numpy(), masks=masks.numpy())
    print("rows", tuple(rows.shape), "masks true fraction", masks.float().mean().item())


if __name__ == "__main__":
    main()
