import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "contexture-nerf_b200")]
import torch
from ctxnerf import ops
dev = torch.device("cuda", 0)
R, S = 1 << 19, 64
raw = torch.randn(R, S, 4, device=dev); raw[..., 3] *= 5
z = torch.sort(torch.rand(R, S, device=dev) * 4 + 2, -1)[0]
d = torch.randn(R, 3, device=dev)
w = torch.rand(R, S, device=dev)
pts = torch.rand(R * 16, 3, device=dev) * 4 - 2
for _ in range(2):
    ops.posenc(pts, 10)
    ops.composite(raw, z, d)
    rr = raw.clone().requires_grad_(True)
    outs = ops.composite(rr, z, d)
    torch.autograd.backward(list(outs), [torch.ones_like(o) for o in outs])
    ops.resample_merge(z, w, 128, det=True)
    with torch.no_grad():
        ops.resample(z[:, :63].contiguous(), w[:, :62].contiguous(), 128, det=True)
    ops.resample_merge(z, w, 128, det=False, seed=5)
    K = [[1111.1, 0, 400.0], [0, 1111.1, 400.0], [0, 0, 1]]
    ops.raygen(1600, 1600, K, torch.eye(4, device=dev)[:3].contiguous(), n_samples=64, near=2., far=6., perturb=True, seed=1, want_viewdirs=True)
torch.cuda.synchronize()
print("ok")
