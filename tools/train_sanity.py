"""A few hundred training steps on a fixed synthetic target: the loss must fall (exercises the overlapped backward,
Adam and weight re-packing repeatedly); prints the loss every 50 steps."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "contexture-nerf_b200")]
import torch
from ctxnerf.train import NerfTrainer
from ctxnerf.workloads import orbit_camera
dev = torch.device("cuda:0")
H = W = 64
K, c2w = orbit_camera(H, W, focal=80.0)
torch.manual_seed(0)
tr = NerfTrainer(H, W, K, c2w, N_samples=64, N_importance=128, perturb=1.0, device=dev, seed=0, lr=5e-4)
yy, xx = torch.meshgrid(torch.linspace(0, 1, H), torch.linspace(0, 1, W), indexing="ij")
target = torch.stack([xx, yy, 0.5 + 0.5 * torch.sin(6.28 * xx) * torch.cos(6.28 * yy)], -1).reshape(-1, 3).to(dev)
idx = torch.arange(H * W, device=dev)
first = last = None
for step in range(int(sys.argv[1]) if len(sys.argv) > 1 else 300):
    loss = tr.step(idx, target)
    if step % 50 == 0 or step == 299:
        l = loss.item()
        first = l if first is None else first
        last = l
        print(step, round(l, 5), flush=True)
assert last < 0.5 * first, (first, last)
print("ok: loss", first, "->", last)
