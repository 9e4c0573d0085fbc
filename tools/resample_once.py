import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "contexture-nerf_b200")]
import torch
from ctxnerf import ops
dev = torch.device("cuda:0")
R, S = 1 << 18, 64
z = torch.sort(torch.rand(R, S, device=dev) * 4 + 2, -1)[0]; w = torch.rand(R, S, device=dev)
bins = z[:, :S - 1].contiguous(); wc = w[:, :S - 2].contiguous()
with torch.no_grad():
    for _ in range(2):
        ops.resample(bins, wc, 128, det=True)
        ops.resample_merge(z, w, 128, det=False, seed=3)
torch.cuda.synchronize()
print("ok")
