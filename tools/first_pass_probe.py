"""Does the first timed pass of a process differ from later ones?  20-step passes of the captured training step, with
different pauses between them (the bench's value pass is the first pass of its process)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "contexture-nerf_b200")]
import torch
from ctxnerf.train import NerfTrainer
from ctxnerf.workloads import orbit_camera
dev = torch.device("cuda", 0)
K, c2w = orbit_camera()
tr = NerfTrainer(800, 800, K, c2w, perturb=1.0, white_bkgd=True, device=dev, seed=0)
idx = [torch.randint(0, 640000, (4096,), device=dev) for _ in range(4)]
tgt = [torch.rand(4096, 3, device=dev) for _ in range(4)]
def run(n):
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(n): tr.step(idx[i % 4], tgt[i % 4])
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
for i in range(5): tr.step(idx[i % 4], tgt[i % 4])      # the bench's warm-up: eager step, capture, replays
torch.cuda.synchronize()
out = []
for pause in (0.0, 0.0, 0.5, 0.5, 2.0, 2.0, 0.0, 5.0, 0.0):
    time.sleep(pause)
    out.append((pause, round(run(20), 4)))
print("pause before pass [s], ms/step over 20 steps:", out)
