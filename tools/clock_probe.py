"""Effective SM clock under the training step: torch.cuda._sleep(N cycles) timed right after a burst of steps."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "contexture-nerf_b200")]
import torch
from ctxnerf.train import NerfTrainer
from ctxnerf.workloads import orbit_camera
dev = torch.device("cuda", 0)
K, c2w = orbit_camera()
tr = NerfTrainer(800, 800, K, c2w, device=dev, seed=0)
idx = torch.randint(0, 640000, (4096,), device=dev); tgt = torch.rand(4096, 3, device=dev)
def probe(tag):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); torch.cuda._sleep(20_000_000); b.record(); torch.cuda.synchronize()
    print(f"{tag}: {20_000_000 / (a.elapsed_time(b) * 1e-3) / 1e6:.0f} MHz effective", flush=True)
probe("cold")
for n in (5, 50, 200, 500):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): tr.step(idx, tgt)
    b.record()
    torch.cuda._sleep(20_000_000)
    c = torch.cuda.Event(enable_timing=True); c.record(); torch.cuda.synchronize()
    print(f"after {n} steps: {a.elapsed_time(b) / n:.3f} ms/step; sleep probe {20_000_000 / (b.elapsed_time(c) * 1e-3) / 1e6:.0f} MHz", flush=True)
    os.system("nvidia-smi --query-gpu=clocks.sm,power.draw,clocks_event_reasons.sw_power_cap --format=csv,noheader")
