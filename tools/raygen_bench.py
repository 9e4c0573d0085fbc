"""get_rays + stratified depths at 2^20-2^21 rays as a fraction of the measured HBM copy bandwidth (24 + 12 + 4 S B per ray
with view directions)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "contexture-nerf_b200")]
import torch
from ctxnerf import ops
dev = torch.device("cuda:0")
HBM = 6540.2e9
def timeit(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); [fn() for _ in range(n)]; b.record(); torch.cuda.synchronize(); return a.elapsed_time(b) / n * 1e-3
K = [[1111.1, 0, 512.0], [0, 1111.1, 512.0], [0, 0, 1]]
c2w = torch.eye(4, device=dev)[:3].contiguous()
H = W = 1024
for S in (32, 64, 128, 192):
    R = H * W
    outs = (torch.empty(R, 3, device=dev), torch.empty(R, 3, device=dev), torch.empty(R, 3, device=dev), torch.empty(R, S, device=dev))
    res = []
    for perturb in (False, True):
        t = timeit(lambda: ops.raygen(H, W, K, c2w, n_samples=S, near=2., far=6., perturb=perturb, seed=1, want_viewdirs=True, out=outs))
        res.append((round(R * (36 + 4 * S) / t / HBM, 3), round(t * 1e3, 3)))
    idx = torch.randint(0, R, (4096,), device=dev)
    o4 = tuple(t[:4096] for t in outs)
    t4 = timeit(lambda: ops.raygen(H, W, K, c2w, ray_idx=idx, n_samples=S, near=2., far=6., perturb=True, seed=1, want_viewdirs=True, out=o4), 50)
    print(f"S={S}: linspace depths {res[0]}, jittered {res[1]} (fraction of copy bandwidth, ms); 4096-ray batch {t4*1e6:.1f} us")
