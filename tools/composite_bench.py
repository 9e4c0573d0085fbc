"""raw2outputs forward / backward at 2^20 rays as fractions of the measured HBM copy bandwidth (24S+36 / 40S+36 B per ray)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "contexture-nerf_b200")]
import torch
from ctxnerf import ops
dev = torch.device("cuda:0")
HBM = 6540.2e9
def timeit(fn, n=10):
    for _ in range(4): fn()
    torch.cuda.synchronize(); a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); [fn() for _ in range(n)]; b.record(); torch.cuda.synchronize(); return a.elapsed_time(b) / n * 1e-3
R = 1 << (19 if "--small" in sys.argv else 20)
S_LIST = [int(a) for a in sys.argv[1:] if a.isdigit()] or [32, 64, 128, 192, 256, 384, 512]
for S in S_LIST:
    raw = torch.randn(R, S, 4, device=dev); raw[..., 3] *= 5
    z = torch.sort(torch.rand(R, S, device=dev) * 4 + 2, -1)[0]; d = torch.randn(R, 3, device=dev)
    rr = raw.clone().requires_grad_(True); outs = ops.composite(rr, z, d)
    gs = [torch.randn_like(o) for o in outs]
    def bw():
        rr.grad = None; torch.autograd.backward(list(outs), gs, retain_graph=True)
    res = []
    for rep in range(2):
        with torch.no_grad():
            t = timeit(lambda: ops.composite(raw, z, d))
        tb = timeit(bw)
        res.append((round(R * (24 * S + 36) / t / HBM, 3), round(R * (40 * S + 36) / tb / HBM, 3)))
    print(f"S={S}: fwd/bwd fraction of copy bandwidth {res}")
    del raw, z, d, rr, outs, gs
