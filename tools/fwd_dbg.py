import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "contexture-nerf_b200")]
import torch
from ctxnerf import run_nerf_helpers as rh, _lib
dev = torch.device("cuda", 0)
torch.manual_seed(0)
net = rh.NeRF().to(dev)
R, S = 4096, 192
o = torch.randn(R, 3, device=dev); d = torch.randn(R, 3, device=dev); d = d / d.norm(dim=-1, keepdim=True)
z = torch.sort(torch.rand(R, S, device=dev) * 4 + 2, -1)[0]
with torch.no_grad():
    for flags in (0, 1, 2, 3):
        _lib.lib().ctx_mlp_set_debug(flags)
        for _ in range(2):
            net.forward_rays(o, d, d, z)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); net.forward_rays(o, d, d, z); b.record()
        torch.cuda.synchronize()
        print(f"debug flags {flags} (1 = no epilogue body, 2 = no MMAs): {a.elapsed_time(b):.3f} ms", flush=True)
    _lib.lib().ctx_mlp_set_debug(0)
