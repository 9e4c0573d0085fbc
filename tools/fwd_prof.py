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
prof = torch.zeros(148 * 16, dtype=torch.int64, device=dev)
names = {14: "epi2.ldtm", 15: "epi2.consume", 1: "epi2.emit", 2: "mma.wait_act", 3: "mma.wait_full", 4: "mma.wait_peer", 9: "mma.issue", 5: "mma.total",
         6: "epi2.wait_acc", 7: "epi2.body", 8: "epi2.total", 13: "epi2.encode", 10: "epi9.wait_acc", 11: "epi9.body", 12: "epi9.total"}
TRAIN = "train" in sys.argv
if TRAIN: sys.argv.remove("train")
with (torch.enable_grad() if TRAIN else torch.no_grad()):
    for flags in [int(x) for x in (sys.argv[1:] or ["0"])]:
        _lib.lib().ctx_mlp_set_debug(flags)
        for _ in range(2):
            net.forward_rays(o, d, d, z)
        torch.cuda.synchronize()
        prof.zero_()
        _lib.lib().ctx_mlp_set_prof_buffer(_lib.ptr(prof))
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); net.forward_rays(o, d, d, z); b.record()
        torch.cuda.synchronize()
        _lib.lib().ctx_mlp_set_prof_buffer(None)
        print(f"== debug flags {flags}: kernel {a.elapsed_time(b):.3f} ms")
        p = prof.cpu().reshape(148, 16).double()
        for i, n in names.items():
            for par, tag in ((0, "leader"), (1, "peer")):
                col = p[par::2, i]
                nz = col[col > 0]
                if len(nz): print(f"   {n:14s} {tag:6s} mean {nz.mean().item():10.0f}  (per phase {nz.mean().item()/420:7.0f})")
    _lib.lib().ctx_mlp_set_debug(0)
