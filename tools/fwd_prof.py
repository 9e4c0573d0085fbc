import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "contexture-nerf_b200")]
import torch
from ctxnerf import _lib as _l0; _l0.use_diag_lib()   # diagnostics build (include/ctxnerf_diag.h)
from ctxnerf import run_nerf_helpers as rh, _lib
dev = torch.device("cuda", 0)
torch.manual_seed(0)
net = rh.NeRF().to(dev)
R, S = 4096, 192
o = torch.randn(R, 3, device=dev); d = torch.randn(R, 3, device=dev); d = d / d.norm(dim=-1, keepdim=True)
z = torch.sort(torch.rand(R, S, device=dev) * 4 + 2, -1)[0]
prof = torch.zeros(148 * 16, dtype=torch.int64, device=dev)
names = {0: "mmaB.wait_act", 1: "mmaB.wait_full", 2: "mma.wait_act", 3: "mma.wait_full", 4: "mma.wait_peer", 9: "mma.issue", 5: "mma.total",
         6: "epi2.wait_acc", 7: "epi2.body", 8: "epi2.total", 13: "epi2.encode", 10: "epi9.wait_acc", 11: "epi9.body", 12: "epi9.total"}
hang = torch.zeros(8 + 64 * 4, dtype=torch.int64).pin_memory()
_lib.lib().ctx_mlp_set_hang_buffer(hang.data_ptr())
def report():
    k = int(hang[0])
    print("HANG REPORT: waiters", k)
    TAGS = {1: "producer.empty", 2: "relay.full", 3: "issuerA.act", 13: "issuerB.act", 4: "issuerA.full", 14: "issuerB.full",
            5: "epi.accA", 15: "epi.accB"}
    for i in range(min(k, 60)):
        e = hang[8 + 4 * i: 12 + 4 * i].tolist()
        print("   ", TAGS.get(e[0], e[0]), "block", e[1] >> 32, "warp", e[1] & 0xffffffff, "info", e[2], "parity", e[3])
import atexit; atexit.register(report)
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
