"""Stress the forward kernel for races: repeat launches, compare bitwise with the first result."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "contexture-nerf_b200")]
import torch
from ctxnerf import _lib as _l0; _l0.use_diag_lib()   # diagnostics build (include/ctxnerf_diag.h)
from ctxnerf import _lib
from ctxnerf import run_nerf_helpers as rh


def make_net(dev, views, seed, in_pts=63, out_ch=4):
    torch.manual_seed(seed)
    if views:
        net = rh.NeRF(D=8, W=256, input_ch=in_pts, input_ch_views=27, skips=[4], use_viewdirs=True)
    else:
        net = rh.NeRF2D(D=8, W=256, input_ch=in_pts, output_ch=out_ch, skips=[4])
    with torch.no_grad():
        for p in net.parameters():
            if p.dim() == 1:
                p.uniform_(-0.1, 0.1)
    return net.to(dev)


dev = torch.device("cuda:0")
flags = int(os.environ.get("CTX_DBG", "0"))
_lib.lib().ctx_mlp_set_debug(flags)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 200
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
import atexit
for (views, P, train) in ((False, 1000, False), (False, 4096, False), (True, 777, False), (True, 40000, False), (True, 20000, True)):
    net = make_net(dev, views, seed=P)
    g = torch.Generator().manual_seed(1)
    x = torch.randn(P, 90 if views else 63, generator=g).clamp(-1, 1).to(dev)
    bad = 0
    ref = None
    t0 = time.time()
    for rep in range(n):
      try:
        with (torch.enable_grad() if train else torch.no_grad()):
            out = net(x)
        if rep % 20 == 0: torch.cuda.synchronize()
      except Exception as ex:
        print("EXC at rep", rep, str(ex)[:200]); report(); sys.exit(1)
      if True:
        try:
          if ref is None:
            ref = out.detach().clone()
          elif not torch.equal(out.detach(), ref):
            bad += 1
            if bad <= 2:
                d = (out.detach() - ref).abs()
                rows = (d.max(1)[0] > 0).nonzero().flatten()
                print("   mismatch rep", rep, "rows", len(rows), rows[:8].tolist(), rows[-3:].tolist(), "max", d.max().item(), flush=True)
        except Exception as ex:
          print("EXC at rep", rep, str(ex)[:200]); report(); sys.exit(1)
    try:
        torch.cuda.synchronize()
    except Exception as ex:
        print("EXC at sync", str(ex)[:200]); report(); sys.exit(1)
    print(f"views={views} P={P} train={train}: {bad}/{n-1} mismatching launches, {time.time()-t0:.2f}s", flush=True)
