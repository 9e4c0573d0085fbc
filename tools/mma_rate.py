import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "contexture-nerf_b200")]
import torch
from ctxnerf import _lib as _l0; _l0.use_diag_lib()   # diagnostics build (include/ctxnerf_diag.h)
from ctxnerf import _lib
dev = torch.device("cuda", 0)
out = torch.zeros(512, dtype=torch.int64, device=dev)
src = torch.zeros(64 * 16384, dtype=torch.uint8, device=dev)
for nw, two in ((4, 0), (8, 0), (4, 1), (8, 1)):
    mode = 2 | (64 if two else 0) | (128 if nw == 8 else 0)
    out.zero_()
    _lib.call("ctx_tcgen05_mma_rate", 0, 400, 256, 148, _lib.ptr(out), mode, _lib.ptr(src), _lib.stream_ptr(dev))
    torch.cuda.synchronize()
    o = out.cpu()
    cyc = o[:148].double().mean().item()
    n = o[296:296 + 148].double().mean().item()
    loads = n * (2 if two else 1)
    print(f"{nw} reader warps, {2 if two else 1} LDTM.x32 in flight per warp: {nw*loads*4096/cyc:6.1f} B/clk/SM TMEM read under MMA load; MMA {cyc/(400*16):.1f} cyc")
