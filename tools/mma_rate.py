import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "contexture-nerf_b200")]
import torch
from ctxnerf import _lib
dev = torch.device("cuda", 0)
out = torch.zeros(512, dtype=torch.int64, device=dev)
src = torch.zeros(64 * 16384, dtype=torch.uint8, device=dev)
for iters, mode in ((400, 2), (400, 3), (1, 2), (1, 3)):
    out.zero_()
    _lib.call("ctx_tcgen05_mma_rate", 0, iters, 256, 148, _lib.ptr(out), mode, _lib.ptr(src), _lib.stream_ptr(dev))
    torch.cuda.synchronize()
    o = out.cpu()
    cyc = o[:148].double().mean().item()
    n = o[296:296 + 148].double().mean().item()
    print(f"iters {iters} mode {mode}: MMA {cyc/(iters*16):.1f} cyc/MMA, total {cyc:.0f} cycles; per-warp loop iterations {n:.0f} -> {cyc/max(n,1):.1f} cycles per LDTM(+STS) iteration")
