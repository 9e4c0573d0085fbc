import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "contexture-nerf_b200")]
import torch
from ctxnerf import _lib as _l0; _l0.use_diag_lib()   # diagnostics build (include/ctxnerf_diag.h)
from ctxnerf import _lib
dev = torch.device("cuda", 0)
out = torch.zeros(16, dtype=torch.int64, device=dev)
_lib.call("ctx_tcgen05_sync_cost", _lib.ptr(out), 200, _lib.stream_ptr(dev))
torch.cuda.synchronize()
names = ["commit cta1 local", "commit cta2 local", "commit cta2 multicast", "remote arrive", "local arrive", "try_wait (done)", "try_wait.cluster (done)", "fence.proxy.async", "clock64"]
for n, v in zip(names, out.cpu().tolist()):
    print(f"{n:26s} {v} cycles")
