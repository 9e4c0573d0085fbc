import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "contexture-nerf_b200")]
import torch
from ctxnerf import _lib as _l0; _l0.use_diag_lib()   # diagnostics build (include/ctxnerf_diag.h)
from ctxnerf import _lib
dev = torch.device("cuda", 0)
g = torch.Generator().manual_seed(0)
for (N, K) in ((256, 64), (128, 32), (256, 256), (64, 128), (32, 16)):
    A = torch.randn(256, K, generator=g).to(torch.bfloat16)
    B = torch.randn(N, K, generator=g).to(torch.bfloat16)
    ref = A.float() @ B.float().T
    Ad, Bd = A.to(dev), B.to(dev)          # keep the device tensors alive across the launch
    C = torch.zeros(256, N, device=dev)
    _lib.call("ctx_tcgen05_selftest2", _lib.ptr(Ad), _lib.ptr(Bd), _lib.ptr(C), N, K, _lib.stream_ptr(dev))
    torch.cuda.synchronize()
    err = (C.cpu() - ref).abs()
    print(f"selftest2 N={N} K={K} maxerr={err.max().item():.4g}", flush=True)
