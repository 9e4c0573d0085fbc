import os, sys
sys.path[:0] = ["/root/repo", "/root/repo/contexture-nerf_b200", "/root/repo/tests"]
import torch
from oracle import nerf_oracle as orc
import test_gpu_mlp as T
dev = torch.device("cuda:0")
if "selftest" in sys.argv:
    T.test_tcgen05_descriptor_conventions(dev, 0)
    T.test_tcgen05_descriptor_conventions(dev, 1)
for (views, P) in ((False, 1000), (False, 4096), (True, 777)):
    net, params = T._net(dev, views, seed=P, in_pts=63, out_ch=4)
    g = torch.Generator().manual_seed(1)
    x = torch.randn(P, 90 if views else 63, generator=g).clamp(-1, 1)
    for rep in range(3):
        with torch.no_grad():
            out = net(x.to(dev)).cpu()
        ref = orc.mlp_forward_bf16(params, x, input_ch_views=27 if views else 0)
        err = (out - ref).abs()
        print(views, P, rep, "per-channel max err", [round(e, 4) for e in err.max(0)[0].tolist()])
        bad = (err.max(1)[0] > 1e-2).nonzero().flatten()
        print("   bad rows", len(bad), bad[:12].tolist(), bad[-5:].tolist())
        if len(bad): print(out[bad[:2]], ref[bad[:2]])
