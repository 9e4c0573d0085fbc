"""Time the fused get_texture_map (res 1024, the reference default) forward and forward+backward."""
import json, os, sys, torch
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "contexture-nerf_b200"))
from ctxnerf import run_nerf_helpers as rh
from ctxnerf.texture import get_texture_map
from ctxnerf import _lib

dev = torch.device("cuda:0")
torch.manual_seed(0)
net = rh.NeRF2D(D=8, W=256, input_ch=42, output_ch=3, skips=[4]).to(dev)
res = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
target = torch.rand(1, 3, res, res, device=dev)
out = {}
def timeit(fn, n=10, w=3):
    for _ in range(w): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
def fwd():
    with torch.no_grad():
        get_texture_map(net, res)
def fwdbwd():
    net.zero_grad(set_to_none=True)
    tex, _ = get_texture_map(net, res)
    ((tex - target) ** 2).mean().backward()
P = res * res
macs = 42*256 + 6*256*256 + (256+42)*256 + 256*3
t = timeit(fwd); out["fwd_ms"] = t; out["fwd_tflops"] = 2*macs*P/t/1e9
t = timeit(fwdbwd); out["fwdbwd_ms"] = t; out["fwdbwd_tflops"] = 6*macs*P/t/1e9
# the unfused way through the same kernels: materialised grid + embed + MLP + torch tanh
emb, _ = rh.get_embedder(10, 0, input_dims=2)
lin = torch.linspace(0, 1, res, device=dev)
def unfused():
    with torch.no_grad():
        u, v = torch.meshgrid(lin, lin, indexing="xy")
        uv = torch.stack([u, v], -1).reshape(-1, 2)
        o = net(emb(uv))
        ((o.tanh() + 1) / 2).reshape(1, res, res, 3).permute(0, 3, 1, 2).contiguous()
out["unfused_fwd_ms"] = timeit(unfused)
out["res"] = res
print(json.dumps(out))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/texture_bench.json", "w"))

# ---- texture_mapping + mask/background at the reference's render shape: 7 views x 1200 x 1200, 1024^2 x 3 atlas ----
from ctxnerf.texture import texture_mapping
B, Hh, Ww = 7, 1200, 1200
uv = torch.rand(B, Hh, Ww, 2, device=dev)
uv = (uv * 0.001 + torch.stack(torch.meshgrid(torch.linspace(0.05, 0.95, Hh, device=dev), torch.linspace(0.05, 0.95, Ww, device=dev),
                                             indexing="ij"), -1)[None]).clamp(0, 1).contiguous()   # smooth uv, as a rasteriser gives
mask = (torch.rand(B, Hh, Ww, 1, device=dev) > 0.4).float()
atlas = torch.rand(1, 3, 1024, 1024, device=dev, requires_grad=True)
gout = torch.randn(B, Hh, Ww, 3, device=dev)
def tm_fwd():
    with torch.no_grad():
        texture_mapping(uv, atlas, "bilinear", mask=mask, background=1.0)
def tm_fwdbwd():
    atlas.grad = None
    texture_mapping(uv, atlas, "bilinear", mask=mask, background=1.0).backward(gout)
def lib_fwdbwd():   # the library route the reference takes (kaolin -> torch grid_sample), for scale
    atlas.grad = None
    g = uv.reshape(B, -1, 1, 2) * 2 - 1
    g = torch.stack([g[..., 0], -g[..., 1]], -1)
    img = torch.nn.functional.grid_sample(atlas.expand(B, -1, -1, -1), g, mode="bilinear", align_corners=False, padding_mode="border")
    img = img.permute(0, 2, 3, 1).reshape(B, Hh, Ww, 3)
    img = img * mask + 1.0 * (1 - mask)
    img.backward(gout)
npx = B * Hh * Ww
t = timeit(tm_fwd); out["texmap_fwd_ms_v_along_x"] = t; out["texmap_fwd_gbs_v_along_x"] = npx * (8 + 4 + 12) / t / 1e6
uv = torch.flip(uv, dims=[-1]).contiguous()     # u along the image x axis: neighbouring pixels read neighbouring texels
t = timeit(tm_fwd); out["texmap_fwd_ms"] = t; out["texmap_fwd_gbs"] = npx * (8 + 4 + 12) / t / 1e6
t = timeit(tm_fwdbwd); out["texmap_fwdbwd_ms"] = t
out["torch_grid_sample_fwdbwd_ms"] = timeit(lib_fwdbwd)
print(json.dumps(out))
json.dump(out, open("gpurun_out/texture_bench.json", "w"))
