"""Quick device-side timings of the individual kernels (CUDA events)."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "contexture-nerf_b200")]
import torch
from ctxnerf import ops, run_nerf_helpers as rh

dev = torch.device("cuda", 0)
PEAK_HBM = 6540.2e9
PEAK_TF = 1666.0e12


def timeit(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for a, b in evs:
        a.record(); fn(); b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in evs)
    return ts[len(ts) // 2] * 1e-3


def main():
    res = {}
    torch.manual_seed(0)
    net = rh.NeRF().to(dev)
    for R, S in ((4096, 64), (4096, 192), (65536, 64)):
        o = torch.randn(R, 3, device=dev); d = torch.randn(R, 3, device=dev)
        d = d / d.norm(dim=-1, keepdim=True)
        z = torch.sort(torch.rand(R, S, device=dev) * 4 + 2, -1)[0]
        with torch.no_grad():
            t = timeit(lambda: net.forward_rays(o, d, d, z))
        fl = 2 * 593408 * R * S
        res[f"mlp_fwd_R{R}_S{S}"] = dict(ms=t * 1e3, tflops=fl / t / 1e12, frac=fl / t / PEAK_TF)
    for R, S in ((1 << 20, 64), (1 << 20, 192), (1 << 18, 512)):
        raw = torch.randn(R, S, 4, device=dev); raw[..., 3] *= 5
        z = torch.sort(torch.rand(R, S, device=dev) * 4 + 2, -1)[0]
        d = torch.randn(R, 3, device=dev)
        t = timeit(lambda: ops.composite(raw, z, d))
        by = R * (24 * S + 36)
        res[f"composite_fwd_R{R}_S{S}"] = dict(ms=t * 1e3, gbs=by / t / 1e9, frac=by / t / PEAK_HBM)
        rawg = raw.clone().requires_grad_(True)
        outs = ops.composite(rawg, z, d)
        g = [torch.randn_like(o_) for o_ in outs]
        g[1].zero_()
        t = timeit(lambda: torch.autograd.grad(outs, rawg, g, retain_graph=True))
        by = R * (40 * S + 36)
        res[f"composite_bwd_R{R}_S{S}"] = dict(ms=t * 1e3, gbs=by / t / 1e9, frac=by / t / PEAK_HBM)
        w = torch.rand(R, S, device=dev)
        N = 128 if S == 64 else S
        t = timeit(lambda: ops.resample_merge(z, w, N, det=True))
        by = R * (4 * (S - 1) + 4 * (S - 2) + 4 * N + 4 * (S + N))
        res[f"resample_merge_R{R}_S{S}_N{N}"] = dict(ms=t * 1e3, gbs=by / t / 1e9, frac=by / t / PEAK_HBM)
        bins = z[:, :S - 1].contiguous(); w2 = w[:, :S - 2].contiguous()
        t = timeit(lambda: ops.resample_raw(bins, w2, N, det=True))
        by = R * (4 * (S - 1) + 4 * (S - 2) + 4 * N + 8 * N)
        res[f"resample_R{R}_B{S-1}_N{N}"] = dict(ms=t * 1e3, gbs=by / t / 1e9, frac=by / t / PEAK_HBM)
    n = 1 << 24
    x = torch.rand(n, 3, device=dev) * 4 - 2
    t = timeit(lambda: ops.posenc(x, 10))
    res["posenc_L10_n16M"] = dict(ms=t * 1e3, gbs=n * 264 / t / 1e9, frac=n * 264 / t / PEAK_HBM)
    K = [[1111.1, 0, 400.0], [0, 1111.1, 400.0], [0, 0, 1]]
    c2w = torch.eye(4, device=dev)[:3].contiguous()
    t = timeit(lambda: ops.raygen(3200, 3200, K, c2w, n_samples=64, near=2., far=6., perturb=True, seed=1, want_viewdirs=True))
    by = 3200 * 3200 * (36 + 256)
    res["raygen_3200x3200_S64"] = dict(ms=t * 1e3, gbs=by / t / 1e9, frac=by / t / PEAK_HBM)
    for k, v in res.items():
        print(k, {a: round(b, 4) for a, b in v.items()})
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(res, open("gpurun_out/perf_probe.json", "w"), indent=1)


if __name__ == "__main__":
    main()
