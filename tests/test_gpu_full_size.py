"""The BASELINE configurations at FULL size on the GPU, through properties that do not need the oracle to finish
(it takes minutes per 10^5 rays) plus an oracle check on a strided subset of the same launch:

* config 2 (800x800 = 640 000 rays, coarse 64 + fine 128): merged depths sorted and inside [near, far], weights >= 0,
  acc = sum(weights) <= 1, white-background rgb in [0, 1]; the render is independent of how the pixels are batched
  (whole image == two halves == a shuffled pixel list un-shuffled, bit for bit) and of repetition;
* config 4 (1024x1024 x 192 samples, bounding sphere): the same batching invariance for the single-pass form;
* config 5 (2^20 rays): raw2outputs and sample_pdf on the largest sweep shapes -- batching invariance, sortedness and
  range of the samples, and a strided 2048-ray subset of the SAME launch against the CPU oracle (bit-exact samples
  and indices, 1e-5 compositing).
Everything goes through the C-ABI (ctx_render_rays / ctx_composite_fwd / ctx_resample_fwd)."""
import numpy as np
import pytest
import torch

from oracle import nerf_oracle as orc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def trainer(cuda):
    from ctxnerf.train import NerfTrainer
    from ctxnerf.workloads import orbit_camera
    K, c2w = orbit_camera()
    return NerfTrainer(800, 800, K, c2w, perturb=0.0, white_bkgd=True, device=cuda, seed=5)


def _same(a, b):
    return torch.equal(torch.nan_to_num(a, nan=-7.0), torch.nan_to_num(b, nan=-7.0))


def test_config2_full_image_properties_and_batching_invariance(cuda, trainer):
    tr = trainer
    n = tr.H * tr.W
    full = tr._render(None, perturb=False)
    rgb, disp, acc, w, depth = full["comp_f"]
    z = full["z_f"]
    torch.cuda.synchronize()
    assert z.shape == (n, tr.N_samples + tr.N_importance)
    assert bool((z[:, 1:] >= z[:, :-1]).all()), "merged depths must be sorted"
    assert float(z.min()) >= tr.near - 1e-6 and float(z.max()) <= tr.far + 1e-6
    assert bool(torch.isfinite(w).all()) and float(w.min()) >= 0.0
    assert float(acc.max()) <= 1.0 + 1e-5 and float(acc.min()) >= 0.0
    assert float((acc - w.sum(-1)).abs().max()) <= 1e-5
    assert bool(torch.isfinite(rgb).all()) and float(rgb.min()) >= -1e-6 and float(rgb.max()) <= 1.0 + 1e-5
    assert float(depth.min()) >= 0.0 and float(depth.max()) <= tr.far * (1.0 + 1e-5)
    # the same image as two halves, and as a shuffled pixel list: bit-identical
    idx = torch.arange(n, device=cuda)
    lo, hi = tr.render(idx[: n // 2 + 77]), tr.render(idx[n // 2 + 77:])
    assert _same(torch.cat([lo["rgb_map"], hi["rgb_map"]]), rgb)
    assert _same(torch.cat([lo["disp_map"], hi["disp_map"]]), disp)
    assert _same(torch.cat([lo["depth_map"], hi["depth_map"]]), depth)
    perm = torch.randperm(n, device=cuda, generator=torch.Generator(device=cuda).manual_seed(1))
    sh = tr.render(perm)
    back = torch.empty_like(rgb)
    back[perm] = sh["rgb_map"]
    assert _same(back, rgb)
    # and of repetition
    again = tr.render(None)
    assert _same(again["rgb_map"], rgb) and _same(again["acc_map"], acc)


def test_config4_view_batching_invariance(cuda, trainer):
    from ctxnerf.workloads import multiview_cameras
    cams, sph = multiview_cameras()
    Kv, cv = cams[3]
    H = W = 1024
    full = trainer.render_view(H, W, Kv, cv, n_samples=192, sphere=sph)
    torch.cuda.synchronize()
    acc = full["acc_map"].reshape(-1)
    assert bool(torch.isfinite(full["rgb_map"]).all()) and float(acc.max()) <= 1.0 + 1e-5 and float(acc.min()) >= 0.0
    idx = torch.arange(H * W, device=cuda)
    parts = [trainer.render_view(H, W, Kv, cv, n_samples=192, sphere=sph, ray_idx=c) for c in idx.split(400_003)]
    assert _same(torch.cat([p["rgb_map"] for p in parts]).reshape(H, W, 3), full["rgb_map"])
    assert _same(torch.cat([p["depth_map"] for p in parts]).reshape(H, W), full["depth_map"])


@pytest.mark.parametrize("S", [64, 192, 512])
def test_config5_raw2outputs_at_2_pow_20_rays(cuda, S):
    from ctxnerf import ops
    R = 1 << 20 if S <= 192 else 1 << 18
    g = torch.Generator(device=cuda).manual_seed(S)
    raw = torch.randn(R, S, 4, device=cuda, generator=g)
    raw[..., 3] *= 5.0
    z = torch.sort(torch.rand(R, S, device=cuda, generator=g) * 4 + 2, -1)[0]
    d = torch.randn(R, 3, device=cuda, generator=g)
    rgb, disp, acc, w, depth = ops.composite(raw, z, d, None, True)
    torch.cuda.synchronize()
    assert bool(torch.isfinite(w).all()) and float(w.min()) >= 0.0 and float(acc.max()) <= 1.0 + 1e-5
    assert float((w.sum(-1) - acc).abs().max()) <= 1e-5 * max(1.0, S / 64)
    # batching invariance (the grid-stride loop and the rays-per-warp grouping must not leak between rays)
    cut = R // 3 + 5
    a, b = ops.composite(raw[:cut], z[:cut], d[:cut], None, True), ops.composite(raw[cut:], z[cut:], d[cut:], None, True)
    for full_t, pa, pb in zip((rgb, disp, acc, w, depth), a, b):
        assert _same(torch.cat([pa, pb]), full_t)
    # a strided subset of the same launch against the oracle
    sel = torch.arange(0, R, R // 2048, device=cuda)[:2048]
    ref = orc.raw2outputs(raw[sel].cpu(), z[sel].cpu(), d[sel].cpu(), white_bkgd=True)
    for name, got, want in zip(("rgb", "disp", "acc", "weights", "depth"), (rgb, disp, acc, w, depth), ref):
        got = got[sel].cpu()
        ok = torch.isfinite(want)
        tol = 1e-5 * want[ok].abs().clamp_min(1.0 if name != "weights" else 1e-3) * max(1.0, S / 64)
        assert bool(((got[ok] - want[ok]).abs() <= tol + 2.4e-7).all()), name
        assert bool((torch.isnan(got) == torch.isnan(want)).all()), name


@pytest.mark.parametrize("S,N", [(64, 128), (256, 512)])
def test_config5_sample_pdf_at_2_pow_20_rays(cuda, S, N):
    from ctxnerf import ops
    R = 1 << 20 if S <= 64 else 1 << 18
    g = torch.Generator(device=cuda).manual_seed(N)
    z = torch.sort(torch.rand(R, S, device=cuda, generator=g) * 4 + 2, -1)[0]
    w = torch.rand(R, S, device=cuda, generator=g) ** 3
    bins, wc = z[:, :S - 1].contiguous(), w[:, :S - 2].contiguous()
    smp, inds = ops.resample_raw(bins, wc, N, det=True)
    torch.cuda.synchronize()
    assert bool((smp[:, 1:] >= smp[:, :-1]).all()), "det=True samples are non-decreasing"
    assert bool((smp >= bins[:, :1]).all()) and bool((smp <= bins[:, -1:]).all())
    assert int(inds.min()) >= 1 and int(inds.max()) <= S - 1
    cut = R // 2 + 3
    a, _ = ops.resample_raw(bins[:cut], wc[:cut], N, det=True)
    b, _ = ops.resample_raw(bins[cut:], wc[cut:], N, det=True)
    assert torch.equal(torch.cat([a, b]), smp)
    sel = torch.arange(0, R, R // 2048, device=cuda)[:2048]
    s_ref, i_ref = orc.sample_pdf(bins[sel].cpu(), wc[sel].cpu(), N, det=True, return_inds=True)
    assert torch.equal(inds[sel].cpu(), i_ref) and torch.equal(smp[sel].cpu(), s_ref)
    # the fused form on the depths themselves: merged row = sort(cat[z, samples]) (checked on the subset)
    zs, z_all = ops.resample_merge(z, w, N, det=True)
    torch.cuda.synchronize()
    assert bool((z_all[:, 1:] >= z_all[:, :-1]).all())
    zs_ref = orc.sample_pdf(0.5 * (z[sel, 1:] + z[sel, :-1]).cpu(), w[sel, 1:-1].cpu(), N, det=True)
    assert torch.equal(zs[sel].cpu(), zs_ref)
    assert torch.equal(z_all[sel].cpu(), torch.sort(torch.cat([z[sel].cpu(), zs_ref], -1), -1)[0])
