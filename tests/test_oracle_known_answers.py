"""Analytic known-answer tests for the restated upstream functions (SURVEY.md
8c): raw2outputs / render_rays have no implementation or test in the reference,
so these closed forms are what pins them."""
import math

import numpy as np
import pytest
import torch

from oracle import nerf_oracle as orc


def test_constant_density_closed_form():
    R, S, sigma, delta = 3, 16, 0.7, 0.25
    z = (2.0 + delta * torch.arange(S, dtype=torch.float32)).expand(R, S)
    d = torch.tensor([[0.0, 0.0, -1.0]]).expand(R, 3)
    raw = torch.zeros(R, S, 4)
    raw[..., 3] = sigma
    rgb, disp, acc, w, depth = orc.raw2outputs(raw, z, d)
    e = math.exp(-sigma * delta)
    expect_T = torch.tensor([(e + 1e-10) ** i for i in range(S)])
    alpha = torch.full((S,), 1 - e)
    alpha[-1] = 1.0
    torch.testing.assert_close(w[0], alpha * expect_T, rtol=1e-5, atol=1e-7)
    torch.testing.assert_close(acc, torch.ones(R), rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(rgb, torch.full((R, 3), 0.5), rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(depth[0], (alpha * expect_T * z[0]).sum(), rtol=1e-5, atol=1e-6)


def test_empty_ray_and_white_background():
    raw = -torch.rand(4, 8, 4) - 0.1
    z = torch.linspace(2, 6, 8).expand(4, 8)
    d = torch.randn(4, 3)
    rgb, disp, acc, w, depth = orc.raw2outputs(raw, z, d)
    assert (w == 0).all() and (acc == 0).all() and (rgb == 0).all()
    assert torch.isnan(disp).all()
    rgb_w = orc.raw2outputs(raw, z, d, white_bkgd=True)[0]
    assert (rgb_w == 1).all()


def test_single_opaque_sample():
    raw = torch.full((1, 10, 4), -5.0)
    raw[0, 6, 3] = 1e4
    raw[0, 6, :3] = torch.tensor([10.0, -10.0, 0.0])
    z = torch.linspace(2, 6, 10)[None]
    rgb, disp, acc, w, depth = orc.raw2outputs(raw, z, torch.tensor([[0.0, 1.0, 0.0]]))
    assert w[0, 6] == 1.0 and w[0].sum() == 1.0
    assert depth[0] == z[0, 6]
    torch.testing.assert_close(rgb[0], torch.tensor([1.0, 0.0, 0.5]), atol=1e-4, rtol=0)


def test_sample_pdf_uniform_and_one_hot():
    bins = torch.linspace(2, 6, 17)[None]
    s = orc.sample_pdf(bins, torch.zeros(1, 16), 33, det=True)
    u = torch.linspace(0, 1, 33)
    torch.testing.assert_close(s[0], 2 + 4 * u, rtol=1e-5, atol=1e-5)
    w = torch.zeros(1, 16)
    w[0, 5] = 1.0
    s = orc.sample_pdf(bins, w, 64, det=True)
    inside = (s >= bins[0, 5] - 1e-4) & (s <= bins[0, 6] + 1e-4)
    assert inside[0, 1:-1].all()
    assert (s[0, 1:] >= s[0, :-1]).all()


def test_get_rays_centre_and_corner():
    K = [[1111.0, 0, 400.0], [0, 1111.0, 400.0], [0, 0, 1]]
    c2w = torch.eye(4)[:3]
    ro, rd = orc.get_rays(800, 800, K, c2w)
    assert torch.equal(rd[400, 400], torch.tensor([0.0, 0.0, -1.0]))
    torch.testing.assert_close(rd[0, 0], torch.tensor([-400 / 1111.0, 400 / 1111.0, -1.0]))
    assert (ro == 0).all()


def test_channel_order_and_skip_layout():
    x = torch.tensor([[0.5, -0.25, 1.0]])
    y = orc.posenc(x, 2)
    expect = torch.cat([x, torch.sin(x), torch.cos(x), torch.sin(2 * x), torch.cos(2 * x)], -1)
    assert torch.equal(y, expect)
    p = orc.init_mlp_params(63, 4)
    assert p["pts_linears.5.weight"].shape == (256, 256 + 63)
    # columns [:63] of the skip layer multiply x
    xin = torch.randn(2, 63)
    p0 = {k: v.clone() for k, v in p.items()}
    p0["pts_linears.5.weight"][:, :63] = 0
    assert not torch.allclose(orc.mlp_forward(p, xin), orc.mlp_forward(p0, xin))
    assert orc.mlp_macs(63, 4) == 492032 and orc.mlp_macs(42, 3) == 481024
    assert orc.mlp_macs(63, 4, input_ch_views=27) == 593408


def test_render_rays_contract_and_determinism():
    g = torch.Generator().manual_seed(0)
    p = orc.init_mlp_params(63, 4, input_ch_views=27, generator=g)
    q = lambda pts, vd, prm: orc.run_network(pts, vd, prm)
    R = 8
    o = torch.zeros(R, 3)
    d = torch.randn(R, 3, generator=g)
    vd = d / d.norm(dim=-1, keepdim=True)
    rb = torch.cat([o, d, torch.full((R, 1), 2.0), torch.full((R, 1), 6.0), vd], -1)
    out = orc.render_rays(rb, p, q, 16, N_importance=8, network_fine=p, white_bkgd=True)
    for k in ("rgb_map", "disp_map", "acc_map", "rgb0", "disp0", "acc0", "z_std"):
        assert k in out
    assert out["z_vals"].shape == (R, 24) and (out["z_vals"][:, 1:] >= out["z_vals"][:, :-1]).all()
    out2 = orc.render_rays(rb, p, q, 16, N_importance=8, network_fine=p, white_bkgd=True)
    assert torch.equal(out["rgb_map"], out2["rgb_map"])


def test_stratified_bounds():
    near, far = torch.full((5, 1), 2.0), torch.full((5, 1), 6.0)
    z0 = orc.stratified_z(near, far, 64)
    assert torch.equal(z0[0], torch.linspace(0, 1, 64) * 0 + (2.0 * (1 - torch.linspace(0, 1, 64)) + 6.0 * torch.linspace(0, 1, 64)))
    j = torch.rand(5, 64)
    z = orc.stratified_z(near, far, 64, jitter=j)
    mids = 0.5 * (z0[:, 1:] + z0[:, :-1])
    assert (z[:, 1:-1] >= mids[:, :-1]).all() and (z[:, 1:-1] <= mids[:, 1:]).all()


def test_texture_uv_grid_ordering():
    """'xy' meshgrid: u varies fastest (columns), v along rows; end points exact (textured_mesh.py:268-272)."""
    g = orc.texture_uv_grid(5)
    assert g.shape == (25, 2)
    assert g[0].tolist() == [0.0, 0.0] and g[4].tolist() == [1.0, 0.0]
    assert g[5].tolist() == [0.0, 0.25] and g[24].tolist() == [1.0, 1.0]
    torch.manual_seed(0)
    p = orc.init_mlp_params(42, 3, W=32)
    tex, raw = orc.texture_map(p, 8)
    assert tex.shape == (1, 3, 8, 8) and raw.shape == (64, 3)
    assert torch.allclose(tex[0, :, 2, 3], (torch.tanh(raw[2 * 8 + 3]) + 1) / 2)


def test_texture_mapping_restatement():
    """kaolin's texture_mapping as restated (render.py:135): v points up, texel centres reproduce the texels, uv outside
    [0,1] clamps to the border, and the mask / background lines of render.py:137-140."""
    res = 8
    tex = torch.arange(res * res, dtype=torch.float32).reshape(1, 1, res, res)
    xs = (torch.arange(res) + 0.5) / res
    uv = torch.stack(torch.meshgrid(xs, xs, indexing="xy"), -1)[None]
    for mode in ("nearest", "bilinear"):
        out = orc.texture_mapping(uv, tex, mode)[0, ..., 0]
        assert torch.allclose(out, torch.flip(tex[0, 0], dims=[0]), atol=1e-5)
    corner = orc.texture_mapping(torch.tensor([[[-0.3, 1.7]]]), tex, "bilinear")
    assert corner.item() == tex[0, 0, 0, 0].item()
    mask = torch.tensor([[[1.0], [0.0]]])
    img = orc.render_composite(torch.tensor([[[0.5, 0.5], [0.5, 0.5]]]), tex, mask, 1.0)
    assert img[0, 1, 0].item() == 1.0 and img[0, 0, 0].item() != 1.0


# ---------------------------------------------------------------------------------------------------------
# raw2outputs is unpinned (absent from the reference): cross-check the torch oracle against a second,
# independently written float64 restatement (oracle/raw2outputs_fp64.py: scalar loops, running transmittance,
# hand-derived O(S^2) gradient).
@pytest.mark.parametrize("white", [False, True])
def test_raw2outputs_agrees_with_independent_fp64_restatement(white):
    from oracle.raw2outputs_fp64 import raw2outputs_fp64
    R, S = 24, 40
    raw, z, d = orc.cfg1_inputs(R, S, seed=3)
    z = torch.sort(torch.rand(R, S, generator=torch.Generator().manual_seed(4)) * 4 + 2, -1)[0]
    raw[0, :, 3] = -1.0                   # an empty ray (acc = 0, disp = NaN)
    raw[1, 7, 3] = 1e4                    # an opaque sample
    d[2] *= 3.0                           # non-unit direction: distances scale with its norm
    got = orc.raw2outputs(raw, z, d, white_bkgd=white)
    ref = raw2outputs_fp64(raw.numpy(), z.numpy(), d.numpy(), white_bkgd=white)
    for name, a in zip(("rgb", "disp", "acc", "weights", "depth"), got):
        b = torch.from_numpy(ref[name]).float()
        if name == "disp":
            assert torch.equal(torch.isnan(a), torch.isnan(b))
            ok = ~torch.isnan(b)
            torch.testing.assert_close(a[ok], b[ok], rtol=2e-5, atol=1e-7)
        else:
            # fp32 oracle vs fp64 restatement: rounding of S-term sums / products only
            torch.testing.assert_close(a, b, rtol=2e-5, atol=2e-6, msg=lambda m: f"{name}: {m}")


def test_raw2outputs_gradient_agrees_with_hand_derived_fp64():
    from oracle.raw2outputs_fp64 import raw2outputs_fp64_backward
    R, S = 6, 17
    g = torch.Generator().manual_seed(8)
    raw = torch.randn(R, S, 4, generator=g, dtype=torch.float64)
    raw[..., 3] *= 2.0
    z = torch.sort(torch.rand(R, S, generator=g, dtype=torch.float64) * 4 + 2, -1)[0]
    d = torch.randn(R, 3, generator=g, dtype=torch.float64)
    g_rgb, g_acc = torch.randn(R, 3, generator=g, dtype=torch.float64), torch.randn(R, generator=g, dtype=torch.float64)
    g_w, g_depth = torch.randn(R, S, generator=g, dtype=torch.float64), torch.randn(R, generator=g, dtype=torch.float64)
    for white in (False, True):
        r0 = raw.clone().requires_grad_(True)
        rgb, disp, acc, w, depth = orc.raw2outputs(r0, z, d, white_bkgd=white)      # the oracle runs in any dtype
        ((rgb * g_rgb).sum() + (acc * g_acc).sum() + (w * g_w).sum() + (depth * g_depth).sum()).backward()
        ref = raw2outputs_fp64_backward(raw.numpy(), z.numpy(), d.numpy(), g_rgb.numpy(), g_acc.numpy(), g_w.numpy(),
                                        g_depth.numpy(), white_bkgd=white)
        torch.testing.assert_close(r0.grad, torch.from_numpy(ref), rtol=1e-9, atol=1e-12)


def test_view_weight_masks_known_answer():
    """trainer.py:155-249: a face seen by two views keeps only the pixels of the view with the larger z-normal."""
    face_idx = torch.full((2, 1, 2, 3), -1, dtype=torch.int64)
    face_idx[0, 0, 0, :2] = 0          # view 0 sees face 0 on two pixels, face 1 on one
    face_idx[0, 0, 1, 2] = 1
    face_idx[1, 0, 0, 0] = 0           # view 1 sees face 0 (with a larger z) and face 2
    face_idx[1, 0, 1, 1] = 2
    normals = torch.zeros(2, 3, 3)
    normals[0, 2] = torch.tensor([0.3, 0.5, 0.9])
    normals[1, 2] = torch.tensor([0.7, 0.9, -0.2])     # face 1 is NOT visible in view 1: its 0.9 must not count
    rows = orc.create_face_view_map(face_idx)
    assert rows.tolist() == [[0, 0, 0, 0], [0, 0, 0, 1], [1, 0, 1, 2], [0, 1, 0, 0], [2, 1, 1, 1]]
    m = orc.compare_face_normals_between_views(rows, normals, face_idx)
    expect = torch.ones(2, 1, 2, 3, dtype=torch.bool)
    expect[0, 0, 0, :2] = False        # face 0: view 1 wins (0.7 > 0.3)
    assert torch.equal(m, expect)
