"""CPU oracle for the NeRF ray-march path.  TEST INFRASTRUCTURE ONLY.

This module is a CPU (torch fp32) restatement of the algorithm behind
``/root/reference/src/run_nerf_helpers.py`` plus the upstream
``raw2outputs`` / ``render_rays`` contract the reference only points at
(comment at ``src/run_nerf_helpers.py:131-133``).  It exists so the CUDA
kernels can be checked; nothing in the product package imports it.  Only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import it.

Pinning status (see DESIGN.md "Oracle"):
  * posenc, coordinate MLP (NeRF2D), get_rays, ndc_rays, sample_pdf are pinned
    against the reference itself: ``tests/golden/make_golden.py`` imports the
    unmodified reference module in the build container and the fixtures it
    writes are committed under ``tests/golden/``.
  * raw2outputs, render_rays, run_network and the view-direction head do not
    exist in the reference tree.  They follow
    the un-vendored, un-pinned upstream ``yenchenlin/nerf-pytorch`` module
    ``run_nerf.py``; for those functions **parity is unpinned** beyond the
    analytic known-answer tests in ``tests/test_oracle_known_answers.py``.

Numeric conventions fixed here (they make "bit-exact" well defined):
  * every elementwise product/sum is a separate fp32 operation (no FMA), as in
    eager PyTorch;
  * ``sample_pdf``: the normalising sum is the exactly-rounded fp32 value of
    the fp64 sum (torch's own CPU ``sum`` is a SIMD-width dependent cascade and
    is not reproducible across hosts); the CDF is the fp64 running sum rounded
    to fp32 per element, which is what torch's CPU ``cumsum`` does.  The
    inverse-CDF stage can be driven with an injected ``cdf`` so that stage is
    comparable bit for bit with the reference.
"""
from __future__ import annotations

import math
from typing import Callable, Dict, Optional, Sequence, Tuple

import numpy as np
import torch

# --------------------------------------------------------------------------
# a1/a2  positional encoding          reference: src/run_nerf_helpers.py:15-45
# --------------------------------------------------------------------------

def frequency_bands(num_freqs: int, log_sampling: bool = True) -> torch.Tensor:
    """Frequencies of the encoder (reference :29-34). Exact powers of two when
    log_sampling is on."""
    top = float(num_freqs - 1)
    if log_sampling:
        return 2.0 ** torch.linspace(0.0, top, steps=num_freqs)
    return torch.linspace(1.0, 2.0 ** top, steps=num_freqs)


def posenc(x: torch.Tensor, num_freqs: int, include_input: bool = True,
           log_sampling: bool = True) -> torch.Tensor:
    """[..., d] -> [..., d*(include_input + 2*num_freqs)].

    Channel order (reference :24-39): x, then for each band sin(f*x), cos(f*x),
    each block d wide.  The argument ``x*f`` is rounded to fp32 before the
    trigonometric call (:38)."""
    parts = [x] if include_input else []
    for f in frequency_bands(num_freqs, log_sampling):
        arg = x * f
        parts.append(torch.sin(arg))
        parts.append(torch.cos(arg))
    return torch.cat(parts, dim=-1)


def posenc_out_dim(d: int, num_freqs: int, include_input: bool = True) -> int:
    return d * ((1 if include_input else 0) + 2 * num_freqs)


# --------------------------------------------------------------------------
# a4/a5  coordinate MLP                reference: src/run_nerf_helpers.py:68-135
# --------------------------------------------------------------------------

def init_mlp_params(input_ch: int, output_ch: int, D: int = 8, W: int = 256,
                    skips: Sequence[int] = (4,), input_ch_views: int = 0,
                    generator: Optional[torch.Generator] = None) -> Dict[str, torch.Tensor]:
    """Parameter dict with the reference's names/shapes (:81-83, :97) and init
    (:99-104: kaiming-normal fan_in/relu weights, nn.Linear default biases).
    With ``input_ch_views > 0`` the upstream view-direction head is added
    (commented lines :86-95): feature_linear, alpha_linear, views_linears.0,
    rgb_linear."""
    g = generator

    def kaiming(o, i):
        return torch.randn(o, i, generator=g) * math.sqrt(2.0 / i)

    def default_w(o, i):  # nn.Linear default: U(-1/sqrt(i), 1/sqrt(i))
        b = 1.0 / math.sqrt(i)
        return (torch.rand(o, i, generator=g) * 2 - 1) * b

    def default_b(o, i):
        b = 1.0 / math.sqrt(i)
        return (torch.rand(o, generator=g) * 2 - 1) * b

    p: Dict[str, torch.Tensor] = {}
    fan = input_ch
    for l in range(D):
        p[f"pts_linears.{l}.weight"] = kaiming(W, fan)
        p[f"pts_linears.{l}.bias"] = default_b(W, fan)
        fan = W + input_ch if l in skips else W
    if input_ch_views > 0:
        p["feature_linear.weight"] = default_w(W, W)
        p["feature_linear.bias"] = default_b(W, W)
        p["alpha_linear.weight"] = default_w(1, W)
        p["alpha_linear.bias"] = default_b(1, W)
        p["views_linears.0.weight"] = default_w(W // 2, W + input_ch_views)
        p["views_linears.0.bias"] = default_b(W // 2, W + input_ch_views)
        p["rgb_linear.weight"] = default_w(3, W // 2)
        p["rgb_linear.bias"] = default_b(3, W // 2)
    else:
        p["output_linear.weight"] = kaiming(output_ch, W)
        p["output_linear.bias"] = default_b(output_ch, W)
    return p


def mlp_forward(p: Dict[str, torch.Tensor], x: torch.Tensor, D: int = 8,
                skips: Sequence[int] = (4,), input_ch_views: int = 0) -> torch.Tensor:
    """NeRF2D.forward (:106-135).  Skip concat puts the *input first* (:115).
    No output activation.  With views: x = [pts_enc | view_enc]; output is
    [rgb(3) | alpha(1)] as in the commented upstream branch (:117-127)."""
    if input_ch_views > 0:
        pts, views = x[..., :-input_ch_views], x[..., -input_ch_views:]
    else:
        pts, views = x, None
    h = pts
    for l in range(D):
        h = torch.relu(h @ p[f"pts_linears.{l}.weight"].T + p[f"pts_linears.{l}.bias"])
        if l in skips:
            h = torch.cat([pts, h], dim=-1)
    if views is None:
        return h @ p["output_linear.weight"].T + p["output_linear.bias"]
    alpha = h @ p["alpha_linear.weight"].T + p["alpha_linear.bias"]
    feat = h @ p["feature_linear.weight"].T + p["feature_linear.bias"]
    hv = torch.cat([feat, views], dim=-1)
    hv = torch.relu(hv @ p["views_linears.0.weight"].T + p["views_linears.0.bias"])
    rgb = hv @ p["rgb_linear.weight"].T + p["rgb_linear.bias"]
    return torch.cat([rgb, alpha], dim=-1)


def mlp_macs(input_ch: int, output_ch: int, D: int = 8, W: int = 256,
             skips: Sequence[int] = (4,), input_ch_views: int = 0) -> int:
    """Multiply-accumulates per evaluated point."""
    macs, fan = 0, input_ch
    for l in range(D):
        macs += fan * W
        fan = W + input_ch if l in skips else W
    if input_ch_views > 0:
        macs += W * W + W + (W + input_ch_views) * (W // 2) + (W // 2) * 3
    else:
        macs += W * output_ch
    return macs


# --------------------------------------------------------------------------
# a6/a7/a8  rays                      reference: src/run_nerf_helpers.py:139-178
# --------------------------------------------------------------------------

def get_rays(H: int, W: int, K, c2w: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """Pinhole rays, OpenGL camera (:139-148).  Pixel (row y, col x) gets
    dir_cam = ((x-cx)/fx, -(y-cy)/fy, -1); the rotation is a broadcast multiply
    followed by a 3-term sum, i.e. ((a*r0 + b*r1) + c*r2) in fp32 without FMA.
    rays_o is a stride-0 expansion of the camera centre."""
    c2w = torch.as_tensor(c2w, dtype=torch.float32)
    xs = torch.linspace(0, W - 1, W)
    ys = torch.linspace(0, H - 1, H)
    px = xs[None, :].expand(H, W)
    py = ys[:, None].expand(H, W)
    a = (px - K[0][2]) / K[0][0]
    b = -(py - K[1][2]) / K[1][1]
    c = -torch.ones_like(a)
    rot = c2w[:3, :3]
    comps = []
    for k in range(3):
        comps.append((a * rot[k, 0] + b * rot[k, 1]) + c * rot[k, 2])
    rays_d = torch.stack(comps, dim=-1)
    rays_o = c2w[:3, -1].expand(rays_d.shape)
    return rays_o, rays_d


def ndc_rays(H, W, focal, near, rays_o: torch.Tensor, rays_d: torch.Tensor):
    """Forward-facing NDC warp (:161-178), same operation order."""
    t = -(near + rays_o[..., 2]) / rays_d[..., 2]
    o = rays_o + t[..., None] * rays_d
    sx = -1.0 / (W / (2.0 * focal))
    sy = -1.0 / (H / (2.0 * focal))
    oz = o[..., 2]
    o_ndc = torch.stack([sx * o[..., 0] / oz, sy * o[..., 1] / oz, 1.0 + 2.0 * near / oz], -1)
    dz = rays_d[..., 2]
    d_ndc = torch.stack([sx * (rays_d[..., 0] / dz - o[..., 0] / oz),
                         sy * (rays_d[..., 1] / dz - o[..., 1] / oz),
                         -2.0 * near / oz], -1)
    return o_ndc, d_ndc


# --------------------------------------------------------------------------
# stratified depths (upstream render_rays)
# --------------------------------------------------------------------------

def stratified_z(near: torch.Tensor, far: torch.Tensor, n_samples: int,
                 lindisp: bool = False, jitter: Optional[torch.Tensor] = None) -> torch.Tensor:
    """near/far: [R,1].  jitter: None (perturb == 0) or uniforms [R, n_samples]."""
    t = torch.linspace(0.0, 1.0, steps=n_samples)
    if lindisp:
        z = 1.0 / (1.0 / near * (1.0 - t) + 1.0 / far * t)
    else:
        z = near * (1.0 - t) + far * t
    z = z.expand(near.shape[0], n_samples)
    if jitter is not None:
        mids = 0.5 * (z[..., 1:] + z[..., :-1])
        upper = torch.cat([mids, z[..., -1:]], -1)
        lower = torch.cat([z[..., :1], mids], -1)
        z = lower + (upper - lower) * jitter
    return z


# --------------------------------------------------------------------------
# a9  inverse-CDF resampling           reference: src/run_nerf_helpers.py:182-225
# --------------------------------------------------------------------------

def pdf_to_cdf(weights: torch.Tensor) -> torch.Tensor:
    """(:184-187) with the summation order fixed (module docstring)."""
    w = weights + 1e-5
    total = w.double().sum(-1, keepdim=True).float()
    pdf = w / total
    cdf = torch.cumsum(pdf.double(), -1).float()
    return torch.cat([torch.zeros_like(cdf[..., :1]), cdf], -1)


def det_uniforms(n: int) -> torch.Tensor:
    return torch.linspace(0.0, 1.0, steps=n)


def sample_pdf(bins: torch.Tensor, weights: torch.Tensor, n_samples: int, det: bool = False,
               u: Optional[torch.Tensor] = None, cdf: Optional[torch.Tensor] = None,
               return_inds: bool = False):
    """bins [R,B], weights [R,B-1] -> samples [R,n_samples] (:182-225).

    ``u`` overrides the uniforms (reference: torch.rand, :194); ``cdf``
    overrides stage 1 so that stage 2 (:208-223) can be compared bit for bit
    with the reference's own cdf."""
    if cdf is None:
        cdf = pdf_to_cdf(weights)
    B = cdf.shape[-1]
    if u is None:
        if det:
            u = det_uniforms(n_samples).expand(*cdf.shape[:-1], n_samples)
        else:
            u = torch.rand(*cdf.shape[:-1], n_samples)
    u = u.contiguous()
    inds = torch.searchsorted(cdf.contiguous(), u, right=True)      # first idx with cdf > u
    lo = (inds - 1).clamp(min=0)
    hi = inds.clamp(max=B - 1)
    cdf_lo, cdf_hi = torch.gather(cdf, -1, lo), torch.gather(cdf, -1, hi)
    bin_lo, bin_hi = torch.gather(bins, -1, lo), torch.gather(bins, -1, hi)
    denom = cdf_hi - cdf_lo
    denom = torch.where(denom < 1e-5, torch.ones_like(denom), denom)
    frac = (u - cdf_lo) / denom
    samples = bin_lo + frac * (bin_hi - bin_lo)
    if return_inds:
        return samples, inds
    return samples


# --------------------------------------------------------------------------
# a10  volume compositing (upstream raw2outputs) -- UNPINNED
# --------------------------------------------------------------------------

def raw2outputs(raw: torch.Tensor, z_vals: torch.Tensor, rays_d: torch.Tensor,
                raw_noise_std: float = 0.0, white_bkgd: bool = False,
                noise: Optional[torch.Tensor] = None):
    """raw [R,S,4], z_vals [R,S], rays_d [R,3] ->
    (rgb_map [R,3], disp_map [R], acc_map [R], weights [R,S], depth_map [R])."""
    gaps = z_vals[..., 1:] - z_vals[..., :-1]
    gaps = torch.cat([gaps, torch.full_like(gaps[..., :1], 1e10)], -1)
    gaps = gaps * torch.norm(rays_d[..., None, :], dim=-1)
    colour = torch.sigmoid(raw[..., :3])
    sigma = raw[..., 3]
    if noise is not None:
        sigma = sigma + noise
    elif raw_noise_std > 0.0:
        sigma = sigma + torch.randn(sigma.shape) * raw_noise_std
    alpha = 1.0 - torch.exp(-torch.relu(sigma) * gaps)
    trans = torch.cumprod(torch.cat([torch.ones_like(alpha[..., :1]), 1.0 - alpha + 1e-10], -1), -1)[..., :-1]
    weights = alpha * trans
    rgb_map = torch.sum(weights[..., None] * colour, -2)
    depth_map = torch.sum(weights * z_vals, -1)
    acc_map = torch.sum(weights, -1)
    disp_map = 1.0 / torch.max(1e-10 * torch.ones_like(depth_map), depth_map / acc_map)
    if white_bkgd:
        rgb_map = rgb_map + (1.0 - acc_map[..., None])
    return rgb_map, disp_map, acc_map, weights, depth_map


# --------------------------------------------------------------------------
# a11  run_network / render_rays (upstream) -- UNPINNED
# --------------------------------------------------------------------------

def run_network(pts: torch.Tensor, viewdirs: Optional[torch.Tensor], params, *, L_pts: int = 10,
                L_dirs: int = 4, D: int = 8, skips=(4,), bf16_operands: bool = False) -> torch.Tensor:
    """pts [R,S,3] (+ viewdirs [R,3]) -> raw [R,S,4].  ``bf16_operands``
    emulates the kernel's arithmetic (bf16 activations and weights, fp32
    accumulation) so the 2e-2 tolerance can be tightened in tests."""
    R, S, _ = pts.shape
    enc = posenc(pts.reshape(-1, 3), L_pts)
    views_ch = 0
    if viewdirs is not None:
        dirs = viewdirs[:, None, :].expand(R, S, 3).reshape(-1, 3)
        enc = torch.cat([enc, posenc(dirs, L_dirs)], -1)
        views_ch = posenc_out_dim(3, L_dirs)
    if bf16_operands:
        out = mlp_forward_bf16(params, enc, D=D, skips=skips, input_ch_views=views_ch)
    else:
        out = mlp_forward(params, enc, D=D, skips=skips, input_ch_views=views_ch)
    return out.reshape(R, S, -1)


def _q(t: torch.Tensor) -> torch.Tensor:
    return t.to(torch.bfloat16).to(torch.float32)


def mlp_forward_bf16(p, x, D=8, skips=(4,), input_ch_views=0):
    """mlp_forward with every matmul operand rounded to bf16, fp32 accumulate --
    the arithmetic of the tcgen05 kernel.  The biases of the GEMM layers are bf16
    too (the 2-CTA kernel feeds them through the MMA on a constant-1 input
    channel); the small heads (alpha, rgb, output) keep fp32 biases."""
    if input_ch_views > 0:
        pts, views = x[..., :-input_ch_views], x[..., -input_ch_views:]
    else:
        pts, views = x, None
    pts_q = _q(pts)
    h = pts_q
    for l in range(D):
        h = torch.relu(h @ _q(p[f"pts_linears.{l}.weight"]).T + _q(p[f"pts_linears.{l}.bias"]))
        h = _q(h)
        if l in skips:
            h = torch.cat([pts_q, h], -1)
    if views is None:
        return h @ _q(p["output_linear.weight"]).T + p["output_linear.bias"]
    alpha = h @ _q(p["alpha_linear.weight"]).T + p["alpha_linear.bias"]
    feat = _q(h @ _q(p["feature_linear.weight"]).T + _q(p["feature_linear.bias"]))
    hv = torch.cat([feat, _q(views)], -1)
    hv = _q(torch.relu(hv @ _q(p["views_linears.0.weight"]).T + _q(p["views_linears.0.bias"])))
    rgb = hv @ _q(p["rgb_linear.weight"]).T + p["rgb_linear.bias"]
    return torch.cat([rgb, alpha], -1)


def render_rays(ray_batch: torch.Tensor, network_fn, network_query_fn: Callable, N_samples: int,
                retraw: bool = False, lindisp: bool = False, perturb: float = 0.0,
                N_importance: int = 0, network_fine=None, white_bkgd: bool = False,
                raw_noise_std: float = 0.0, jitter: Optional[torch.Tensor] = None,
                u: Optional[torch.Tensor] = None) -> Dict[str, torch.Tensor]:
    """Coarse + fine render of a ray batch [R, 8 or 11] (upstream contract).
    ``jitter``/``u`` let a test inject the random numbers the CUDA path used."""
    rays_o, rays_d = ray_batch[:, 0:3], ray_batch[:, 3:6]
    viewdirs = ray_batch[:, -3:] if ray_batch.shape[-1] > 8 else None
    near, far = ray_batch[:, 6:7], ray_batch[:, 7:8]
    R = ray_batch.shape[0]
    if perturb > 0.0 and jitter is None:
        jitter = torch.rand(R, N_samples)
    z = stratified_z(near, far, N_samples, lindisp, jitter if perturb > 0.0 else None)
    pts = rays_o[:, None, :] + rays_d[:, None, :] * z[..., None]
    raw = network_query_fn(pts, viewdirs, network_fn)
    rgb, disp, acc, weights, depth = raw2outputs(raw, z, rays_d, raw_noise_std, white_bkgd)
    out: Dict[str, torch.Tensor] = {}
    if N_importance > 0:
        out["rgb0"], out["disp0"], out["acc0"] = rgb, disp, acc
        z_mid = 0.5 * (z[..., 1:] + z[..., :-1])
        z_new = sample_pdf(z_mid, weights[..., 1:-1], N_importance, det=(perturb == 0.0), u=u).detach()
        z, _ = torch.sort(torch.cat([z, z_new], -1), -1)
        pts = rays_o[:, None, :] + rays_d[:, None, :] * z[..., None]
        raw = network_query_fn(pts, viewdirs, network_fine if network_fine is not None else network_fn)
        rgb, disp, acc, weights, depth = raw2outputs(raw, z, rays_d, raw_noise_std, white_bkgd)
        out["z_std"] = torch.std(z_new, dim=-1, unbiased=False)
    out["rgb_map"], out["disp_map"], out["acc_map"] = rgb, disp, acc
    out["depth_map"], out["weights"], out["z_vals"] = depth, weights, z
    if retraw:
        out["raw"] = raw
    return out


def texture_uv_grid(res: int) -> torch.Tensor:
    """UV grid of get_texture_map (/root/reference/src/models/textured_mesh.py:268-272): meshgrid of two
    linspace(0,1,res) with 'xy' indexing, stacked (u,v) and flattened row-major -> [res*res, 2]."""
    lin = torch.linspace(0, 1, res, dtype=torch.float32)
    u, v = torch.meshgrid(lin, lin, indexing="xy")
    return torch.stack([u, v], -1).reshape(-1, 2)


def texture_map(params, res: int, multires: int = 10, bf16_operands: bool = False):
    """get_texture_map (textured_mesh.py:266-301): -> (texture [1,3,res,res], mlp_output [res*res,3])."""
    enc = posenc(texture_uv_grid(res), multires)
    fwd = mlp_forward_bf16 if bf16_operands else mlp_forward
    out = fwd(params, enc)
    tex = (torch.tanh(out) + 1) / 2
    return tex.reshape(1, res, res, 3).permute(0, 3, 1, 2), out


def texture_map_only_valid_areas(params, interpolated_uvs: torch.Tensor, face_idx: torch.Tensor, multires: int = 10,
                                 bf16_operands: bool = False) -> torch.Tensor:
    """get_texture_map_only_valid_areas from the rasteriser's outputs on (textured_mesh.py:328-347): the MLP at the
    covered texels only, colours scaled by 0.8/0.5 (unscale_image :336-338), zeros elsewhere -> [1,3,res,res]."""
    mask = (face_idx >= 0).squeeze(0)
    uvs = interpolated_uvs.squeeze(0)[mask]
    res = mask.shape[-1]
    final = torch.zeros(mask.shape[0], res, 3)
    if uvs.shape[0] > 0:
        fwd = mlp_forward_bf16 if bf16_operands else mlp_forward
        out = fwd(params, posenc(uvs, multires))
        final[mask] = out / 0.5 * 0.8
    return final.permute(2, 0, 1).unsqueeze(0)


def texture_mapping(uv: torch.Tensor, texture: torch.Tensor, mode: str = "bilinear") -> torch.Tensor:
    """kaolin.render.mesh.texture_mapping as called at /root/reference/src/models/render.py:135 (kaolin is third
    party, un-vendored, no version pinned -> parity unpinned; this restates its documented behaviour): uv [B,...,2] in
    [0,1] with v pointing up, texture [B,C,H,W]; uv -> grid_sample coordinates (2u-1, -(2v-1)),
    align_corners=False, padding_mode='border'; returns [B,...,C]."""
    B = uv.shape[0]
    dims = uv.shape[1:-1]
    g = uv.reshape(B, -1, 1, 2) * 2.0 - 1.0
    g = torch.stack([g[..., 0], -g[..., 1]], -1)
    out = torch.nn.functional.grid_sample(texture, g, mode=mode, align_corners=False, padding_mode="border")
    return out.permute(0, 2, 3, 1).reshape(B, *dims, texture.shape[1])


def render_composite(uv, texture, mask, background: float = 1.0, mode: str = "bilinear"):
    """render.py:133-140: texture lookup, times the coverage mask, plus a constant background outside it."""
    img = texture_mapping(uv, texture, mode) * mask
    return img + background * (1 - mask)


# --------------------------------------------------------------------------
# view-weight masks              reference: src/training/trainer.py:155-249
# --------------------------------------------------------------------------

def create_face_view_map(face_idx: torch.Tensor) -> torch.Tensor:
    """trainer.py:155-216: rows (face, view, i, j) of every pixel whose face id is >= 0, (view, pixel) order."""
    V, _, H, W = face_idx.shape
    flat = face_idx.reshape(V, -1)
    view, pix = torch.meshgrid(torch.arange(V), torch.arange(H * W), indexing="ij")
    rows = torch.stack([flat.flatten(), view.flatten(), pix.flatten() // W, pix.flatten() % W], dim=1)
    return rows[flat.flatten() >= 0]


def compare_face_normals_between_views(face_view_map: torch.Tensor, face_normals: torch.Tensor,
                                       face_idx: torch.Tensor) -> torch.Tensor:
    """trainer.py:218-249.  torch_scatter.scatter_max(src, index, dim=0) (:227, third party, not in this image) is
    restated with Tensor.scatter_reduce_('amax', include_self=False): the per-index maximum of the source rows."""
    V, _, H, W = face_idx.shape
    masks = torch.full((V, 1, H, W), True, dtype=torch.bool)
    face_ids, views = face_view_map[:, 0], face_view_map[:, 1]
    i, j = face_view_map[:, 2], face_view_map[:, 3]
    z = face_normals[views, 2, face_ids]
    n_faces = int(face_ids.max().item()) + 1 if face_ids.numel() else 0
    max_z = torch.full((n_faces,), float("-inf")).scatter_reduce_(0, face_ids, z, "amax", include_self=False)
    unworthy = z < max_z[face_ids]
    masks[views, 0, i, j] = ~unworthy
    return masks


def img2mse(x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
    """reference :9"""
    return torch.mean((x - y) ** 2)


# --------------------------------------------------------------------------
# synthetic workloads named by BASELINE.json
# --------------------------------------------------------------------------

def lego_like_camera(H: int = 800, W: int = 800, focal: float = 1111.1, radius: float = 4.0311,
                     elev_deg: float = 30.0, azim_deg: float = 0.0):
    """cfg 2 camera: intrinsics K and a camera-to-world [3,4] looking at the
    origin from (radius, elevation, azimuth), OpenGL axes (x right, y up, -z view)."""
    K = [[focal, 0.0, 0.5 * W], [0.0, focal, 0.5 * H], [0.0, 0.0, 1.0]]
    e, a = math.radians(elev_deg), math.radians(azim_deg)
    eye = np.array([radius * math.cos(e) * math.sin(a), radius * math.sin(e),
                    radius * math.cos(e) * math.cos(a)], dtype=np.float64)
    fwd = -eye / np.linalg.norm(eye)
    right = np.cross(fwd, np.array([0.0, 1.0, 0.0]))
    right /= np.linalg.norm(right)
    up = np.cross(right, fwd)
    c2w = np.stack([right, up, -fwd, eye], axis=1).astype(np.float32)
    return K, torch.from_numpy(c2w)


def cfg1_inputs(R: int = 4096, S: int = 64, seed: int = 0):
    """BASELINE configs[0]: synthetic raw densities/colours on R rays x S samples."""
    g = torch.Generator().manual_seed(seed)
    z = torch.linspace(2.0, 6.0, S).expand(R, S).contiguous()
    d = torch.randn(R, 3, generator=g)
    d = d / d.norm(dim=-1, keepdim=True)
    raw = torch.randn(R, S, 4, generator=g)
    raw[..., 3] *= 5.0
    return raw, z, d


def cfg1_chain(raw, z, d, raw_fine, n_importance: int = 128):
    """raw2outputs(S) -> sample_pdf(det) -> sort/merge -> raw2outputs(S+Ni)."""
    rgb0, disp0, acc0, w, depth0 = raw2outputs(raw, z, d)
    z_mid = 0.5 * (z[..., 1:] + z[..., :-1])
    zs = sample_pdf(z_mid, w[..., 1:-1], n_importance, det=True)
    z_all, _ = torch.sort(torch.cat([z, zs], -1), -1)
    return (rgb0, disp0, acc0, w, depth0), zs, z_all, raw2outputs(raw_fine, z_all, d)
