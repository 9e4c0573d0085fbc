"""Drop-in replacement for ``src/run_nerf_helpers.py`` of ConTEXTure-NeRF.

Same names, argument meaning and return conventions as
/root/reference/src/run_nerf_helpers.py (Embedder :15, get_embedder :48,
NeRF2D :68, get_rays :139, get_rays_np :151, ndc_rays :161, sample_pdf :182,
img2mse/mse2psnr/to8b :9-11) plus the upstream functions the reference only
points at (:131-133): NeRF (view-direction MLP), raw2outputs, run_network,
render_rays.  Every tensor op runs in libctxnerf.so (hand-written sm_100a
kernels behind the C-ABI of include/ctxnerf.h); there is no eager-PyTorch or
CPU fallback -- inputs living on the CPU are moved to the current CUDA device.

The reference does ``from src.run_nerf_helpers import *`` (trainer.py:33) and
relies on ``torch, nn, F, np`` leaking through, so no ``__all__`` is defined.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F
import numpy as np

from . import ops as _ops
from . import _lib as _clib
from .mlp import FusedMLPBase as _FusedMLPBase


def _cuda_device():
    if not torch.cuda.is_available():
        raise _clib.CtxNerfError("ctxnerf needs a CUDA device (B200); there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def _to_cuda(t, dtype=torch.float32):
    if t is None:
        return None
    if not torch.is_tensor(t):
        t = torch.as_tensor(np.asarray(t), dtype=dtype)
    if not t.is_cuda:
        t = t.to(_cuda_device())
    if t.is_floating_point() and t.dtype != dtype:
        t = t.to(dtype)       # the kernels read float*: a float64 / half batch is converted, never reinterpreted
    return t


# Misc (reference :9-11)
img2mse = lambda x, y: torch.mean((x - y) ** 2)
mse2psnr = lambda x: -10. * torch.log(x) / torch.log(torch.tensor([10.], device=x.device))
to8b = lambda x: (255 * np.clip(x, 0, 1)).astype(np.uint8)


# Positional encoding (reference :15-45)
class Embedder:
    def __init__(self, **kwargs):
        self.kwargs = kwargs
        self.create_embedding_fn()

    def create_embedding_fn(self):
        kw = self.kwargs
        d = kw['input_dims']
        fns = list(kw['periodic_fns'])
        if fns and fns != [torch.sin, torch.cos]:
            raise _clib.CtxNerfError("the fused encoder implements periodic_fns=[torch.sin, torch.cos] only")
        self._L = int(kw['num_freqs']) if fns else 0
        self._inc = bool(kw['include_input'])
        self._log = bool(kw['log_sampling'])
        if self._L > 0 and kw['max_freq_log2'] != self._L - 1:
            raise _clib.CtxNerfError("the fused encoder expects max_freq_log2 == num_freqs - 1")
        self.out_dim = d * ((1 if self._inc else 0) + 2 * self._L)
        # per-block views of the fused result, kept for attribute compatibility
        self.embed_fns = [
            (lambda x, i=i: self.embed(x)[..., i * d:(i + 1) * d]) for i in range(self.out_dim // d)]

    def embed(self, inputs):
        return _ops.posenc(_to_cuda(inputs), self._L, self._inc, self._log)


def get_embedder(multires, i=0, input_dims=2):
    """``input_dims`` is an extension: the reference hard-codes 2 (UV texture
    coordinates, :55-56); the volumetric path passes 3."""
    if i == -1:
        return nn.Identity(), input_dims
    embedder_obj = Embedder(include_input=True, input_dims=input_dims, max_freq_log2=multires - 1,
                            num_freqs=multires, log_sampling=True, periodic_fns=[torch.sin, torch.cos])
    embed = lambda x, eo=embedder_obj: eo.embed(x)
    return embed, embedder_obj.out_dim


# Model (reference :68-135)
class NeRF2D(_FusedMLPBase):
    def __init__(self, D=8, W=256, input_ch=3, output_ch=4, skips=[4]):
        super(NeRF2D, self).__init__()
        self.D = D
        self.W = W
        self.input_ch = input_ch
        self.skips = skips
        self.pts_linears = nn.ModuleList(
            [nn.Linear(input_ch, W)] +
            [nn.Linear(W, W) if i not in self.skips else nn.Linear(W + input_ch, W) for i in range(D - 1)])
        self.output_linear = nn.Linear(W, output_ch)
        for layer in self.pts_linears:
            nn.init.kaiming_normal_(layer.weight, mode='fan_in', nonlinearity='relu')
        nn.init.kaiming_normal_(self.output_linear.weight, mode='fan_in', nonlinearity='relu')
        self._setup(D, W, skips, input_ch, 0, output_ch)

    def _param_list(self):
        ps = []
        for l in self.pts_linears:
            ps += [l.weight, l.bias]
        return ps + [self.output_linear.weight, self.output_linear.bias]

    def forward(self, x):
        return self._run(x=_to_cuda(x))


class NeRF(_FusedMLPBase):
    """Upstream view-direction MLP (the branch kept as comments at reference
    :86-95, :117-127): input [pts_enc | view_enc] -> [rgb(3) | alpha(1)]."""

    def __init__(self, D=8, W=256, input_ch=63, input_ch_views=27, output_ch=4, skips=[4], use_viewdirs=True):
        super(NeRF, self).__init__()
        self.D, self.W, self.input_ch, self.input_ch_views = D, W, input_ch, input_ch_views
        self.skips, self.use_viewdirs = skips, use_viewdirs
        self.pts_linears = nn.ModuleList(
            [nn.Linear(input_ch, W)] +
            [nn.Linear(W, W) if i not in self.skips else nn.Linear(W + input_ch, W) for i in range(D - 1)])
        if use_viewdirs:
            self.feature_linear = nn.Linear(W, W)
            self.alpha_linear = nn.Linear(W, 1)
            self.views_linears = nn.ModuleList([nn.Linear(input_ch_views + W, W // 2)])
            self.rgb_linear = nn.Linear(W // 2, 3)
        else:
            self.output_linear = nn.Linear(W, output_ch)
        for layer in self.pts_linears:
            nn.init.kaiming_normal_(layer.weight, mode='fan_in', nonlinearity='relu')
        self._setup(D, W, skips, input_ch, input_ch_views if use_viewdirs else 0, 4 if use_viewdirs else output_ch)

    def _param_list(self):
        ps = []
        for l in self.pts_linears:
            ps += [l.weight, l.bias]
        if self.use_viewdirs:
            for m in (self.feature_linear, self.alpha_linear, self.views_linears[0], self.rgb_linear):
                ps += [m.weight, m.bias]
        else:
            ps += [self.output_linear.weight, self.output_linear.bias]
        return ps

    def forward(self, x):
        x = _to_cuda(x)
        if not self.use_viewdirs:
            x = x[..., :self.input_ch]
        return self._run(x=x)

    def forward_rays(self, rays_o, rays_d, viewdirs, z_vals):
        """Fused query: points o + d*z are formed and encoded inside the MLP
        kernel (nothing but raw [R,S,4] touches HBM)."""
        v = viewdirs if self.use_viewdirs else None
        f = _ops._f32c
        _ops._need_cuda(rays_o, rays_d, v, z_vals)
        return self._run(rays=(f(rays_o), f(rays_d), f(v), f(z_vals)))


# Ray helpers (reference :139-178)
def get_rays(H, W, K, c2w):
    c2w = _to_cuda(c2w)
    r = _ops.raygen(H, W, K, c2w)
    rays_d = r["rays_d"].reshape(H, W, 3)
    rays_o = c2w[:3, -1].to(torch.float32).expand(rays_d.shape)   # stride-0 view, as the reference returns
    return rays_o, rays_d


def get_rays_np(H, W, K, c2w):
    rays_o, rays_d = get_rays(H, W, K, torch.as_tensor(np.asarray(c2w), dtype=torch.float32))
    return np.broadcast_to(np.asarray(c2w, dtype=np.float32)[:3, -1], (H, W, 3)), rays_d.cpu().numpy()


def ndc_rays(H, W, focal, near, rays_o, rays_d):
    rays_d = _to_cuda(rays_d)
    rays_o = _to_cuda(rays_o).expand(rays_d.shape)
    return _ops.ndc(H, W, focal, near, rays_o, rays_d)


# Hierarchical sampling (reference :182-225)
def sample_pdf(bins, weights, N_samples, det=False, pytest=False):
    bins, weights = _to_cuda(bins), _to_cuda(weights)
    u = None
    if pytest:  # reference :196-205: numpy's fixed random numbers
        np.random.seed(0)
        new_shape = list(weights.shape[:-1]) + [N_samples]
        if det:
            u = np.broadcast_to(np.linspace(0., 1., N_samples), new_shape)
        else:
            u = np.random.rand(*new_shape)
        u = torch.Tensor(np.ascontiguousarray(u)).to(bins.device)
    return _ops.resample(bins, weights, N_samples, det=det, u=u)


# Volume compositing (upstream raw2outputs; reference comment :131-133)
def raw2outputs(raw, z_vals, rays_d, raw_noise_std=0, white_bkgd=False, pytest=False):
    raw, z_vals, rays_d = _to_cuda(raw), _to_cuda(z_vals), _to_cuda(rays_d)
    noise = None
    if raw_noise_std > 0.:
        if pytest:
            np.random.seed(0)
            noise = torch.Tensor(np.random.rand(*list(raw[..., 3].shape)) * raw_noise_std).to(raw.device)
        else:
            noise = torch.randn(raw[..., 3].shape, device=raw.device) * raw_noise_std
    rgb_map, disp_map, acc_map, weights, depth_map = _ops.composite(raw, z_vals, rays_d, noise, white_bkgd)
    return rgb_map, disp_map, acc_map, weights, depth_map


def batchify(fn, chunk):
    """Upstream helper; the fused kernel tiles internally so chunk is only honoured for foreign fns."""
    if chunk is None:
        return fn
    return lambda inputs: torch.cat([fn(inputs[i:i + chunk]) for i in range(0, inputs.shape[0], chunk)], 0)


def run_network(inputs, viewdirs, fn, embed_fn, embeddirs_fn, netchunk=1024 * 64):
    """Upstream run_network: embed points (+ expanded view directions), apply the MLP."""
    inputs_flat = torch.reshape(inputs, [-1, inputs.shape[-1]])
    embedded = embed_fn(inputs_flat)
    if viewdirs is not None:
        input_dirs = viewdirs[:, None].expand(inputs.shape)
        input_dirs_flat = torch.reshape(input_dirs, [-1, input_dirs.shape[-1]])
        embedded = torch.cat([embedded, embeddirs_fn(input_dirs_flat)], -1)
    chunked = isinstance(fn, _FusedMLPBase)
    outputs_flat = fn(embedded) if chunked else batchify(fn, netchunk)(embedded)
    return torch.reshape(outputs_flat, list(inputs.shape[:-1]) + [outputs_flat.shape[-1]])


class FusedQuery:
    """``network_query_fn`` marker: tells render_rays that ``network_fn`` is a
    ctxnerf ``NeRF`` and that points should be generated/encoded inside the MLP
    kernel.  Calling it like upstream's query fn also works."""

    def __init__(self, multires=10, multires_views=4):
        self.multires, self.multires_views = multires, multires_views
        self.embed_fn, _ = get_embedder(multires, 0, input_dims=3)
        self.embeddirs_fn, _ = get_embedder(multires_views, 0, input_dims=3)

    def __call__(self, pts, viewdirs, network_fn):
        use_dirs = viewdirs if getattr(network_fn, "use_viewdirs", True) else None
        return run_network(pts, use_dirs, network_fn, self.embed_fn, self.embeddirs_fn)


def render_rays(ray_batch, network_fn, network_query_fn, N_samples, retraw=False, lindisp=False, perturb=0.,
                N_importance=0, network_fine=None, white_bkgd=False, raw_noise_std=0., verbose=False,
                pytest=False):
    """Upstream render_rays contract: coarse pass, importance
    resampling, fine pass.  Returns dict(rgb_map, disp_map, acc_map[, raw]
    [, rgb0, disp0, acc0, z_std])."""
    ray_batch = _to_cuda(ray_batch)
    N_rays = ray_batch.shape[0]
    rays_o, rays_d = ray_batch[:, 0:3], ray_batch[:, 3:6]
    viewdirs = ray_batch[:, -3:] if ray_batch.shape[-1] > 8 else None
    jitter = None
    if perturb > 0. and pytest:
        np.random.seed(0)
        jitter = torch.Tensor(np.random.rand(N_rays, N_samples)).to(ray_batch.device)
    z_vals = _ops.stratified(ray_batch[:, 6], ray_batch[:, 7], N_samples, lindisp=lindisp,
                             perturb=perturb > 0., jitter=jitter)
    fused = isinstance(network_query_fn, FusedQuery) and isinstance(network_fn, NeRF)
    ro_c, rd_c = _ops._f32c(rays_o), _ops._f32c(rays_d)
    vd_c = _ops._f32c(viewdirs)

    def query(z, net):
        if fused and isinstance(net, NeRF) and (vd_c is not None or not net.use_viewdirs):
            return net.forward_rays(ro_c, rd_c, vd_c, z)
        pts = rays_o[..., None, :] + rays_d[..., None, :] * z[..., :, None]
        return network_query_fn(pts, viewdirs, net)

    raw = query(z_vals, network_fn)
    rgb_map, disp_map, acc_map, weights, depth_map = raw2outputs(raw, z_vals, rays_d, raw_noise_std, white_bkgd,
                                                                 pytest=pytest)
    if N_importance > 0:
        rgb_map_0, disp_map_0, acc_map_0 = rgb_map, disp_map, acc_map
        u = None
        if pytest:
            np.random.seed(0)
            shp = (N_rays, N_importance)
            u = np.broadcast_to(np.linspace(0., 1., N_importance), shp) if perturb == 0. else np.random.rand(*shp)
            u = torch.Tensor(np.ascontiguousarray(u)).to(ray_batch.device)
        # fused: mid-points, pdf of weights[...,1:-1], inverse CDF, merge + sort (detached, as upstream)
        z_samples, z_vals = _ops.resample_merge(z_vals, weights, N_importance, det=(perturb == 0.), u=u)
        run_fn = network_fn if network_fine is None else network_fine
        raw = query(z_vals, run_fn)
        rgb_map, disp_map, acc_map, weights, depth_map = raw2outputs(raw, z_vals, rays_d, raw_noise_std,
                                                                     white_bkgd, pytest=pytest)
    ret = {'rgb_map': rgb_map, 'disp_map': disp_map, 'acc_map': acc_map}
    if retraw:
        ret['raw'] = raw
    if N_importance > 0:
        ret['rgb0'] = rgb_map_0
        ret['disp0'] = disp_map_0
        ret['acc0'] = acc_map_0
        ret['z_std'] = torch.std(z_samples, dim=-1, unbiased=False)
    return ret


# ------------------------------------------------------------------------------------------------------------
# Upstream outer driver: nerf-pytorch's run_nerf.py render / batchify_rays / render_path, the
# callers of render_rays that the reference's pointer comment (src/run_nerf_helpers.py:131-133) refers to.  The
# module is third party and un-vendored, so these follow its published call contract (argument names, defaults,
# return order); nothing here has a counterpart to check against inside /root/reference.
def batchify_rays(rays_flat, chunk=1024 * 32, **kwargs):
    """render_rays over ``chunk``-sized slices of a [N, 8|11] ray batch; dict of concatenated outputs."""
    parts = {}
    for i in range(0, rays_flat.shape[0], chunk):
        ret = render_rays(rays_flat[i:i + chunk], **kwargs)
        for k, v in ret.items():
            parts.setdefault(k, []).append(v)
    return {k: torch.cat(v, 0) for k, v in parts.items()}


def render(H, W, K, chunk=1024 * 32, rays=None, c2w=None, ndc=True, near=0., far=1., use_viewdirs=False,
           c2w_staticcam=None, **kwargs):
    """Render a full image (``c2w`` given) or an explicit ray batch (``rays`` = (rays_o, rays_d)).
    Returns ``[rgb_map, disp_map, acc_map, extras]`` with the maps shaped like the input rays."""
    if c2w is not None:
        rays_o, rays_d = get_rays(H, W, K, c2w)
    else:
        rays_o, rays_d = rays
    rays_o, rays_d = _to_cuda(rays_o), _to_cuda(rays_d)
    viewdirs = None
    if use_viewdirs:
        viewdirs = rays_d
        if c2w_staticcam is not None:   # view directions of `c2w`, geometry of the static camera
            rays_o, rays_d = get_rays(H, W, K, c2w_staticcam)
        viewdirs = viewdirs / torch.norm(viewdirs, dim=-1, keepdim=True)
        viewdirs = viewdirs.reshape(-1, 3).float()
    sh = rays_d.shape
    if ndc:
        rays_o, rays_d = ndc_rays(H, W, K[0][0], 1., rays_o, rays_d)
    rays_o = rays_o.reshape(-1, 3).float()
    rays_d = rays_d.reshape(-1, 3).float()
    near_t = near * torch.ones_like(rays_d[..., :1])
    far_t = far * torch.ones_like(rays_d[..., :1])
    rays_cat = torch.cat([rays_o, rays_d, near_t, far_t], -1)
    if use_viewdirs:
        rays_cat = torch.cat([rays_cat, viewdirs], -1)
    all_ret = batchify_rays(rays_cat, chunk, **kwargs)
    for k in all_ret:
        all_ret[k] = all_ret[k].reshape(list(sh[:-1]) + list(all_ret[k].shape[1:]))
    k_extract = ['rgb_map', 'disp_map', 'acc_map']
    ret_list = [all_ret[k] for k in k_extract]
    ret_dict = {k: all_ret[k] for k in all_ret if k not in k_extract}
    return ret_list + [ret_dict]


def render_path(render_poses, hwf, K, chunk, render_kwargs, gt_imgs=None, savedir=None, render_factor=0):
    """Render one image per pose; returns (rgbs, disps) as stacked numpy arrays.  (Writing PNGs needs imageio,
    which this image does not have: ``savedir`` is accepted and ignored unless imageio imports.)"""
    H, W, focal = hwf
    if render_factor != 0:
        H, W, focal = H // render_factor, W // render_factor, focal / render_factor
        K = [[K[0][0] / render_factor, 0., K[0][2] / render_factor],
             [0., K[1][1] / render_factor, K[1][2] / render_factor], [0., 0., 1.]]
    rgbs, disps = [], []
    for i, c2w in enumerate(render_poses):
        with torch.no_grad():
            rgb, disp, acc, _ = render(H, W, K, chunk=chunk, c2w=torch.as_tensor(c2w)[:3, :4], **render_kwargs)
        rgbs.append(rgb.cpu().numpy())
        disps.append(disp.cpu().numpy())
        if savedir is not None:
            try:
                import imageio
                imageio.imwrite(f"{savedir}/{i:03d}.png", to8b(rgbs[-1]))
            except ImportError:
                pass
    return np.stack(rgbs, 0), np.stack(disps, 0)
