"""The coarse+fine NeRF training / rendering step built from the ctxnerf
kernels -- the hot path BASELINE.json measures.

One step = ray generation fused with stratified sampling -> coarse MLP (points
encoded in shared memory) -> raw2outputs -> sample_pdf + merge -> fine MLP ->
raw2outputs -> img2mse(rgb)+img2mse(rgb0) -> hand-written backward of every
stage -> one gradient all-reduce -> Adam.  It is upstream's render()+train
iteration with the same hyper-parameter names; every numeric
stage is a libctxnerf.so kernel, orchestrated here without autograd.
"""
from __future__ import annotations

import ctypes
from typing import Optional

import os

import torch

from . import ops
from ._lib import CtxNerfError, call, ptr, stream_ptr
from .dist import FlatBucket, world
from .mlp import forward_raw
from .mlp_bwd import mlp_backward, mlp_dgrad, mlp_wgrad
from .run_nerf_helpers import NeRF


class NerfTrainer:
    def __init__(self, H, W, K, c2w, near=2.0, far=6.0, N_samples=64, N_importance=128, perturb=1.0,
                 white_bkgd=True, lindisp=False, lr=5e-4, betas=(0.9, 0.999), eps=1e-8, device=None, seed=0,
                 multires=10, multires_views=4):
        self.device = torch.device(device if device is not None else ("cuda", torch.cuda.current_device()))
        self.H, self.W, self.K = H, W, K
        self.c2w = torch.as_tensor(c2w, dtype=torch.float32).to(self.device).contiguous()
        self.near, self.far = float(near), float(far)
        self.N_samples, self.N_importance = int(N_samples), int(N_importance)
        self.perturb, self.white_bkgd, self.lindisp = float(perturb), bool(white_bkgd), bool(lindisp)
        self.lr, self.betas, self.eps = lr, betas, eps
        # identical initial weights on every rank: construct under a fixed seed
        st = torch.random.get_rng_state()
        torch.manual_seed(seed)
        in_pts, in_views = 3 * (1 + 2 * multires), 3 * (1 + 2 * multires_views)
        self.coarse = NeRF(input_ch=in_pts, input_ch_views=in_views).to(self.device)
        self.fine = NeRF(input_ch=in_pts, input_ch_views=in_views).to(self.device)
        torch.random.set_rng_state(st)
        for net in (self.coarse, self.fine):
            net.L_pts, net.L_dirs = multires, multires_views
            net._ensure()
        self.bucket = FlatBucket([self.coarse, self.fine])
        self.exp_avg = torch.zeros_like(self.bucket.flat)
        self.exp_avg_sq = torch.zeros_like(self.bucket.flat)
        self.step_count = 0
        self.rank, self.world_size = world()
        self._loss = torch.zeros(1, device=self.device)
        self.timers = None      # optional dict name -> list[(start_event, end_event)]
        # backward overlap (see step): side stream + SM budget of the coarse chain
        self.overlap_backward = os.environ.get("CTXNERF_OVERLAP", "1") != "0"
        self.side_sms = int(os.environ.get("CTXNERF_SIDE_SMS", "44"))
        self._side = torch.cuda.Stream(device=self.device)
        self._ev0, self._ev1 = torch.cuda.Event(), torch.cuda.Event()

    # ------------------------------------------------------------------ utils
    def _timed(self, name, fn):
        if self.timers is None:
            return fn()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        r = fn()
        b.record()
        self.timers.setdefault(name, []).append((a, b))
        return r

    def _render(self, ray_idx, save_acts, seed, perturb):
        S, Ni = self.N_samples, self.N_importance
        jit = bool(perturb)
        r = ops.raygen(self.H, self.W, self.K, self.c2w, ray_idx=ray_idx, n_samples=S, near=self.near, far=self.far,
                       lindisp=self.lindisp, perturb=jit, seed=seed, want_viewdirs=True)
        o, d, v, z_c = r["rays_o"], r["rays_d"], r["viewdirs"], r["z_vals"]
        R = o.shape[0]
        raw_c, acts_c, Pc, pk_c = self._timed("mlp_fwd_coarse",
                                              lambda: forward_raw(self.coarse, rays=(o, d, v, z_c), save_acts=save_acts))
        comp_c = self._composite(raw_c, z_c, d, R, S)
        zs, z_f = ops.resample_merge(z_c, comp_c[3], Ni, det=not jit, seed=seed + 1)
        raw_f, acts_f, Pf, pk_f = self._timed("mlp_fwd_fine",
                                              lambda: forward_raw(self.fine, rays=(o, d, v, z_f), save_acts=save_acts))
        comp_f = self._composite(raw_f, z_f, d, R, S + Ni)
        return dict(d=d, z_c=z_c, z_f=z_f, raw_c=raw_c, raw_f=raw_f, acts_c=acts_c, acts_f=acts_f, Pc=Pc, Pf=Pf,
                    pk_c=pk_c, pk_f=pk_f, comp_c=comp_c, comp_f=comp_f, R=R)

    def _composite(self, raw, z, d, R, S):
        dev = self.device
        rgb = torch.empty(R, 3, device=dev)
        disp = torch.empty(R, device=dev)
        acc = torch.empty(R, device=dev)
        w = torch.empty(R, S, device=dev)
        depth = torch.empty(R, device=dev)
        call("ctx_composite_fwd", ptr(raw), ptr(z), ptr(d), None, R, S, int(self.white_bkgd), ptr(rgb), ptr(disp),
             ptr(acc), ptr(w), ptr(depth), stream_ptr(dev))
        return rgb, disp, acc, w, depth

    def _composite_bwd(self, raw, z, d, R, S, g_rgb):
        g_raw = torch.empty_like(raw)
        call("ctx_composite_bwd", ptr(raw), ptr(z), ptr(d), None, R, S, int(self.white_bkgd), ptr(g_rgb), None,
             None, None, None, ptr(g_raw), stream_ptr(self.device))
        return g_raw

    # ------------------------------------------------------------------- API
    @torch.no_grad()
    def render(self, ray_idx: Optional[torch.Tensor] = None):
        """Inference render of the given pixels (None = the whole H x W image)."""
        with torch.cuda.device(self.device):
            # evaluation is deterministic, as upstream's render_kwargs_test (perturb=False, raw_noise_std=0):
            # linspace depths and det=True importance sampling; the jitter belongs to step() only
            fwd = self._render(ray_idx, save_acts=False, seed=0, perturb=False)
        rgb, disp, acc, _, depth = fwd["comp_f"]
        return dict(rgb_map=rgb, disp_map=disp, acc_map=acc, depth_map=depth, rgb0=fwd["comp_c"][0])

    @torch.no_grad()
    def render_view(self, H, W, K, c2w, n_samples=192, sphere=None, near=None, far=None):
        """Single-pass render of a whole view with the fine network (BASELINE config 4: per-ray near/far from
        the bounding sphere of the normalised mesh, `n_samples` depths per ray, no hierarchical pass)."""
        dev = self.device
        with torch.cuda.device(dev):
            r = ops.raygen(H, W, K, torch.as_tensor(c2w, dtype=torch.float32).to(dev), n_samples=n_samples,
                           near=self.near if near is None else near, far=self.far if far is None else far,
                           sphere=sphere, want_viewdirs=True)
            raw, _, _, _ = forward_raw(self.fine, rays=(r["rays_o"], r["rays_d"], r["viewdirs"], r["z_vals"]))
            rgb, disp, acc, _, depth = self._composite(raw, r["z_vals"], r["rays_d"], H * W, n_samples)
        return dict(rgb_map=rgb.reshape(H, W, 3), disp_map=disp.reshape(H, W), acc_map=acc.reshape(H, W),
                    depth_map=depth.reshape(H, W))

    @torch.no_grad()
    def step(self, ray_idx: torch.Tensor, target: torch.Tensor, optimizer_step: bool = True) -> torch.Tensor:
        """One training step on this rank's ray batch; returns the loss (device scalar)."""
        dev = self.device
        if not (torch.is_tensor(target) and target.is_cuda and target.device == dev):
            raise CtxNerfError("NerfTrainer.step: target must be a CUDA tensor on the trainer's device "
                               "(use step_from_host for host buffers)")
        target = ops._f32c(target)
        if target.numel() != 3 * ray_idx.numel():
            raise CtxNerfError("NerfTrainer.step: target must hold one rgb triple per ray")
        with torch.cuda.device(dev):
            seed = ops.new_seed() if self.perturb > 0.0 else 0
            self.bucket.zero_grad()
            f = self._render(ray_idx, save_acts=True, seed=seed, perturb=self.perturb > 0.0)
            R, S, Sf = f["R"], self.N_samples, self.N_samples + self.N_importance
            g_rgb = torch.empty(R, 3, device=dev)
            g_rgb0 = torch.empty(R, 3, device=dev)
            call("ctx_mse_fwd_bwd", ptr(f["comp_f"][0]), ptr(f["comp_c"][0]), ptr(target), R * 3, 1.0,
                 ptr(self._loss), ptr(g_rgb), ptr(g_rgb0), stream_ptr(dev))
            g_raw_f = self._composite_bwd(f["raw_f"], f["z_f"], f["d"], R, Sf, g_rgb)
            g_raw_c = self._composite_bwd(f["raw_c"], f["z_c"], f["d"], R, S, g_rgb0)
            sinks_f, sinks_c = self.bucket.sinks_for(self.fine), self.bucket.sinks_for(self.coarse)
            dacts_f = self._timed("mlp_dgrad_fine",
                                  lambda: mlp_dgrad(self.fine, f["pk_f"], f["acts_f"], f["Pf"], g_raw_f))
            if self.overlap_backward:
                # The two networks' backward chains are independent, and their kernels bind on different
                # resources: dgrad on the tensor pipe / epilogue, wgrad on HBM reads.  The coarse chain runs on a
                # side stream inside a 36-SM budget while wgrad of the fine network streams its records through
                # the other 112 SMs; wgrad of the coarse network follows on whatever SMs come free.
                main = torch.cuda.current_stream(dev)
                self._ev0.record(main)
                self._side.wait_event(self._ev0)
                if self.timers is not None:
                    g0 = torch.cuda.Event(enable_timing=True)
                    g0.record(main)
                with torch.cuda.stream(self._side):
                    dacts_c = self._timed("mlp_dgrad_coarse",
                                          lambda: mlp_dgrad(self.coarse, f["pk_c"], f["acts_c"], f["Pc"], g_raw_c,
                                                            max_sms=self.side_sms))
                    self._timed("mlp_wgrad_coarse",
                                lambda: mlp_wgrad(self.coarse, f["acts_c"], dacts_c, f["Pc"], sinks_c))
                    self._ev1.record(self._side)
                self._timed("mlp_wgrad_fine",
                            lambda: mlp_wgrad(self.fine, f["acts_f"], dacts_f, f["Pf"], sinks_f,
                                              max_sms=148 - self.side_sms))
                main.wait_event(self._ev1)
                if self.timers is not None:   # span of the concurrent group on the main stream
                    g1 = torch.cuda.Event(enable_timing=True)
                    g1.record(main)
                    self.timers.setdefault("bwd_overlap_group", []).append((g0, g1))
            else:
                self._timed("mlp_wgrad_fine", lambda: mlp_wgrad(self.fine, f["acts_f"], dacts_f, f["Pf"], sinks_f))
                dacts_c = self._timed("mlp_dgrad_coarse",
                                      lambda: mlp_dgrad(self.coarse, f["pk_c"], f["acts_c"], f["Pc"], g_raw_c))
                self._timed("mlp_wgrad_coarse", lambda: mlp_wgrad(self.coarse, f["acts_c"], dacts_c, f["Pc"], sinks_c))
            self.bucket.all_reduce()
            if optimizer_step:
                self.step_count += 1
                call("ctx_adam_step", ptr(self.bucket.flat), ptr(self.bucket.grad), ptr(self.exp_avg),
                     ptr(self.exp_avg_sq), self.bucket.numel, self.lr, self.betas[0], self.betas[1], self.eps,
                     self.step_count, 0.0, 1.0 / self.world_size, stream_ptr(dev))
                self.coarse._packed.invalidate()
                self.fine._packed.invalidate()
        return self._loss

    def step_from_host(self, ray_idx_pinned: torch.Tensor, target_pinned: torch.Tensor,
                       loss_pinned: torch.Tensor) -> "torch.cuda.Event":
        """End-to-end entry: host (pinned) inputs in, loss back to the host.  Asynchronous: the returned event marks
        the arrival of the loss in ``loss_pinned``, so a training loop can submit step i+1 before it reads the loss
        of step i (the launches of the next step then hide behind the GPU work of this one)."""
        idx = ray_idx_pinned.to(self.device, non_blocking=True)
        tgt = target_pinned.to(self.device, non_blocking=True)
        loss = self.step(idx, tgt)
        loss_pinned.copy_(loss, non_blocking=True)
        done = torch.cuda.Event()
        done.record(torch.cuda.current_stream(self.device))
        return done     # loss_pinned is valid once this event has completed
