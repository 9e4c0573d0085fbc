"""Second, independent restatement of upstream ``raw2outputs`` in numpy float64.
TEST INFRASTRUCTURE ONLY (same rules as nerf_oracle.py: imported by tests/ only).

Why it exists: ``raw2outputs`` is absent from /root/reference (only the pointer
comment at src/run_nerf_helpers.py:131-133), so ``nerf_oracle.raw2outputs`` --
a vectorised torch restatement -- is **unpinned**.  This file states the same
published equations (Mildenhall et al., NeRF, eq. 3; upstream
yenchenlin/nerf-pytorch ``run_nerf.py: raw2outputs``) a second
time, written differently on purpose: scalar loops over rays and samples, a
running transmittance instead of ``cumprod``, float64 throughout, and a
hand-derived reverse-mode gradient instead of autograd.  The two restatements
share no code; tests/test_oracle_known_answers.py checks them against each
other (forward to fp32 rounding, backward against fp64 autograd of the torch
oracle).  Agreement of two independent derivations is weaker than a golden
vector from the reference, and DESIGN.md keeps saying "parity unpinned" for
this function.
"""
from __future__ import annotations

import math

import numpy as np


def raw2outputs_fp64(raw, z_vals, rays_d, white_bkgd=False, noise=None):
    """raw [R,S,4], z_vals [R,S], rays_d [R,3] (array-likes) ->
    dict(rgb [R,3], disp [R], acc [R], weights [R,S], depth [R]) in float64."""
    raw = np.asarray(raw, dtype=np.float64)
    z = np.asarray(z_vals, dtype=np.float64)
    d = np.asarray(rays_d, dtype=np.float64)
    R, S, _ = raw.shape
    rgb = np.zeros((R, 3)); disp = np.zeros(R); acc = np.zeros(R); depth = np.zeros(R)
    w = np.zeros((R, S))
    for r in range(R):
        norm = math.sqrt(d[r, 0] ** 2 + d[r, 1] ** 2 + d[r, 2] ** 2)
        T = 1.0                                         # transmittance in front of sample s
        for s in range(S):
            delta = (z[r, s + 1] - z[r, s]) if s + 1 < S else 1e10
            delta *= norm
            sigma = raw[r, s, 3] + (0.0 if noise is None else float(noise[r, s]))
            alpha = 1.0 - math.exp(-max(sigma, 0.0) * delta)
            w[r, s] = alpha * T
            T *= (1.0 - alpha + 1e-10)
            for c in range(3):
                rgb[r, c] += w[r, s] / (1.0 + math.exp(-raw[r, s, c]))
            depth[r] += w[r, s] * z[r, s]
            acc[r] += w[r, s]
        if acc[r] == 0.0:
            disp[r] = float("nan")                      # 0/0 -> NaN survives torch.max
        else:
            disp[r] = 1.0 / max(1e-10, depth[r] / acc[r])
        if white_bkgd:
            rgb[r] += 1.0 - acc[r]
    return dict(rgb=rgb, disp=disp, acc=acc, weights=w, depth=depth)


def raw2outputs_fp64_backward(raw, z_vals, rays_d, g_rgb, g_acc, g_weights, g_depth, white_bkgd=False):
    """Hand-derived gradient of  sum(rgb*g_rgb) + sum(acc*g_acc) + sum(weights*g_weights) + sum(depth*g_depth)
    with respect to raw, float64 (disp is left out: it is a function of depth and acc).

    With t_s = 1 - alpha_s + 1e-10, T_s = prod_{k<s} t_k, w_s = alpha_s T_s and
    G_s = dL/dw_s = g_w[s] + sum_c g_rgb[c] sigmoid(raw[s,c]) + g_depth z_s + g_acc (- sum_c g_rgb[c] if white):
        dL/dalpha_s = T_s G_s - sum_{j>s} G_j alpha_j T_j / t_s
    computed here with the explicit O(S^2) double loop (no scan, no division by t_s: the product over
    k in (s, j) is rebuilt), which is the definition, not the kernel's algorithm."""
    raw = np.asarray(raw, dtype=np.float64)
    z = np.asarray(z_vals, dtype=np.float64)
    d = np.asarray(rays_d, dtype=np.float64)
    R, S, _ = raw.shape
    g_raw = np.zeros_like(raw)
    for r in range(R):
        norm = math.sqrt(float(np.dot(d[r], d[r])))
        delta = np.empty(S); alpha = np.empty(S); t = np.empty(S); T = np.empty(S); col = np.empty((S, 3))
        run = 1.0
        for s in range(S):
            delta[s] = ((z[r, s + 1] - z[r, s]) if s + 1 < S else 1e10) * norm
            alpha[s] = 1.0 - math.exp(-max(raw[r, s, 3], 0.0) * delta[s])
            t[s] = 1.0 - alpha[s] + 1e-10
            T[s] = run
            run *= t[s]
            for c in range(3):
                col[s, c] = 1.0 / (1.0 + math.exp(-raw[r, s, c]))
        ga = float(g_acc[r]) - (float(np.sum(g_rgb[r])) if white_bkgd else 0.0)
        G = np.array([float(g_weights[r, s]) + float(np.dot(g_rgb[r], col[s])) + float(g_depth[r]) * z[r, s] + ga
                      for s in range(S)])
        for s in range(S):
            behind = 0.0
            for j in range(s + 1, S):
                between = 1.0
                for k in range(s + 1, j):
                    between *= t[k]
                behind += G[j] * alpha[j] * T[s] * between       # d w_j / d t_s  (t_s itself left out)
            g_alpha = T[s] * G[s] - behind
            # d alpha / d sigma = delta * exp(-sigma delta) for sigma > 0
            g_raw[r, s, 3] = g_alpha * delta[s] * math.exp(-raw[r, s, 3] * delta[s]) if raw[r, s, 3] > 0 else 0.0
            wv = alpha[s] * T[s]
            for c in range(3):
                g_raw[r, s, c] = wv * float(g_rgb[r, c]) * col[s, c] * (1.0 - col[s, c])
    return g_raw
