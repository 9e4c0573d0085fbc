"""Generate the committed golden fixtures from the UNMODIFIED reference.

Run in the build container only (it needs /root/reference, which does not
exist on the GPU box):

    python tests/golden/make_golden.py

It imports /root/reference/src/run_nerf_helpers.py as is, runs every function
of the hot path that exists there on small seeded inputs, and stores inputs
and outputs in tests/golden/ref_*.npz.  The tests compare the oracle (and, on
the GPU, the CUDA kernels) with these files.
"""
import importlib.util
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/src/run_nerf_helpers.py"


def load_reference():
    spec = importlib.util.spec_from_file_location("ref_run_nerf_helpers", REF)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def main():
    ref = load_reference()
    torch.manual_seed(1234)
    out = {}

    # ---- Embedder: d=2/L=10 (live texture path), d=3/L=10, d=3/L=4 ------------
    for tag, d, L in (("uv", 2, 10), ("pts", 3, 10), ("dirs", 3, 4)):
        x = (torch.rand(48, d) * 4 - 2)
        eo = ref.Embedder(include_input=True, input_dims=d, max_freq_log2=L - 1, num_freqs=L,
                          log_sampling=True, periodic_fns=[torch.sin, torch.cos])
        out[f"emb_{tag}_x"] = x.numpy()
        out[f"emb_{tag}_y"] = eo.embed(x).numpy()
        assert eo.out_dim == d * (1 + 2 * L)
    fn, od = ref.get_embedder(10)
    out["get_embedder10_outdim"] = np.array(od)
    out["get_embedder10_y"] = fn(torch.from_numpy(out["emb_uv_x"])).numpy()

    # ---- NeRF2D 63->4 and 42->3 ------------------------------------------------
    # (W=64 keeps the committed fixture small; layer structure, skip position and
    #  parameter order are those of the W=256 net)
    for tag, cin, cout in (("vol", 63, 4), ("tex", 42, 3)):
        torch.manual_seed(7)
        net = ref.NeRF2D(D=8, W=64, input_ch=cin, output_ch=cout, skips=[4])
        x = torch.randn(32, cin)
        y = net(x)
        (y ** 2).sum().backward()
        out[f"mlp_{tag}_x"] = x.numpy()
        out[f"mlp_{tag}_y"] = y.detach().numpy()
        for name, prm in net.named_parameters():
            out[f"mlp_{tag}_p/{name}"] = prm.detach().numpy()
            out[f"mlp_{tag}_g/{name}"] = prm.grad.numpy()

    # ---- get_rays / get_rays_np -----------------------------------------------
    H, W = 12, 20
    K = [[23.5, 0.0, 9.75], [0.0, 24.25, 6.5], [0.0, 0.0, 1.0]]
    ang = 0.7
    c2w = torch.tensor([[np.cos(ang), 0.1, np.sin(ang), 1.5],
                        [0.05, 0.98, -0.2, -0.25],
                        [-np.sin(ang), 0.15, np.cos(ang), 3.0]], dtype=torch.float32)
    ro, rd = ref.get_rays(H, W, K, c2w)
    out["rays_K"] = np.array(K, dtype=np.float64)
    out["rays_c2w"] = c2w.numpy()
    out["rays_HW"] = np.array([H, W])
    out["rays_o"] = ro.contiguous().numpy()
    out["rays_d"] = rd.numpy()
    ro_np, rd_np = ref.get_rays_np(H, W, np.array(K, dtype=np.float32), c2w.numpy())
    out["rays_np_d"] = rd_np.astype(np.float32)

    # ---- ndc_rays --------------------------------------------------------------
    o_in = torch.randn(40, 3) * 0.3
    d_in = torch.randn(40, 3)
    d_in[:, 2] = -d_in[:, 2].abs() - 0.2
    on, dn = ref.ndc_rays(378, 504, 407.5, 1.0, o_in, d_in)
    out["ndc_in_o"], out["ndc_in_d"] = o_in.numpy(), d_in.numpy()
    out["ndc_o"], out["ndc_d"] = on.numpy(), dn.numpy()
    out["ndc_args"] = np.array([378, 504, 407.5, 1.0])

    # ---- sample_pdf: det, explicit-u, all-zero and one-hot weights ------------
    R, B, N = 24, 63, 128
    bins = torch.sort(torch.rand(R, B) * 4 + 2, -1)[0]
    w = torch.rand(R, B - 1)
    w[3] = 0.0                       # all-zero weights -> uniform
    w[4] = 0.0
    w[4, 17] = 1.0                   # one-hot -> denom<1e-5 branch
    w[5] *= (torch.rand(B - 1) < 0.2).float()   # sparse
    # the reference's own cdf (stage 1), to pin stage 2 bit for bit
    wp = w + 1e-5
    pdf = wp / torch.sum(wp, -1, keepdim=True)
    cdf = torch.cat([torch.zeros(R, 1), torch.cumsum(pdf, -1)], -1)
    s_det = ref.sample_pdf(bins, w, N, det=True)
    out["pdf_bins"], out["pdf_w"], out["pdf_cdf_ref"] = bins.numpy(), w.numpy(), cdf.numpy()
    out["pdf_det"] = s_det.numpy()
    u_det = torch.linspace(0., 1., N).expand(R, N).contiguous()
    out["pdf_det_inds"] = torch.searchsorted(cdf, u_det, right=True).numpy()
    out["pdf_linspace128"] = torch.linspace(0., 1., N).numpy()
    for n in (64, 192, 256, 512, 7):
        out[f"pdf_linspace{n}"] = torch.linspace(0., 1., n).numpy()
    # random u: replay torch.rand through the global RNG
    torch.manual_seed(99)
    s_rand = ref.sample_pdf(bins, w, N, det=False)
    torch.manual_seed(99)
    u_rand = torch.rand(R, N)
    out["pdf_u_rand"], out["pdf_rand"] = u_rand.numpy(), s_rand.numpy()
    out["pdf_rand_inds"] = torch.searchsorted(cdf, u_rand, right=True).numpy()
    # non-contiguous weights slice as render_rays passes it
    wfull = torch.rand(R, B + 1)
    out["pdf_wfull"] = wfull.numpy()
    out["pdf_det_slice"] = ref.sample_pdf(bins, wfull[..., 1:-1], N, det=True).numpy()

    # ---- misc ------------------------------------------------------------------
    a, b = torch.rand(16, 3), torch.rand(16, 3)
    out["mse_a"], out["mse_b"] = a.numpy(), b.numpy()
    out["mse"] = ref.img2mse(a, b).numpy()
    out["psnr"] = ref.mse2psnr(ref.img2mse(a, b)).numpy()
    out["to8b"] = ref.to8b(a.numpy() * 1.4 - 0.2)

    np.savez_compressed(os.path.join(HERE, "ref_run_nerf_helpers.npz"), **out)
    print("wrote", os.path.join(HERE, "ref_run_nerf_helpers.npz"), len(out), "arrays")

    # ---- cross-check the oracle against the live reference at larger sizes ----
    sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
    from oracle import nerf_oracle as orc
    torch.manual_seed(5)
    R = 4096
    bins = torch.sort(torch.rand(R, 63) * 4 + 2, -1)[0]
    w = torch.rand(R, 62) ** 4
    wp = w + 1e-5
    cdf = torch.cat([torch.zeros(R, 1), torch.cumsum(wp / torch.sum(wp, -1, keepdim=True), -1)], -1)
    s_ref = ref.sample_pdf(bins, w, 128, det=True)
    s_st2, i_st2 = orc.sample_pdf(bins, w, 128, det=True, cdf=cdf, return_inds=True)
    s_orc, i_orc = orc.sample_pdf(bins, w, 128, det=True, return_inds=True)
    i_ref = torch.searchsorted(cdf, torch.linspace(0, 1, 128).expand(R, 128).contiguous(), right=True)
    print("stage-2 bit-exact vs reference:", bool((s_st2 == s_ref).all()), bool((i_st2 == i_ref).all()))
    print("end-to-end: index agreement %.6f, max |d sample| %.3e" %
          ((i_orc == i_ref).float().mean().item(), (s_orc - s_ref).abs().max().item()))
    x = torch.rand(4096, 3) * 4 - 2
    eo = ref.Embedder(include_input=True, input_dims=3, max_freq_log2=9, num_freqs=10,
                      log_sampling=True, periodic_fns=[torch.sin, torch.cos])
    print("posenc bit-exact:", bool((orc.posenc(x, 10) == eo.embed(x)).all()))
    K2, c2w2 = orc.lego_like_camera()
    ro, rd = ref.get_rays(800, 800, K2, c2w2)
    ro2, rd2 = orc.get_rays(800, 800, K2, c2w2)
    print("get_rays 800x800 bit-exact:", bool((rd == rd2).all() and (ro == ro2).all()))


if __name__ == "__main__":
    main()
