/* ctxnerf.h -- C-ABI of libctxnerf.so: the B200 (sm_100a) NeRF ray-march path.
 *
 * The reference (zaiisao/ConTEXTure-NeRF) has no FFI: its boundary for this
 * path is the Python surface of src/run_nerf_helpers.py.  Each entry point
 * below is what a ctypes binding placed in that module calls instead of the
 * eager-PyTorch body it replaces (file:line given per function; INTEGRATION.md
 * shows the binding).
 *
 * Conventions
 *  - every pointer is a DEVICE pointer unless marked "host"; tensors are
 *    contiguous row-major fp32 unless a stride argument is given;
 *  - the library never allocates, frees or synchronises; work is enqueued on
 *    `stream` (a cudaStream_t passed as void*), launchers are re-entrant;
 *  - return 0 on success, a positive cudaError_t on a CUDA failure, a negative
 *    CTX_ERR_* on a bad argument.  Nothing throws across the ABI;
 *  - "nullable" arguments may be NULL.
 */
#ifndef CTXNERF_H
#define CTXNERF_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CTXNERF_ABI_VERSION 3
#define CTX_ERR_BAD_ARG (-1)
#define CTX_ERR_UNSUPPORTED (-2)
#define CTX_ERR_NO_NCCL (-3)

int ctx_abi_version(void);
/* static string for a code returned by any ctx_* call */
const char* ctx_error_string(int code);

/* ---- positional encoding: Embedder.embed, src/run_nerf_helpers.py:44-45 ----
 * x [n,d] -> out [n, d*(include_input + 2L)], channel order of :24-39.        */
int ctx_posenc_fwd(const float* x, float* out, int64_t n, int d, int L, int include_input,
                   int log_sampling, void* stream);
int ctx_posenc_bwd(const float* x, const float* g_out, float* g_x, int64_t n, int d, int L,
                   int include_input, int log_sampling, void* stream);

/* ---- ray generation fused with stratified sampling --------------------------
 * get_rays, src/run_nerf_helpers.py:139-148 (+ optional ndc_rays :161-178 and
 * upstream render_rays' depth sampling).  c2w: device [3,4] row-major with
 * leading dimension c2w_ld.  ray_idx (nullable): flat pixel ids y*W+x to
 * generate (training batches); NULL = all H*W pixels in row-major order.
 * n_samples == 0 skips z_vals.  perturb != 0 jitters with `jitter`
 * [n_rays,n_samples] if given, else with Philox(seed).  sphere (host,
 * nullable: cx,cy,cz,r) replaces near/far by the ray/sphere interval.
 * viewdirs, z_vals, near_far ([n,2]) are nullable outputs.                     */
int ctx_raygen_fwd(int H, int W, float fx, float fy, float cx, float cy, const float* c2w,
                   int c2w_ld, const int64_t* ray_idx, int64_t n_rays, int use_ndc,
                   float ndc_focal, float ndc_near, int n_samples, float near, float far,
                   int lindisp, int perturb, const float* jitter, uint64_t seed,
                   const uint64_t* seed_dev, int use_sphere, const float* sphere, float* rays_o,
                   float* rays_d, float* viewdirs, float* z_vals, float* near_far, void* stream);
/* seed_dev (nullable, here and in ctx_resample_fwd): a DEVICE counter added to `seed` inside the kernel, so that a
 * captured CUDA graph of the training step draws fresh Philox numbers on every replay (ctx_adam_step_dev bumps it). */

/* z_vals [R,S] from per-ray near/far (ray_batch[:,6], ray_batch[:,7] upstream) */
int ctx_stratified_fwd(const float* near, int64_t near_stride, const float* far,
                       int64_t far_stride, int64_t R, int S, int lindisp, int perturb,
                       const float* jitter, uint64_t seed, float* z_vals, void* stream);

/* ---- ndc_rays, src/run_nerf_helpers.py:161-178 ------------------------------ */
int ctx_ndc_fwd(int H, int W, float focal, float near, const float* rays_o, const float* rays_d,
                int64_t n, float* o_out, float* d_out, void* stream);
int ctx_ndc_bwd(int H, int W, float focal, float near, const float* rays_o, const float* rays_d,
                const float* g_o, const float* g_d, int64_t n, float* g_rays_o, float* g_rays_d,
                void* stream);

/* ---- raw2outputs (absent from the reference; pointer comment at
 * src/run_nerf_helpers.py:131-133; spec SURVEY.md 8c-S1) -----------------------
 * raw [R,S,4], z_vals [R,S], rays_d [R,3], noise (nullable, [R,S], already
 * scaled by raw_noise_std) -> rgb [R,3], disp [R], acc [R], weights [R,S],
 * depth [R].  Any S up to 16384 (rays longer than 512 samples run a chunked variant
 * of the same scan).                                                             */
int ctx_composite_fwd(const float* raw, const float* z_vals, const float* rays_d,
                      const float* noise, int64_t R, int S, int white_bkgd, float* rgb_map,
                      float* disp_map, float* acc_map, float* weights, float* depth_map,
                      void* stream);
/* g_* inputs nullable (treated as zero); writes g_raw [R,S,4] */
int ctx_composite_bwd(const float* raw, const float* z_vals, const float* rays_d,
                      const float* noise, int64_t R, int S, int white_bkgd, const float* g_rgb,
                      const float* g_disp, const float* g_acc, const float* g_weights,
                      const float* g_depth, float* g_raw, void* stream);

/* ---- sample_pdf, src/run_nerf_helpers.py:182-225 ----------------------------
 * bins [R,B] with row stride bins_stride (mid_bins != 0: `bins` is z [R,B+1]
 * and the mid-points .5*(z[i+1]+z[i]) are formed in-kernel), weights [R,B-1]
 * with row stride w_stride (the reference passes the view weights[...,1:-1]).
 * cdf_in (nullable [R,B]) injects stage 1.  u (nullable [R,N]) injects the
 * uniforms; else det -> linspace(0,1,N), otherwise Philox(seed).
 * Outputs: samples [R,N]; inds (nullable, int64 [R,N]) = searchsorted result;
 * z_all (nullable [R,Sm+N]) = sort(cat[z_merge[R,Sm], samples]).               */
int ctx_resample_fwd(const float* bins, int64_t bins_stride, int mid_bins, const float* weights,
                     int64_t w_stride, const float* cdf_in, const float* u, int det,
                     uint64_t seed, const uint64_t* seed_dev, int64_t R, int B, int N, float* samples,
                     int64_t* inds, const float* z_merge, int64_t zm_stride, int Sm, float* z_all,
                     void* stream);
/* d samples / d weights -> g_weights [R,B-1] (contiguous) */
int ctx_resample_bwd(const float* bins, int64_t bins_stride, int mid_bins, const float* weights,
                     int64_t w_stride, const float* u, int det, uint64_t seed, int64_t R, int B,
                     int N, const float* g_samples, float* g_weights, void* stream);

/* Fused training form of raw2outputs (NerfTrainer.step): forward + img2mse(rgb_map, target)
 * (src/run_nerf_helpers.py:9) + backward in ONE pass over raw.  g_raw [R,S,4] = d loss / d raw with
 * loss = sum((rgb_map - target)^2) * loss_scale (loss_scale = 1/(3R) for the image mean); loss[0] += that value
 * (zero it once per step; the coarse and the fine pass add into the same scalar).  weights [R,S] and rgb_map [R,3]
 * are optional outputs (the coarse pass feeds sample_pdf with its weights).                        */
int ctx_composite_train(const float* raw, const float* z_vals, const float* rays_d, const float* noise,
                        int64_t R, int S, int white_bkgd, const float* target, float loss_scale, float* loss,
                        float* g_raw, float* weights, float* rgb_map, void* stream);

/* ---- coordinate MLP: NeRF2D, src/run_nerf_helpers.py:68-135 (+ the upstream
 * view-direction head kept as comments :86-95, :117-127) ----------------------
 * tcgen05/TMEM kernel, bf16 operands, fp32 accumulation, hidden width 256.
 * `net` is a HOST blob of ctx_mlp_net_bytes() bytes filled by ctx_mlp_describe:
 * D pts layers, skip_mask bit i set <=> "i in skips" (input re-concatenated
 * after layer i, :114-115), in_pts/in_views real encoding widths (63|42 / 27|0),
 * out_ch 4|3 (must be 4 with views).                                            */
int ctx_mlp_net_bytes(void);
int ctx_mlp_describe(int D, uint32_t skip_mask, int in_pts, int in_views, int out_ch, void* net);
/* params: HOST array of n_params DEVICE pointers (fp32, nn.Linear [out,in]) in
 * module order: pts_linears.{0..D-1}.{weight,bias}, then output_linear.{w,b}
 * or feature_linear, alpha_linear, views_linears.0, rgb_linear {w,b}.
 * wpacked [w_bytes], wtpacked [wt_bytes] (nullable), fparams [n_fparams]: device
 * buffers sized from the net blob (see contexture-nerf_b200/csrc/mlp_desc.h).   */
int ctx_mlp_pack(const void* net, const float* const* params, int n_params, void* wpacked,
                 void* wtpacked, float* fparams, void* stream);
/* mode 2: UV-grid texture query, see ctx_tanh01_fwd below.
 * mode 0: x [P,x_ld] holds the already-encoded input [pts_enc | view_enc];
 * mode 1: the P = R*S points o + d*z (z [R,S]) and the view directions are
 * encoded in-kernel straight into shared memory (Embedder.embed :44-45, L_pts
 * / L_dirs frequencies).  out [P,out_ch] raw network output (no activation,
 * :129).  acts (nullable): activation records for ctx_mlp_bwd, sized for a
 * multiple of 4 tiles of 128 points (the kernel works on 512 points per SM pair).*/
int ctx_mlp_fwd(const void* net, const void* wpacked, const float* fparams, int mode, const float* x,
                int x_ld, const float* rays_o, const float* rays_d, const float* viewdirs,
                const float* z, int S, int L_pts, int L_dirs, int64_t P, float* out, void* acts,
                void* stream);

/* same with an SM budget (max_sms > 0: at most that many SMs, 0 = the whole GPU), see ctx_mlp_dgrad_ex, and with
 * mode 3: x holds RAW coordinates [*, x_ld] (x_ld = 2 or 3, nets without views), encoded in-kernel with L_pts
 * frequencies; gather (nullable, int64 [P]): point p reads row gather[p] of x -- the MLP evaluated only at the
 * texels a mesh uses (get_texture_map_only_valid_areas, src/models/textured_mesh.py:303-347).                  */
int ctx_mlp_fwd_ex(const void* net, const void* wpacked, const float* fparams, int mode, const float* x,
                   int x_ld, const float* rays_o, const float* rays_d, const float* viewdirs,
                   const float* z, int S, int L_pts, int L_dirs, int64_t P, float* out, void* acts,
                   const int64_t* gather, int max_sms, void* stream);

/* Backward of ctx_mlp_fwd w.r.t. the parameters (hand-written dgrad + wgrad
 * tcgen05 kernels; the encoded inputs are data and get no gradient).  g_out
 * [P,out_ch]; acts = records written by ctx_mlp_fwd; dacts = scratch of the same
 * size; grads = HOST array of DEVICE pointers in the order of ctx_mlp_pack's
 * params; gradients are ACCUMULATED (+=) into them.                              */
int ctx_mlp_bwd(const void* net, const void* wtpacked, const float* fparams, const float* g_out,
                const void* acts, void* dacts, int64_t P, float* const* grads, int n_grads,
                const float* const* params, float* scratch, void* stream);
/* the two halves of ctx_mlp_bwd, separately launchable (and separately timed by bench.py):
 * dgrad fills the dZ records from g_out, wgrad reduces records + dZ records into the gradients. */
int ctx_mlp_dgrad(const void* net, const void* wtpacked, const float* fparams, const float* g_out,
                  const void* acts, void* dacts, int64_t P, void* stream);
/* params (HOST array of DEVICE pointers, order of ctx_mlp_pack) and scratch (device,
 * ctx_mlp_wgrad_scratch_floats() floats, zeroed by the launcher) are read only for a view-direction net: the
 * feature layer is linear, so its weight gradients and those of views_linears.0[:, :256] are rebuilt from
 * G = [dZ_views | g_out]^T h and the weights instead of from records of the feature activations (which are never
 * written).  Both may be NULL for a net without views.                                                        */
int ctx_mlp_wgrad_scratch_floats(void);
int ctx_mlp_wgrad(const void* net, const void* acts, const void* dacts, int64_t P, float* const* grads,
                  int n_grads, const float* const* params, float* scratch, void* stream);
/* the same with an SM budget (max_sms > 0: at most that many SMs are occupied; 0 = the whole GPU), so that the
 * tensor-bound dgrad of one network can run beside the HBM-bound wgrad of the other on disjoint SM pairs
 * (NerfTrainer.step: dgrad_coarse || wgrad_fine).  dgrad accepts any budget >= 2 (one cluster); wgrad needs two
 * SMs per (layer, segment) job -- 28 for the view-direction net -- and returns CTX_ERR_UNSUPPORTED below that.  */
int ctx_mlp_dgrad_ex(const void* net, const void* wtpacked, const float* fparams, const float* g_out,
                     const void* acts, void* dacts, int64_t P, int max_sms, void* stream);
int ctx_mlp_wgrad_ex(const void* net, const void* acts, const void* dacts, int64_t P, float* const* grads,
                     int n_grads, const float* const* params, float* scratch, int max_sms, void* stream);

/* ---- fused texture map: get_texture_map, src/models/textured_mesh.py:266-301 -----------------------
 * ctx_mlp_fwd mode 2 generates the res x res UV grid (meshgrid of linspace(0,1,res), 'xy' indexing, :269-272),
 * encodes it (L_pts frequencies, 2-D) in shared memory and runs the 42->3 texture MLP: pass x = NULL, S = res,
 * P = res*res.  ctx_tanh01_fwd then applies (tanh+1)/2 (:299) and writes the NCHW layout [C, P] of
 * `.reshape(1,res,res,3).permute(0,3,1,2)`; ctx_tanh01_bwd is its backward (g_raw_in nullable: extra gradient
 * arriving directly on mlp_output).                                                                      */
int ctx_tanh01_fwd(const float* raw, float* out, int64_t P, int C, void* stream);
int ctx_tanh01_bwd(const float* raw, const float* g_tex, const float* g_raw_in, float* g_raw, int64_t P, int C,
                   void* stream);

/* ---- texture lookup of the rasterised mesh: kal.render.mesh.texture_mapping + mask / background composite,
 * src/models/render.py:133-140.  uv [B,N,2] in [0,1] (v up), tex [tex_batch,C,H,W] with
 * tex_batch == B or 1 (one atlas shared by every view: its gradient is the sum over the views), mode 0 = nearest,
 * 1 = bilinear (grid_sample, align_corners=False, padding border).  mask (nullable, [B,N]): out = sample*mask +
 * bg*(1-mask), bg (nullable) = C background values.  Only the texture receives a gradient (upstream detaches uv);
 * ctx_texmap_bwd ACCUMULATES into g_tex.                                                                      */
int ctx_texmap_fwd(const float* uv, const float* tex, const float* mask, const float* bg, float* out, int64_t B,
                   int64_t N, int tex_batch, int C, int H, int W, int mode, void* stream);
int ctx_texmap_bwd(const float* uv, const float* mask, const float* g_out, float* g_tex, int64_t B, int64_t N,
                   int tex_batch, int C, int H, int W, int mode, void* stream);

/* dst [n,C] = 0 ; dst[idx[m]] = src[m] * scale (m < M)  -- `final_image[mask] = colors` of
 * get_texture_map_only_valid_areas (:345) -- and its backward g_src[m] = g_dst[idx[m]] * scale.                  */
int ctx_rows_scatter(const float* src, const int64_t* idx, float scale, float* dst, int64_t M, int64_t n, int C,
                     void* stream);
int ctx_rows_gather(const float* g_dst, const int64_t* idx, float scale, float* g_src, int64_t M, int C,
                    void* stream);

/* ---- training-step glue ------------------------------------------------------
 * img2mse(a,t) + img2mse(b,t) (src/run_nerf_helpers.py:9) and its gradient in one
 * pass: loss[0] = mean((a-t)^2) [+ mean((b-t)^2)], g_a = 2(a-t)/n*scale (b, g_*
 * nullable).  n = number of elements.                                            */
int ctx_mse_fwd_bwd(const float* a, const float* b, const float* target, int64_t n, float scale,
                    float* loss, float* g_a, float* g_b, void* stream);
/* torch.optim.Adam update (src/training/trainer.py:603) over flat fp32 buckets;
 * step >= 1 is the bias-correction step, grad_scale multiplies the gradient
 * (1/world_size after a sum all-reduce).                                         */
int ctx_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n,
                  float lr, float beta1, float beta2, float eps, int step, float weight_decay,
                  float grad_scale, void* stream);

/* Device-resident step state for a CUDA-graph-captured training step: counters = two
 * uint64 on the device, [0] Philox seed offset (see seed_dev above), [1] Adam step count.  ctx_step_tick runs first
 * in a step: counters[0] += 2, counters[1] += 1, loss[0] = 0 (loss nullable).  ctx_adam_step_dev is ctx_adam_step
 * with the step count read from counters[1].                                                                  */
int ctx_step_tick(void* counters, float* loss, void* stream);
int ctx_adam_step_dev(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n,
                      float lr, float beta1, float beta2, float eps, const void* counters, float weight_decay,
                      float grad_scale, void* stream);

/* ---- view-weight masks: ConTEXTure.create_face_view_map / compare_face_normals_between_views,
 * src/training/trainer.py:155-249 (the reference uses torch_scatter.scatter_max, :227).  face_idx [V,H,W] int64
 * (< 0 = background), face_normals [V,3,F] fp32.  mask [V,H,W] bytes: 1 unless the pixel's face has a larger
 * z-normal in another view that shows it.  visible [V*F] bytes / maxz [F] floats: caller-owned scratch.            */
int ctx_view_weight_masks(const float* face_normals, const int64_t* face_idx, int V, int F, int H, int W,
                          unsigned char* visible, float* maxz, unsigned char* mask, void* stream);
/* rows (face, view, i, j) of the covered pixels in (view, pixel) order: pass 0 fills block_counts
 * [ctx_face_view_map_blocks(V*H*W) + 1] (exclusive offsets; the last entry = number of rows N), pass 1 writes
 * rows [N,4] int64.                                                                                              */
int64_t ctx_face_view_map_blocks(int64_t n_pixels);
int ctx_face_view_map(const int64_t* face_idx, int V, int H, int W, int64_t* block_counts, int64_t* rows, int pass,
                      void* stream);

/* ---- render_rays, the coarse + fine driver (SURVEY.md 8b "fused driver", 8c S2; upstream run_nerf.py, to which
 * src/run_nerf_helpers.py:131-133 points): rays + stratified depths -> network on n_samples points -> raw2outputs ->
 * sample_pdf(z_mid, weights[...,1:-1], n_importance, det = !perturb) -> sort(cat[z, z_samples]) -> network_fine on
 * n_samples + n_importance points -> raw2outputs, as six launches on `stream`.  Inference form (no records, no
 * density noise).  Every buffer is the caller's; R = n_rays, S = n_samples, Ni = n_importance.                    */
typedef struct CtxNet {
  const void* desc;      /* host blob of ctx_mlp_describe */
  const void* wpacked;   /* device, ctx_mlp_pack */
  const float* fparams;  /* device, ctx_mlp_pack */
} CtxNet;
typedef struct CtxRenderArgs {
  int H, W;                          /* get_rays(H, W, K, c2w): K = [[fx,0,cx],[0,fy,cy],[0,0,1]] */
  float fx, fy, cx, cy;
  const float* c2w;                  /* device [3,4], leading dimension c2w_ld */
  int c2w_ld;
  const int64_t* ray_idx;            /* nullable: flat pixel ids y*W+x; NULL = pixels 0..n_rays-1 */
  int64_t n_rays;
  float near, far;
  int lindisp, perturb;              /* perturb != 0: jittered depths and random u (Philox(seed [+ *seed_dev])) */
  uint64_t seed;
  const uint64_t* seed_dev;          /* nullable device counter added to seed */
  const float* sphere;               /* HOST, nullable: (cx,cy,cz,r) -> per-ray near/far from the ray/sphere interval */
  int n_samples, n_importance;       /* n_importance == 0: single pass, results in the final maps */
  int white_bkgd;
  int L_pts, L_dirs;                 /* encoding frequencies (10 / 4); L_dirs == 0 for a net without view directions */
  int max_sms;                       /* SM budget of the MLP launches (0 = whole GPU) */
  CtxNet coarse, fine;               /* fine.desc == NULL: the coarse network serves both passes */
  /* workspace (device) */
  float *rays_o, *rays_d, *viewdirs; /* [R,3] each; viewdirs nullable when L_dirs == 0 */
  float *z_coarse, *raw_coarse, *weights_coarse;            /* [R,S], [R,S,4], [R,S] */
  float *z_samples, *z_fine, *raw_fine, *weights_fine;      /* [R,Ni], [R,S+Ni], [R,S+Ni,4], [R,S+Ni] */
  /* outputs (device): coarse maps (used when n_importance > 0) and final maps */
  float *rgb0, *disp0, *acc0, *depth0;                      /* [R,3], [R], [R], [R] */
  float *rgb_map, *disp_map, *acc_map, *depth_map;
} CtxRenderArgs;
int ctx_render_rays(const CtxRenderArgs* args, void* stream);
int ctx_render_args_bytes(void);   /* sizeof(CtxRenderArgs) in the library (layout check for bindings) */

/* ---- gradient all-reduce: the one exchange step of the ray-sharded training step, replacing the
 * reduce-add of nn.DataParallel (src/training/trainer.py:134-135).  NCCL is opened at run time (dlopen; `path`
 * nullable = "libnccl.so.2" by soname), so the library has no link-time dependency on it; every call below returns
 * CTX_ERR_NO_NCCL (text in ctx_comm_last_error) when NCCL is absent or reports an error.  The communicator is an
 * opaque handle owned by the caller.  Rendezvous: one rank makes the 128-byte token with ctx_comm_unique_id and the
 * host passes it to the others; ctx_comm_init is collective and binds to the CURRENT device.  ctx_allreduce sums
 * bucket[0..n) in place on `stream`; it can be captured into a CUDA graph (all ranks capture the same sequence).  */
#define CTX_COMM_ID_BYTES 128
int ctx_comm_load(const char* path);
int ctx_comm_version(void);
const char* ctx_comm_last_error(void);
int ctx_comm_unique_id(void* id_out_host);
int ctx_comm_init(void** comm_out, int n_ranks, const void* id_host, int rank);
int ctx_comm_destroy(void* comm);
int ctx_allreduce(void* comm, float* bucket, int64_t n, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CTXNERF_H */
