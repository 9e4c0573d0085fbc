// raw2outputs: density -> alpha, exclusive-cumprod transmittance, weighted
// rgb/depth/acc.  One warp per ray, shuffle scans, 128-bit loads of raw.
//
// Spec: upstream nerf-pytorch raw2outputs as restated in SURVEY.md 8c-S1 (the
// reference tree only carries the pointer comment, src/run_nerf_helpers.py:131-133).
// HBM-bound: fwd moves 24*S+36 B/ray, bwd 40*S+36 B/ray.
#include "ctx_common.cuh"

namespace ctx {

constexpr int kCompWarps = 8;  // warps per CTA

struct SampleTerms {
  float alpha, trans_factor, expo;  // expo = exp(-relu(sigma)*dist) ; trans_factor = 1-alpha+1e-10
};

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

__device__ __forceinline__ SampleTerms sample_terms(float sigma, float dist) {
  SampleTerms t;
  t.expo = expf(-fmaxf(sigma, 0.0f) * dist);
  t.alpha = 1.0f - t.expo;
  t.trans_factor = (1.0f - t.alpha) + 1e-10f;
  return t;
}

// inclusive product scan across the warp
__device__ __forceinline__ float warp_scan_prod(float v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const float n = __shfl_up_sync(CTX_FULL_MASK, v, o);
    if (lane >= o) v *= n;
  }
  return v;
}

// CH = number of 32-sample chunks held in registers (S <= 32*CH)
template <int CH>
__global__ void __launch_bounds__(kCompWarps * 32)
composite_fwd_kernel(const float4* __restrict__ raw, const float* __restrict__ z,
                     const float* __restrict__ rays_d, const float* __restrict__ noise,
                     int64_t R, int S, int white_bkgd,
                     float* __restrict__ rgb_map, float* __restrict__ disp_map,
                     float* __restrict__ acc_map, float* __restrict__ weights,
                     float* __restrict__ depth_map) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (int64_t)blockIdx.x * kCompWarps + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * kCompWarps;
  for (int64_t ray = warp0; ray < R; ray += nwarps) {
    const float dx = rays_d[ray * 3 + 0], dy = rays_d[ray * 3 + 1], dz = rays_d[ray * 3 + 2];
    const float dnorm = sqrtf(dx * dx + dy * dy + dz * dz);
    const int64_t base = ray * S;
    float4 rw[CH];
    float zc[CH], zn[CH], nz[CH];
    // issue every load of the ray up front
#pragma unroll
    for (int c = 0; c < CH; ++c) {
      const int s = c * 32 + lane;
      const bool ok = s < S;
      rw[c] = ok ? __ldg(raw + base + s) : make_float4(0.f, 0.f, 0.f, 0.f);
      zc[c] = ok ? __ldg(z + base + s) : 0.f;
      zn[c] = (s + 1 < S) ? __ldg(z + base + s + 1) : 0.f;
      nz[c] = (noise != nullptr && ok) ? __ldg(noise + base + s) : 0.f;
    }
    float carry = 1.0f, sr = 0.f, sg = 0.f, sb = 0.f, sd = 0.f, sa = 0.f;
#pragma unroll
    for (int c = 0; c < CH; ++c) {
      const int s = c * 32 + lane;
      const bool ok = s < S;
      if (c * 32 >= S) break;
      const float dist = ((s == S - 1) ? 1e10f : (zn[c] - zc[c])) * dnorm;
      SampleTerms t = sample_terms(rw[c].w + nz[c], dist);
      const float f = ok ? t.trans_factor : 1.0f;
      const float incl = warp_scan_prod(f, lane);
      float excl = __shfl_up_sync(CTX_FULL_MASK, incl, 1);
      if (lane == 0) excl = 1.0f;
      const float T = carry * excl;
      const float w = ok ? t.alpha * T : 0.f;
      carry *= __shfl_sync(CTX_FULL_MASK, incl, 31);
      if (ok) {
        weights[base + s] = w;
        sr += w * sigmoidf_(rw[c].x);
        sg += w * sigmoidf_(rw[c].y);
        sb += w * sigmoidf_(rw[c].z);
        sd += w * zc[c];
        sa += w;
      }
    }
    sr = warp_sum(sr); sg = warp_sum(sg); sb = warp_sum(sb); sd = warp_sum(sd); sa = warp_sum(sa);
    if (lane == 0) {
      const float bg = white_bkgd ? (1.0f - sa) : 0.0f;
      rgb_map[ray * 3 + 0] = sr + bg;
      rgb_map[ray * 3 + 1] = sg + bg;
      rgb_map[ray * 3 + 2] = sb + bg;
      depth_map[ray] = sd;
      acc_map[ray] = sa;
      // 1/max(1e-10, depth/acc); NaN when acc == 0 (0/0 propagates through max in torch)
      const float q = sd / sa;
      disp_map[ray] = (q != q) ? q : 1.0f / fmaxf(1e-10f, q);
    }
  }
}

// Backward.  With t_k = 1-alpha_k+1e-10, T_i = prod_{k<i} t_k, w_i = alpha_i T_i and
// G_i = dL/dw_i:   dL/dalpha_i = T_i * (G_i - S_i),
//   S_i = sum_{j>i} G_j alpha_j prod_{i<k<j} t_k      (suffix affine scan; no division
// by t_i, which is 1e-10 at an opaque sample -- SURVEY.md H5).
template <int CH>
__global__ void __launch_bounds__(kCompWarps * 32)
composite_bwd_kernel(const float4* __restrict__ raw, const float* __restrict__ z,
                     const float* __restrict__ rays_d, const float* __restrict__ noise,
                     int64_t R, int S, int white_bkgd,
                     const float* __restrict__ g_rgb, const float* __restrict__ g_disp,
                     const float* __restrict__ g_acc, const float* __restrict__ g_weights,
                     const float* __restrict__ g_depth, float4* __restrict__ g_raw) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (int64_t)blockIdx.x * kCompWarps + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * kCompWarps;
  for (int64_t ray = warp0; ray < R; ray += nwarps) {
    const float dx = rays_d[ray * 3 + 0], dy = rays_d[ray * 3 + 1], dz = rays_d[ray * 3 + 2];
    const float dnorm = sqrtf(dx * dx + dy * dy + dz * dz);
    const int64_t base = ray * S;
    float4 rw[CH];
    float zc[CH], gw[CH], alpha[CH], expo[CH], tf[CH], T[CH], dist[CH], sig[CH];
#pragma unroll
    for (int c = 0; c < CH; ++c) {
      const int s = c * 32 + lane;
      const bool ok = s < S;
      rw[c] = ok ? __ldg(raw + base + s) : make_float4(0.f, 0.f, 0.f, 0.f);
      zc[c] = ok ? __ldg(z + base + s) : 0.f;
      const float zn = (s + 1 < S) ? __ldg(z + base + s + 1) : 0.f;
      const float nz = (noise != nullptr && ok) ? __ldg(noise + base + s) : 0.f;
      gw[c] = (g_weights != nullptr && ok) ? __ldg(g_weights + base + s) : 0.f;
      dist[c] = ((s == S - 1) ? 1e10f : (zn - zc[c])) * dnorm;
      sig[c] = rw[c].w + nz;
    }
    // forward recompute: T_i, alpha_i, and the ray totals needed by the disp gradient
    float carry = 1.0f, sd = 0.f, sa = 0.f;
#pragma unroll
    for (int c = 0; c < CH; ++c) {
      const int s = c * 32 + lane;
      const bool ok = s < S;
      SampleTerms t = sample_terms(sig[c], dist[c]);
      alpha[c] = ok ? t.alpha : 0.f;
      expo[c] = t.expo;
      tf[c] = ok ? t.trans_factor : 1.0f;
      const float incl = warp_scan_prod(tf[c], lane);
      float excl = __shfl_up_sync(CTX_FULL_MASK, incl, 1);
      if (lane == 0) excl = 1.0f;
      T[c] = carry * excl;
      carry *= __shfl_sync(CTX_FULL_MASK, incl, 31);
      const float w = alpha[c] * T[c];
      sd += w * zc[c];
      sa += w;
    }
    sd = warp_sum(sd); sa = warp_sum(sa);
    const float gr = g_rgb ? g_rgb[ray * 3 + 0] : 0.f;
    const float gg = g_rgb ? g_rgb[ray * 3 + 1] : 0.f;
    const float gb = g_rgb ? g_rgb[ray * 3 + 2] : 0.f;
    float gd = g_depth ? g_depth[ray] : 0.f;
    float ga = g_acc ? g_acc[ray] : 0.f;
    const float gdisp = g_disp ? g_disp[ray] : 0.f;
    if (gdisp != 0.f) {  // disp = 1/max(1e-10, q), q = depth/acc
      const float q = sd / sa;
      if (q > 1e-10f) {
        const float dq = -gdisp / (q * q);
        gd += dq / sa;
        ga += -dq * sd / (sa * sa);
      } else if (q != q) {
        gd = q; ga = q;
      }
    }
    if (white_bkgd) ga -= (gr + gg + gb);
    // reverse affine scan over chunks
    float carry_u = 0.f;  // U of the first sample of the following chunk
#pragma unroll
    for (int c = CH - 1; c >= 0; --c) {
      const int s = c * 32 + lane;
      const bool ok = s < S;
      if (c * 32 >= S) continue;
      const float cr = sigmoidf_(rw[c].x), cg = sigmoidf_(rw[c].y), cb = sigmoidf_(rw[c].z);
      const float G = gw[c] + gr * cr + gg * cg + gb * cb + gd * zc[c] + ga;
      float a = tf[c];                      // padded lanes: a = 1, b = 0 (identity map)
      float b = ok ? G * alpha[c] : 0.f;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const float a2 = __shfl_down_sync(CTX_FULL_MASK, a, o);
        const float b2 = __shfl_down_sync(CTX_FULL_MASK, b, o);
        if (lane + o < 32) { b = b + a * b2; a = a * a2; }
      }
      const float U = b + a * carry_u;
      float Snext = __shfl_down_sync(CTX_FULL_MASK, U, 1);
      if (lane == 31) Snext = carry_u;
      carry_u = __shfl_sync(CTX_FULL_MASK, U, 0);
      if (ok) {
        const float g_alpha = T[c] * (G - Snext);
        const float g_sigma = (sig[c] > 0.f) ? g_alpha * expo[c] * dist[c] : 0.f;
        const float w = alpha[c] * T[c];
        float4 o4;
        o4.x = w * gr * cr * (1.0f - cr);
        o4.y = w * gg * cg * (1.0f - cg);
        o4.z = w * gb * cb * (1.0f - cb);
        o4.w = g_sigma;
        g_raw[base + s] = o4;
      }
    }
  }
}

static inline int comp_grid(int64_t R) {
  int64_t blocks = ceil_div(R, kCompWarps);
  const int64_t cap = (int64_t)kNumSMs * 8;  // 8 CTAs of 8 warps = 64 warps/SM, grid-stride beyond
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

}  // namespace ctx

#define CTX_COMP_DISPATCH(KERNEL, ...)                                                   \
  if (S <= 32) KERNEL<1><<<grid, block, 0, st>>>(__VA_ARGS__);                          \
  else if (S <= 64) KERNEL<2><<<grid, block, 0, st>>>(__VA_ARGS__);                     \
  else if (S <= 128) KERNEL<4><<<grid, block, 0, st>>>(__VA_ARGS__);                    \
  else if (S <= 192) KERNEL<6><<<grid, block, 0, st>>>(__VA_ARGS__);                    \
  else if (S <= 256) KERNEL<8><<<grid, block, 0, st>>>(__VA_ARGS__);                    \
  else if (S <= 384) KERNEL<12><<<grid, block, 0, st>>>(__VA_ARGS__);                   \
  else KERNEL<16><<<grid, block, 0, st>>>(__VA_ARGS__);

extern "C" int ctx_composite_fwd(const float* raw, const float* z_vals, const float* rays_d,
                                 const float* noise, int64_t R, int S, int white_bkgd,
                                 float* rgb_map, float* disp_map, float* acc_map, float* weights,
                                 float* depth_map, void* stream) {
  if (R < 0 || S < 1 || S > 512) return CTX_ERR_BAD_ARG;
  if (R == 0) return 0;
  if (!raw || !z_vals || !rays_d || !rgb_map || !disp_map || !acc_map || !weights || !depth_map)
    return CTX_ERR_BAD_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = ctx::comp_grid(R), block = ctx::kCompWarps * 32;
  CTX_COMP_DISPATCH(ctx::composite_fwd_kernel, (const float4*)raw, z_vals, rays_d, noise, R, S,
                    white_bkgd, rgb_map, disp_map, acc_map, weights, depth_map)
  CTX_RETURN_LAST();
}

extern "C" int ctx_composite_bwd(const float* raw, const float* z_vals, const float* rays_d,
                                 const float* noise, int64_t R, int S, int white_bkgd,
                                 const float* g_rgb, const float* g_disp, const float* g_acc,
                                 const float* g_weights, const float* g_depth, float* g_raw,
                                 void* stream) {
  if (R < 0 || S < 1 || S > 512) return CTX_ERR_BAD_ARG;
  if (R == 0) return 0;
  if (!raw || !z_vals || !rays_d || !g_raw) return CTX_ERR_BAD_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = ctx::comp_grid(R), block = ctx::kCompWarps * 32;
  CTX_COMP_DISPATCH(ctx::composite_bwd_kernel, (const float4*)raw, z_vals, rays_d, noise, R, S,
                    white_bkgd, g_rgb, g_disp, g_acc, g_weights, g_depth, (float4*)g_raw)
  CTX_RETURN_LAST();
}
