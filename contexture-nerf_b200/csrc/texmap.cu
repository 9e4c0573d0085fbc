// texture_mapping + mask / background composite of the rasterised mesh.
//
// Reference call site: /root/reference/src/models/render.py:133-140 --
//   image_features = kal.render.mesh.texture_mapping(uv_features, texture_map, mode)   (uv_features is detached, :121)
//   image_features = image_features * mask ; image_features += 1 * (1 - mask)          (white background)
// kaolin's texture_mapping (third party, un-vendored; restated in oracle/nerf_oracle.py) maps uv in [0,1] to
// grid_sample coordinates (u*2-1, -(v*2-1)) and samples with align_corners=False, padding_mode='border'.
// One thread per pixel: coordinates once, then per channel four gathers (bilinear) or one (nearest); the backward
// scatters g_out * mask into the texture gradient with fp32 atomics (only the texture gets a gradient, as upstream).
// HBM-bound: 8 B uv + 4 B mask read, 4C B written per pixel; the texture (12.6 MB at 1024^2 x 3) lives in L2.
#include "ctx_common.cuh"
#include <initializer_list>

namespace ctx {

struct TexCoord {
  int x0, y0, x1, y1;
  float w00, w01, w10, w11;   // (y, x) corner weights
};

// grid_sample's unnormalise (align_corners=False) + border clip, in torch's operation order
__device__ __forceinline__ float tex_unnormalize(float coord, int size) {
  float x = __fmul_rn(__fsub_rn(__fmul_rn(__fadd_rn(coord, 1.0f), (float)size), 1.0f), 0.5f);
  return fminf((float)(size - 1), fmaxf(x, 0.0f));
}

__device__ __forceinline__ TexCoord tex_coord(float u, float v, int H, int W, int bilinear) {
  const float gx = __fsub_rn(__fmul_rn(u, 2.0f), 1.0f);
  const float gy = -__fsub_rn(__fmul_rn(v, 2.0f), 1.0f);
  const float ix = tex_unnormalize(gx, W), iy = tex_unnormalize(gy, H);
  TexCoord t;
  if (bilinear) {
    const float fx = floorf(ix), fy = floorf(iy);
    t.x0 = (int)fx; t.y0 = (int)fy; t.x1 = t.x0 + 1; t.y1 = t.y0 + 1;
    const float ax = ix - fx, ay = iy - fy;        // distance to the north-west texel
    t.w00 = (1.0f - ax) * (1.0f - ay); t.w01 = ax * (1.0f - ay);
    t.w10 = (1.0f - ax) * ay;          t.w11 = ax * ay;
    if (t.x1 > W - 1) { t.x1 = W - 1; t.w01 = 0.f; t.w11 = 0.f; }   // out-of-range corners contribute nothing
    if (t.y1 > H - 1) { t.y1 = H - 1; t.w10 = 0.f; t.w11 = 0.f; }
  } else {
    t.x0 = t.x1 = (int)rintf(ix); t.y0 = t.y1 = (int)rintf(iy);     // nearbyint, as torch
    t.w00 = 1.0f; t.w01 = t.w10 = t.w11 = 0.f;
  }
  return t;
}

__global__ void __launch_bounds__(256)
texmap_fwd_kernel(const float2* __restrict__ uv, const float* __restrict__ tex, const float* __restrict__ mask,
                  const float* __restrict__ bg, float* __restrict__ out, int64_t B, int64_t N, int tex_batch, int C,
                  int H, int W, int bilinear) {
  const int64_t total = B * N;
  const int64_t plane = (int64_t)H * W;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = i / N;
    const float2 c = __ldg(uv + i);
    const TexCoord t = tex_coord(c.x, c.y, H, W, bilinear);
    const float m = mask ? __ldg(mask + i) : 1.0f;
    const float* tb = tex + (tex_batch > 1 ? b : 0) * C * plane;
    const int64_t o00 = (int64_t)t.y0 * W + t.x0, o01 = (int64_t)t.y0 * W + t.x1;
    const int64_t o10 = (int64_t)t.y1 * W + t.x0, o11 = (int64_t)t.y1 * W + t.x1;
    for (int ch = 0; ch < C; ++ch) {
      const float* p = tb + ch * plane;
      float v = t.w00 * __ldg(p + o00);
      if (bilinear) v += t.w01 * __ldg(p + o01) + t.w10 * __ldg(p + o10) + t.w11 * __ldg(p + o11);
      if (mask) v = v * m + (bg ? __ldg(bg + ch) : 0.f) * (1.0f - m);
      out[i * C + ch] = v;
    }
  }
}

__global__ void __launch_bounds__(256)
texmap_bwd_kernel(const float2* __restrict__ uv, const float* __restrict__ mask, const float* __restrict__ g_out,
                  float* __restrict__ g_tex, int64_t B, int64_t N, int tex_batch, int C, int H, int W, int bilinear) {
  const int64_t total = B * N;
  const int64_t plane = (int64_t)H * W;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const float m = mask ? __ldg(mask + i) : 1.0f;
    if (m == 0.f) continue;                          // background pixels carry no texture gradient
    const int64_t b = i / N;
    const float2 c = __ldg(uv + i);
    const TexCoord t = tex_coord(c.x, c.y, H, W, bilinear);
    float* tb = g_tex + (tex_batch > 1 ? b : 0) * C * plane;
    const int64_t o00 = (int64_t)t.y0 * W + t.x0, o01 = (int64_t)t.y0 * W + t.x1;
    const int64_t o10 = (int64_t)t.y1 * W + t.x0, o11 = (int64_t)t.y1 * W + t.x1;
    for (int ch = 0; ch < C; ++ch) {
      const float g = __ldg(g_out + i * C + ch) * m;
      float* p = tb + ch * plane;
      atomicAdd(p + o00, t.w00 * g);
      if (bilinear) {
        if (t.w01 != 0.f) atomicAdd(p + o01, t.w01 * g);
        if (t.w10 != 0.f) atomicAdd(p + o10, t.w10 * g);
        if (t.w11 != 0.f) atomicAdd(p + o11, t.w11 * g);
      }
    }
  }
}

// C == 3, N % 4 == 0, 16-byte aligned rows: four consecutive pixels per thread -- uv as two float4, mask as one,
// the twelve outputs as three float4 stores (48 contiguous bytes), 48 independent gathers in flight per thread;
// blockIdx.y is the view, so no 64-bit division.
template <bool kBwd>
__global__ void __launch_bounds__(256)
texmap_rgb4_kernel(const float4* __restrict__ uv, const float* __restrict__ tex, const float4* __restrict__ mask,
                   const float* __restrict__ bg, float4* __restrict__ io, float* __restrict__ g_tex, int64_t N,
                   int tex_batch, int H, int W, int bilinear) {
  const int64_t b = blockIdx.y;
  const int64_t plane = (int64_t)H * W;
  const int64_t nq = N >> 2;
  const int64_t tex_off = (tex_batch > 1 ? b : 0) * 3 * plane;
  for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < nq; q += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i4 = b * nq + q;                   // index of this group of four pixels
    const float4 c01 = __ldg(uv + 2 * i4), c23 = __ldg(uv + 2 * i4 + 1);
    const float4 m4 = mask ? __ldg(mask + i4) : make_float4(1.f, 1.f, 1.f, 1.f);
    const float us[4] = {c01.x, c01.z, c23.x, c23.z}, vs[4] = {c01.y, c01.w, c23.y, c23.w};
    const float ms[4] = {m4.x, m4.y, m4.z, m4.w};
    float v[12];
    if (kBwd) {
      const float4 a = __ldg(io + 3 * i4), bq = __ldg(io + 3 * i4 + 1), cq = __ldg(io + 3 * i4 + 2);
      v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = bq.x; v[5] = bq.y; v[6] = bq.z; v[7] = bq.w;
      v[8] = cq.x; v[9] = cq.y; v[10] = cq.z; v[11] = cq.w;
    }
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      const TexCoord t = tex_coord(us[p], vs[p], H, W, bilinear);
      const int64_t o00 = (int64_t)t.y0 * W + t.x0, o01 = (int64_t)t.y0 * W + t.x1;
      const int64_t o10 = (int64_t)t.y1 * W + t.x0, o11 = (int64_t)t.y1 * W + t.x1;
#pragma unroll
      for (int ch = 0; ch < 3; ++ch) {
        if (!kBwd) {
          const float* pl = tex + tex_off + ch * plane;
          float r = t.w00 * __ldg(pl + o00);
          if (bilinear) r += t.w01 * __ldg(pl + o01) + t.w10 * __ldg(pl + o10) + t.w11 * __ldg(pl + o11);
          if (mask) r = r * ms[p] + (bg ? __ldg(bg + ch) : 0.f) * (1.0f - ms[p]);
          v[p * 3 + ch] = r;
        } else if (ms[p] != 0.f) {
          float* pl = g_tex + tex_off + ch * plane;
          const float g = v[p * 3 + ch] * ms[p];
          atomicAdd(pl + o00, t.w00 * g);
          if (bilinear) {
            if (t.w01 != 0.f) atomicAdd(pl + o01, t.w01 * g);
            if (t.w10 != 0.f) atomicAdd(pl + o10, t.w10 * g);
            if (t.w11 != 0.f) atomicAdd(pl + o11, t.w11 * g);
          }
        }
      }
    }
    if (!kBwd) {
      io[3 * i4] = make_float4(v[0], v[1], v[2], v[3]);
      io[3 * i4 + 1] = make_float4(v[4], v[5], v[6], v[7]);
      io[3 * i4 + 2] = make_float4(v[8], v[9], v[10], v[11]);
    }
  }
}

// ---- bicubic (the third interpolation mode render.py:9 allows): grid_sample's cubic convolution, A = -0.75,
// align_corners=False, 4 x 4 taps whose INDICES are clipped to the border (padding_mode='border' as kaolin passes).
struct CubicTaps {
  int ix[4], iy[4];
  float cx[4], cy[4];
};
__device__ __forceinline__ void cubic_coeffs(float t, float (&c)[4]) {
  const float A = -0.75f;
  const float x0 = t + 1.0f, x3 = 2.0f - t, x2 = 1.0f - t;
  c[0] = ((A * x0 - 5.0f * A) * x0 + 8.0f * A) * x0 - 4.0f * A;
  c[1] = ((A + 2.0f) * t - (A + 3.0f)) * t * t + 1.0f;
  c[2] = ((A + 2.0f) * x2 - (A + 3.0f)) * x2 * x2 + 1.0f;
  c[3] = ((A * x3 - 5.0f * A) * x3 + 8.0f * A) * x3 - 4.0f * A;
}
__device__ __forceinline__ CubicTaps cubic_taps(float u, float v, int H, int W) {
  const float gx = __fsub_rn(__fmul_rn(u, 2.0f), 1.0f);
  const float gy = -__fsub_rn(__fmul_rn(v, 2.0f), 1.0f);
  // unnormalise WITHOUT clipping (torch clips the tap indices, not the sampling position, in bicubic mode)
  const float ix = __fmul_rn(__fsub_rn(__fmul_rn(__fadd_rn(gx, 1.0f), (float)W), 1.0f), 0.5f);
  const float iy = __fmul_rn(__fsub_rn(__fmul_rn(__fadd_rn(gy, 1.0f), (float)H), 1.0f), 0.5f);
  const float fx = floorf(ix), fy = floorf(iy);
  CubicTaps t;
  cubic_coeffs(ix - fx, t.cx);
  cubic_coeffs(iy - fy, t.cy);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    t.ix[k] = (int)fminf((float)(W - 1), fmaxf(fx - 1.0f + (float)k, 0.0f));
    t.iy[k] = (int)fminf((float)(H - 1), fmaxf(fy - 1.0f + (float)k, 0.0f));
  }
  return t;
}

template <bool kBwd>
__global__ void __launch_bounds__(256)
texmap_bicubic_kernel(const float2* __restrict__ uv, const float* __restrict__ tex, const float* __restrict__ mask,
                      const float* __restrict__ bg, float* __restrict__ io, float* __restrict__ g_tex, int64_t B,
                      int64_t N, int tex_batch, int C, int H, int W) {
  const int64_t total = B * N, plane = (int64_t)H * W;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const float m = mask ? __ldg(mask + i) : 1.0f;
    if (kBwd && m == 0.f) continue;
    const int64_t b = i / N;
    const float2 c = __ldg(uv + i);
    const CubicTaps t = cubic_taps(c.x, c.y, H, W);
    const int64_t off = (tex_batch > 1 ? b : 0) * C * plane;
    for (int ch = 0; ch < C; ++ch) {
      if (!kBwd) {
        const float* p = tex + off + ch * plane;
        float acc = 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float row = 0.f;
#pragma unroll
          for (int k = 0; k < 4; ++k) row += t.cx[k] * __ldg(p + (int64_t)t.iy[j] * W + t.ix[k]);
          acc += t.cy[j] * row;
        }
        if (mask) acc = acc * m + (bg ? __ldg(bg + ch) : 0.f) * (1.0f - m);
        io[i * C + ch] = acc;
      } else {
        float* p = g_tex + off + ch * plane;
        const float g = __ldg(io + i * C + ch) * m;
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
          for (int k = 0; k < 4; ++k) atomicAdd(p + (int64_t)t.iy[j] * W + t.ix[k], t.cy[j] * t.cx[k] * g);
      }
    }
  }
}

static inline bool texmap_vec_ok(int64_t N, int C, std::initializer_list<const void*> ptrs) {
  if (C != 3 || (N & 3) != 0) return false;
  for (const void* p : ptrs)
    if (p && (reinterpret_cast<uintptr_t>(p) & 15) != 0) return false;
  return true;
}

static inline int texmap_grid(int64_t total) {
  int64_t blocks = ceil_div(total, 256);
  const int64_t cap = (int64_t)num_sms() * 8;
  if (blocks > cap) blocks = cap;
  return (int)(blocks < 1 ? 1 : blocks);
}

}  // namespace ctx

namespace ctx {
__global__ void rows_scatter_kernel(const float* __restrict__ src, const long long* __restrict__ idx, float scale,
                                    float* __restrict__ dst, int64_t M, int C) {
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < M * C; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t m = t / C;
    dst[idx[m] * C + (t - m * C)] = src[t] * scale;
  }
}
__global__ void rows_gather_kernel(const float* __restrict__ g_dst, const long long* __restrict__ idx, float scale,
                                   float* __restrict__ g_src, int64_t M, int C) {
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < M * C; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t m = t / C;
    g_src[t] = g_dst[idx[m] * C + (t - m * C)] * scale;
  }
}
}  // namespace ctx

extern "C" int ctx_rows_scatter(const float* src, const int64_t* idx, float scale, float* dst, int64_t M, int64_t n,
                                int C, void* stream) {
  if (M < 0 || n < 0 || C < 1 || !dst) return CTX_ERR_BAD_ARG;
  cudaError_t e = cudaMemsetAsync(dst, 0, (size_t)n * C * sizeof(float), (cudaStream_t)stream);
  if (e != cudaSuccess) return (int)e;
  if (M == 0) return 0;
  if (!src || !idx) return CTX_ERR_BAD_ARG;
  ctx::rows_scatter_kernel<<<ctx::texmap_grid(M * C), 256, 0, (cudaStream_t)stream>>>(src, (const long long*)idx, scale,
                                                                                     dst, M, C);
  CTX_RETURN_LAST();
}

extern "C" int ctx_rows_gather(const float* g_dst, const int64_t* idx, float scale, float* g_src, int64_t M, int C,
                               void* stream) {
  if (M < 0 || C < 1) return CTX_ERR_BAD_ARG;
  if (M == 0) return 0;
  if (!g_dst || !idx || !g_src) return CTX_ERR_BAD_ARG;
  ctx::rows_gather_kernel<<<ctx::texmap_grid(M * C), 256, 0, (cudaStream_t)stream>>>(g_dst, (const long long*)idx, scale,
                                                                                    g_src, M, C);
  CTX_RETURN_LAST();
}

extern "C" int ctx_texmap_fwd(const float* uv, const float* tex, const float* mask, const float* bg, float* out,
                              int64_t B, int64_t N, int tex_batch, int C, int H, int W, int mode, void* stream) {
  if (B < 0 || N < 0 || C < 1 || H < 1 || W < 1 || mode < 0 || mode > 2) return CTX_ERR_BAD_ARG;
  if (tex_batch != 1 && tex_batch != B) return CTX_ERR_BAD_ARG;
  if (B * N == 0) return 0;
  if (!uv || !tex || !out) return CTX_ERR_BAD_ARG;
  if (mode == 2) {
    ctx::texmap_bicubic_kernel<false><<<ctx::texmap_grid(B * N), 256, 0, (cudaStream_t)stream>>>(
        (const float2*)uv, tex, mask, bg, out, nullptr, B, N, tex_batch, C, H, W);
    CTX_RETURN_LAST();
  }
  if (B <= 65535 && ctx::texmap_vec_ok(N, C, {uv, mask, out})) {
    dim3 grid((unsigned)ctx::texmap_grid(N / 4), (unsigned)B);
    ctx::texmap_rgb4_kernel<false><<<grid, 256, 0, (cudaStream_t)stream>>>(
        (const float4*)uv, tex, (const float4*)mask, bg, (float4*)out, nullptr, N, tex_batch, H, W, mode);
    CTX_RETURN_LAST();
  }
  ctx::texmap_fwd_kernel<<<ctx::texmap_grid(B * N), 256, 0, (cudaStream_t)stream>>>(
      (const float2*)uv, tex, mask, bg, out, B, N, tex_batch, C, H, W, mode);
  CTX_RETURN_LAST();
}

// g_tex is ACCUMULATED into (zero it first for a fresh gradient)
extern "C" int ctx_texmap_bwd(const float* uv, const float* mask, const float* g_out, float* g_tex, int64_t B,
                              int64_t N, int tex_batch, int C, int H, int W, int mode, void* stream) {
  if (B < 0 || N < 0 || C < 1 || H < 1 || W < 1 || mode < 0 || mode > 2) return CTX_ERR_BAD_ARG;
  if (tex_batch != 1 && tex_batch != B) return CTX_ERR_BAD_ARG;
  if (B * N == 0) return 0;
  if (!uv || !g_out || !g_tex) return CTX_ERR_BAD_ARG;
  if (mode == 2) {
    ctx::texmap_bicubic_kernel<true><<<ctx::texmap_grid(B * N), 256, 0, (cudaStream_t)stream>>>(
        (const float2*)uv, nullptr, mask, nullptr, const_cast<float*>(g_out), g_tex, B, N, tex_batch, C, H, W);
    CTX_RETURN_LAST();
  }
  if (B <= 65535 && ctx::texmap_vec_ok(N, C, {uv, mask, g_out})) {
    dim3 grid((unsigned)ctx::texmap_grid(N / 4), (unsigned)B);
    ctx::texmap_rgb4_kernel<true><<<grid, 256, 0, (cudaStream_t)stream>>>(
        (const float4*)uv, nullptr, (const float4*)mask, nullptr, (float4*)const_cast<float*>(g_out), g_tex, N,
        tex_batch, H, W, mode);
    CTX_RETURN_LAST();
  }
  ctx::texmap_bwd_kernel<<<ctx::texmap_grid(B * N), 256, 0, (cudaStream_t)stream>>>(
      (const float2*)uv, mask, g_out, g_tex, B, N, tex_batch, C, H, W, mode);
  CTX_RETURN_LAST();
}
