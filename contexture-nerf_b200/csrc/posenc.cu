// Stand-alone multi-frequency positional encoding (the fused MLP kernel has its
// own in-shared-memory producer; this one materialises the encoding for the
// drop-in Embedder.embed and for the config-5 HBM sweep).
//
// Reference: /root/reference/src/run_nerf_helpers.py:15-45.  Channel order
// [x | sin(f0 x) | cos(f0 x) | sin(f1 x) | ...], each block d wide; the argument
// x*f is rounded to fp32 first (:38), then a full-range-reduction sincosf.
// HBM-bound: 4d B read + 4d(1+2L) B written per point (264 B at d=3, L=10).
#include "ctx_common.cuh"

namespace ctx {

constexpr int kEncTile = 64;     // points per CTA tile
constexpr int kEncThreads = 128;

__device__ __forceinline__ float enc_freq(int k, int L, int log_sampling) {
  if (log_sampling) return exp2f((float)k);                      // exact power of two
  return linspace_at(1.0f, exp2f((float)(L - 1)), L, k);         // linear in frequency
}

// One thread per (point, coordinate): it loops over the L frequencies, so no index division sits on the hot
// path and x is read once.  The tile is assembled in shared memory and leaves with 16-byte coalesced stores.
template <int D>
__global__ void __launch_bounds__(kEncThreads)
posenc_fwd_kernel(const float* __restrict__ x, float* __restrict__ out, int64_t n, int d_rt, int L,
                  int inc, int log_sampling) {
  extern __shared__ __align__(16) float tile[];
  const int d = D > 0 ? D : d_rt;
  const int C = d * (inc + 2 * L);
  const int64_t ntiles = ceil_div(n, kEncTile);
  for (int64_t tix = blockIdx.x; tix < ntiles; tix += gridDim.x) {
    const int64_t p0 = tix * kEncTile;
    const int np = (int)min((int64_t)kEncTile, n - p0);
    for (int e = threadIdx.x; e < np * d; e += kEncThreads) {
      const int pt = e / d, j = e - pt * d;
      const float xv = x[p0 * d + e];
      float* row = tile + pt * C;
      if (inc) row[j] = xv;
      float* o = row + inc * d + j;
      if (log_sampling) {
        float f = 1.0f;
        for (int k = 0; k < L; ++k, f *= 2.0f, o += 2 * d) {
          float s, c;
          sincosf(__fmul_rn(xv, f), &s, &c);
          o[0] = s; o[d] = c;
        }
      } else {
        for (int k = 0; k < L; ++k, o += 2 * d) {
          float s, c;
          sincosf(__fmul_rn(xv, enc_freq(k, L, 0)), &s, &c);
          o[0] = s; o[d] = c;
        }
      }
    }
    __syncthreads();
    float* dst = out + p0 * C;
    const int total = np * C;
    if ((total & 3) == 0) {
      const float4* t4 = reinterpret_cast<const float4*>(tile);
      float4* d4 = reinterpret_cast<float4*>(dst);
      for (int e = threadIdx.x; e < (total >> 2); e += kEncThreads) d4[e] = t4[e];
    } else {
      for (int e = threadIdx.x; e < total; e += kEncThreads) dst[e] = tile[e];
    }
    __syncthreads();
  }
}

// g_x[p,j] = inc*g[p,j] + sum_k f_k (cos(f_k x) g_sin[k] - sin(f_k x) g_cos[k])
__global__ void __launch_bounds__(kEncThreads)
posenc_bwd_kernel(const float* __restrict__ x, const float* __restrict__ g_out,
                  float* __restrict__ g_x, int64_t n, int d, int L, int inc, int log_sampling) {
  extern __shared__ __align__(16) float tile[];
  const int C = d * (inc + 2 * L);
  const int64_t ntiles = ceil_div(n, kEncTile);
  for (int64_t tix = blockIdx.x; tix < ntiles; tix += gridDim.x) {
    const int64_t p0 = tix * kEncTile;
    const int np = (int)min((int64_t)kEncTile, n - p0);
    const int total = np * C;
    const float* src = g_out + p0 * C;
    for (int e = threadIdx.x; e < total; e += kEncThreads) tile[e] = src[e];
    __syncthreads();
    for (int e = threadIdx.x; e < np * d; e += kEncThreads) {
      const int pt = e / d, j = e - pt * d;
      const float xv = x[(p0 + pt) * d + j];
      float acc = inc ? tile[pt * C + j] : 0.f;
      for (int k = 0; k < L; ++k) {
        const float f = enc_freq(k, L, log_sampling);
        float s, c;
        sincosf(__fmul_rn(xv, f), &s, &c);
        const float* row = tile + pt * C + inc * d + k * 2 * d;
        acc += f * (c * row[j] - s * row[d + j]);
      }
      g_x[(p0 + pt) * d + j] = acc;
    }
    __syncthreads();
  }
}

}  // namespace ctx

extern "C" int ctx_posenc_fwd(const float* x, float* out, int64_t n, int d, int L, int include_input,
                              int log_sampling, void* stream) {
  if (n < 0 || d < 1 || d > 16 || L < 0 || L > 32 || (!include_input && L == 0)) return CTX_ERR_BAD_ARG;
  if (n == 0) return 0;
  if (!x || !out) return CTX_ERR_BAD_ARG;
  const int C = d * ((include_input ? 1 : 0) + 2 * L);
  const size_t smem = (size_t)ctx::kEncTile * C * sizeof(float);
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(ctx::posenc_fwd_kernel<0>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
  }
  int64_t blocks = ctx::ceil_div(n, ctx::kEncTile);
  const int64_t cap = (int64_t)ctx::num_sms() * 12;
  if (blocks > cap) blocks = cap;
  cudaStream_t st = (cudaStream_t)stream;
  const int inc = include_input ? 1 : 0;
  if (d == 3 && smem <= 48 * 1024)
    ctx::posenc_fwd_kernel<3><<<(int)blocks, ctx::kEncThreads, smem, st>>>(x, out, n, d, L, inc, log_sampling);
  else if (d == 2 && smem <= 48 * 1024)
    ctx::posenc_fwd_kernel<2><<<(int)blocks, ctx::kEncThreads, smem, st>>>(x, out, n, d, L, inc, log_sampling);
  else
    ctx::posenc_fwd_kernel<0><<<(int)blocks, ctx::kEncThreads, smem, st>>>(x, out, n, d, L, inc, log_sampling);
  CTX_RETURN_LAST();
}

extern "C" int ctx_posenc_bwd(const float* x, const float* g_out, float* g_x, int64_t n, int d, int L,
                              int include_input, int log_sampling, void* stream) {
  if (n < 0 || d < 1 || d > 16 || L < 0 || L > 32 || (!include_input && L == 0)) return CTX_ERR_BAD_ARG;
  if (n == 0) return 0;
  if (!x || !g_out || !g_x) return CTX_ERR_BAD_ARG;
  const int C = d * ((include_input ? 1 : 0) + 2 * L);
  const size_t smem = (size_t)ctx::kEncTile * C * sizeof(float);
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(ctx::posenc_bwd_kernel,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
  }
  int64_t blocks = ctx::ceil_div(n, ctx::kEncTile);
  const int64_t cap = (int64_t)ctx::num_sms() * 12;
  if (blocks > cap) blocks = cap;
  ctx::posenc_bwd_kernel<<<(int)blocks, ctx::kEncThreads, smem, (cudaStream_t)stream>>>(
      x, g_out, g_x, n, d, L, include_input ? 1 : 0, log_sampling);
  CTX_RETURN_LAST();
}
