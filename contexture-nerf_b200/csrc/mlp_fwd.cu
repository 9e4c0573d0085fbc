// Fused coordinate-MLP forward on tcgen05 / TMEM (sm_100a).
//
// One persistent CTA per SM.  Per iteration a CTA owns 256 points = two
// 128-row accumulator tiles (2 x 256 TMEM columns).  For every layer the weight
// matrix is streamed ONCE per iteration from L2 in 16 KB K-chunks by 1-D bulk
// TMA copies (4-stage mbarrier ring) and each chunk feeds the MMAs of both
// tiles; activations never leave the SM: the epilogue warps read the fp32
// accumulator from TMEM, add the bias, apply ReLU, round to bf16 and write the
// next layer's A operand straight back into shared memory.  The positional
// encoding of the points (o + d*z, L=10) and of the view directions (L=4) is
// produced directly in shared memory as the layer-0 / view-layer A operand.
// The 1- to 4-wide heads (alpha, rgb, output_linear) are folded into the
// preceding epilogue on the CUDA cores.
//
// Warp roles (320 threads): warp 0 = weight-stream producer (one lane),
// warp 1 = TMEM allocator + MMA issuer (one lane), warps 2-9 = encode/epilogue
// (4 warps per tile; warp%4 selects the TMEM lane quarter).
//
// Reference semantics: NeRF2D.forward, /root/reference/src/run_nerf_helpers.py:106-135
// (and the commented view branch :117-127); Embedder.embed :44-45.
#include "mlp_common.cuh"

namespace ctx {

// encode d-dim coordinate vector into channels [x | sin f0 x | cos f0 x | ...] padded with zeros to `pad`
template <int PAD, int MAXL>
__device__ __forceinline__ void encode_row(uint8_t* tile, int row, const float* xyz, int L, bool valid,
                                           uint8_t* gtile) {
  constexpr int d = 3;
  float v[PAD];
#pragma unroll
  for (int i = 0; i < PAD; ++i) v[i] = 0.f;
  if (valid) {
#pragma unroll
    for (int j = 0; j < 3; ++j)
      if (j < d) v[j] = xyz[j];
#pragma unroll
    for (int k = 0; k < MAXL; ++k) {
      if (k < L) {
        const float f = (float)(1 << k);
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          if (j < d) {
            float s, c;
            fast_sincos(xyz[j] * f, s, c);
            const int b = d + k * 2 * d;
            if (b + d + j < PAD) { v[b + j] = s; v[b + d + j] = c; }
          }
        }
      }
    }
  }
#pragma unroll
  for (int c0 = 0; c0 < PAD; c0 += 8) store_row8(tile, row, c0, v + c0, false, gtile);
}

__global__ void __launch_bounds__(kMlpThreads, 1) mlp_fwd_kernel(const __grid_constant__ MlpFwdArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* h_buf = smem;                                          // kTiles x 64 KB
  uint8_t* x_buf = smem + kTiles * kHBytes;                       // kTiles x 16 KB
  uint8_t* w_buf = x_buf + kTiles * kXBytes;                      // kStages x 16 KB
  MlpSmemCtl* ctl = reinterpret_cast<MlpSmemCtl*>(w_buf + kStages * kStageBytes);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const CtxMlpNet& net = a.net;
  const int64_t n_iters_total = ceil_div(a.P, (int64_t)kTileM * kTiles);

  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) { tc::mbar_init(&ctl->full[s], 1); tc::mbar_init(&ctl->empty[s], 1); }
    for (int t = 0; t < kTiles; ++t) {
      tc::mbar_init(&ctl->acc_full[t], 1);
      tc::mbar_init(&ctl->act_ready[t], kEpiThreadsPerTile);
    }
    tc::mbar_fence_init();
  }
  if (warp == 1) tc::tmem_alloc(&ctl->tmem_base, 512);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = ctl->tmem_base;

  if (warp == 0) {
    // ===================== weight-stream producer =====================
    if (lane == 0) {
      uint32_t g = 0;  // running chunk counter
      for (int64_t it = blockIdx.x; it < n_iters_total; it += gridDim.x) {
        for (int l = 0; l < net.n_layers; ++l) {
          const CtxMlpLayer& L = net.L[l];
          const int nchunks = L.n_x_pre + L.n_h + L.n_x_post;
          const uint32_t bytes = (uint32_t)L.N * CTX_MLP_KC * 2;
          for (int c = 0; c < nchunks; ++c, ++g) {
            const int s = g % kStages;
            const uint32_t ph = (g / kStages) & 1;
            tc::mbar_wait(&ctl->empty[s], ph ^ 1);
            tc::mbar_arrive_expect_tx(&ctl->full[s], bytes);
            tc::bulk_g2s(w_buf + s * kStageBytes, a.wpacked + L.w_off + (size_t)c * bytes, bytes, &ctl->full[s]);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ========================= MMA issuer ==============================
    if (lane == 0) {
      uint32_t g = 0;
      uint32_t act_phase = 0;  // same sequence for both tiles
      for (int64_t it = blockIdx.x; it < n_iters_total; it += gridDim.x) {
        for (int l = 0; l < net.n_layers; ++l) {
          const CtxMlpLayer& L = net.L[l];
          const int nchunks = L.n_x_pre + L.n_h + L.n_x_post;
          const uint32_t idesc = tc::make_idesc_bf16(kTileM, L.N, 0, 0);
          const uint32_t b_lbo = (uint32_t)L.N * 16;
          for (int c = 0; c < nchunks; ++c, ++g) {
            const int s = g % kStages;
            tc::mbar_wait(&ctl->full[s], (g / kStages) & 1);
            tc::tc_fence_after();
            // A source of this chunk
            uint32_t a_off;  // byte offset inside the tile's buffers
            bool from_x;
            if (c < L.n_x_pre) { from_x = true; a_off = c * 4 * kK8Stride; }
            else if (c < L.n_x_pre + L.n_h) { from_x = false; a_off = (c - L.n_x_pre) * 4 * kK8Stride; }
            else { from_x = true; a_off = (c - L.n_x_pre - L.n_h) * 4 * kK8Stride; }
            const uint32_t b_base = tc::smem_u32(w_buf + s * kStageBytes);
#pragma unroll
            for (int t = 0; t < kTiles; ++t) {
              if (c == 0) {
                tc::mbar_wait(&ctl->act_ready[t], act_phase);
                tc::tc_fence_after();
              }
              const uint32_t a_base = tc::smem_u32(from_x ? (x_buf + t * kXBytes) : (h_buf + t * kHBytes)) + a_off;
#pragma unroll
              for (int kk = 0; kk < 2; ++kk) {
                const uint64_t da = tc::make_smem_desc(a_base + kk * 2 * kK8Stride, kK8Stride, 128);
                const uint64_t db = tc::make_smem_desc(b_base + kk * 2 * b_lbo, b_lbo, 128);
                tc::mma_bf16_ss(tmem + t * CTX_MLP_W, da, db, idesc, (c > 0 || kk > 0) ? 1u : 0u);
              }
              if (c == nchunks - 1) tc::mma_commit(&ctl->acc_full[t]);
            }
            tc::mma_commit(&ctl->empty[s]);
          }
          act_phase ^= 1;
        }
      }
    }
  } else {
    // ===================== encode + epilogue warps =====================
    const int t = (warp - 2) >> 2;            // tile handled by this warp
    const int q = warp & 3;                   // TMEM lane quarter this warp may access
    const int row = q * 32 + lane;
    uint8_t* my_h = h_buf + t * kHBytes;
    uint8_t* my_x = x_buf + t * kXBytes;
    const uint32_t my_acc = tmem + t * CTX_MLP_W + ((uint32_t)(q * 32) << 16);
    const float* fp = a.fparams;
    uint32_t acc_phase = 0;
    const bool has_views = net.in_views > 0;

    for (int64_t it = blockIdx.x; it < n_iters_total; it += gridDim.x) {
      const int64_t tile_idx = it * kTiles + t;
      const int64_t p = tile_idx * kTileM + row;
      const bool valid = p < a.P;
      uint8_t* rec = a.acts ? a.acts + (size_t)tile_idx * net.act_tile_bytes : nullptr;
      float dirs[3] = {0.f, 0.f, 0.f};
      // ---- encode the point into the layer-0 A operand ----
      if (a.mode == 1) {
        float xyz[3] = {0.f, 0.f, 0.f};
        if (valid) {
          const int64_t r = p / a.S;
          const float zv = a.z[p];
#pragma unroll
          for (int j = 0; j < 3; ++j) {
            xyz[j] = a.rays_o[r * 3 + j] + a.rays_d[r * 3 + j] * zv;
            if (has_views) dirs[j] = a.viewdirs[r * 3 + j];
          }
        }
        encode_row<CTX_MLP_XP_PAD, 10>(my_x, row, xyz, a.L_pts, valid, rec ? rec + net.xp_slot : nullptr);
      } else {
        float v[CTX_MLP_XP_PAD];
#pragma unroll
        for (int i = 0; i < CTX_MLP_XP_PAD; ++i)
          v[i] = (valid && i < net.in_pts) ? __ldg(a.x + p * a.x_ld + i) : 0.f;
#pragma unroll
        for (int c0 = 0; c0 < CTX_MLP_XP_PAD; c0 += 8)
          store_row8(my_x, row, c0, v + c0, false, rec ? rec + net.xp_slot : nullptr);
      }
      tc::fence_proxy_async_smem();
      tc::mbar_arrive(&ctl->act_ready[t]);

      float alpha = 0.f;
      for (int l = 0; l < net.n_layers; ++l) {
        const CtxMlpLayer& L = net.L[l];
        tc::mbar_wait(&ctl->acc_full[t], acc_phase);
        acc_phase ^= 1;
        tc::tc_fence_after();
        const bool is_final = (L.epi == CTX_EPI_FINAL_VIEWS || L.epi == CTX_EPI_FINAL_OUT);
        const bool write_h = !is_final || rec != nullptr;
        float head[4] = {0.f, 0.f, 0.f, 0.f};
        const float* hw = fp + net.head_off;
        for (int cb = 0; cb < L.N / 32; ++cb) {
          uint32_t vr[32];
          tc::tmem_ld32(my_acc + cb * 32, vr);
          const float4* b4 = reinterpret_cast<const float4*>(fp + L.bias_off + cb * 32);
          float bias[32];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 b = __ldg(b4 + j);
            bias[4 * j] = b.x; bias[4 * j + 1] = b.y; bias[4 * j + 2] = b.z; bias[4 * j + 3] = b.w;
          }
          tc::tmem_wait_ld();
          float v[32];
          uint32_t neg = 0;   // bit (31-j) = sign of pre-activation j (the ReLU mask the dgrad kernel reads)
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            v[j] = __uint_as_float(vr[j]) + bias[j];
            if (rec) neg = __funnelshift_l(__float_as_uint(v[j]), neg, 1);
            if (L.relu) v[j] = fmaxf(v[j], 0.f);
          }
          if (rec && L.mask_slot >= 0)
            reinterpret_cast<uint32_t*>(rec + L.mask_slot)[row * (L.N / 32) + cb] = neg;
          if (L.epi == CTX_EPI_HIDDEN_ALPHA) {
#pragma unroll
            for (int j = 0; j < 32; ++j) alpha = fmaf(bf16_round(v[j]), __ldg(hw + cb * 32 + j), alpha);
          } else if (L.epi == CTX_EPI_FINAL_VIEWS) {
            const float* wr = hw + 260;  // W_rgb [3][128]
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const float hv = bf16_round(v[j]);
              head[0] = fmaf(hv, __ldg(wr + cb * 32 + j), head[0]);
              head[1] = fmaf(hv, __ldg(wr + 128 + cb * 32 + j), head[1]);
              head[2] = fmaf(hv, __ldg(wr + 256 + cb * 32 + j), head[2]);
            }
          } else if (L.epi == CTX_EPI_FINAL_OUT) {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const float hv = bf16_round(v[j]);
#pragma unroll
              for (int o = 0; o < 4; ++o) head[o] = fmaf(hv, __ldg(hw + o * 256 + cb * 32 + j), head[o]);
            }
          }
          if (write_h) {
#pragma unroll
            for (int j = 0; j < 32; j += 8)
              store_row8(is_final ? nullptr : my_h, row, cb * 32 + j, v + j, false, rec ? rec + L.act_slot : nullptr);
          }
        }
        if (L.epi == CTX_EPI_HIDDEN_ALPHA) {
          alpha += __ldg(hw + 256);
          // the point encoding is dead from here on: encode the view direction into the x buffer
          if (a.mode == 1) {
            encode_row<CTX_MLP_XD_PAD, 4>(my_x, row, dirs, a.L_dirs, valid, rec ? rec + net.xd_slot : nullptr);
          } else {
            float vv[CTX_MLP_XD_PAD];
#pragma unroll
            for (int i = 0; i < CTX_MLP_XD_PAD; ++i)
              vv[i] = (valid && i < net.in_views) ? __ldg(a.x + p * a.x_ld + net.in_pts + i) : 0.f;
#pragma unroll
            for (int c0 = 0; c0 < CTX_MLP_XD_PAD; c0 += 8)
              store_row8(my_x, row, c0, vv + c0, false, rec ? rec + net.xd_slot : nullptr);
          }
        }
        if (is_final) {
          if (valid) {
            if (L.epi == CTX_EPI_FINAL_VIEWS) {
              const float* br = hw + 260 + 384;
              float4 o4 = make_float4(head[0] + __ldg(br), head[1] + __ldg(br + 1), head[2] + __ldg(br + 2), alpha);
              *reinterpret_cast<float4*>(a.out + p * 4) = o4;
            } else {
              const float* bo = hw + 1024;
#pragma unroll
              for (int o = 0; o < 4; ++o)
                if (o < net.out_ch) a.out[p * net.out_ch + o] = head[o] + __ldg(bo + o);
            }
          }
        }
        if (!is_final) tc::fence_proxy_async_smem();
        tc::tc_fence_before();
        if (!is_final) tc::mbar_arrive(&ctl->act_ready[t]);
      }
    }
  }

  tc::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tc::tmem_dealloc(tmem, 512);
  }
}

}  // namespace ctx

extern "C" int ctx_mlp_fwd(const void* net_host, const void* wpacked, const float* fparams, int mode,
                           const float* x, int x_ld, const float* rays_o, const float* rays_d,
                           const float* viewdirs, const float* z, int S, int L_pts, int L_dirs, int64_t P,
                           float* out, void* acts, void* stream) {
  if (!net_host || !wpacked || !fparams || !out || P < 0) return CTX_ERR_BAD_ARG;
  if (P == 0) return 0;
  ctx::MlpFwdArgs a;
  a.net = *reinterpret_cast<const CtxMlpNet*>(net_host);
  if (mode == 0) {
    if (!x || x_ld < a.net.in_pts + a.net.in_views) return CTX_ERR_BAD_ARG;
  } else if (mode == 1) {
    if (!rays_o || !rays_d || !z || S < 1 || (a.net.in_views > 0 && !viewdirs)) return CTX_ERR_BAD_ARG;
    if (a.net.in_pts != 3 * (1 + 2 * L_pts) || L_pts > 10) return CTX_ERR_UNSUPPORTED;
    if (a.net.in_views > 0 && (a.net.in_views != 3 * (1 + 2 * L_dirs) || L_dirs > 4)) return CTX_ERR_UNSUPPORTED;
  } else {
    return CTX_ERR_BAD_ARG;
  }
  a.wpacked = (const uint8_t*)wpacked; a.fparams = fparams; a.mode = mode; a.x = x; a.x_ld = x_ld;
  a.rays_o = rays_o; a.rays_d = rays_d; a.viewdirs = viewdirs; a.z = z; a.S = S; a.L_pts = L_pts;
  a.L_dirs = L_dirs; a.P = P; a.out = out; a.acts = (uint8_t*)acts; a.prof = nullptr; a.debug = 0;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(ctx::mlp_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)ctx::kMlpSmemBytes);
    if (e != cudaSuccess) return (int)e;
    attr_set = true;
  }
  const int64_t iters = ctx::ceil_div(P, (int64_t)ctx::kTileM * ctx::kTiles);
  const int grid = (int)(iters < ctx::kNumSMs ? iters : ctx::kNumSMs);
  ctx::mlp_fwd_kernel<<<grid, ctx::kMlpThreads, ctx::kMlpSmemBytes, (cudaStream_t)stream>>>(a);
  CTX_RETURN_LAST();
}
