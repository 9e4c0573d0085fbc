// Hand-written backward of the coordinate MLP on tcgen05 / TMEM (sm_100a).
//
//   dgrad kernel : same persistent two-tile skeleton as the forward.  Per layer
//     (last to first) dH_in = dZ * W (transposed bf16 weight stream, K = output
//     features), the epilogue applies the ReLU sign mask recorded by the forward
//     (1 bit / activation), writes dZ of the previous layer back to shared memory
//     as the next A operand and bulk-stores it to the dZ record for wgrad.  The
//     tiny heads (rgb, alpha, output_linear) are differentiated on the CUDA cores.
//   wgrad kernel : dW = dZ^T * A summed over all points -- a split-K GEMM whose
//     K dimension is the point index.  Both operands are the [points x features]
//     tile images written by forward / dgrad, consumed as MN-major operands, so
//     no transposition pass exists anywhere.  Each CTA owns one (layer, segment)
//     job and a slice of the point tiles, accumulates in TMEM, and flushes with
//     fp32 red.global.add; bias gradients are column sums of dZ taken from the
//     shared-memory tile while the MMAs run.
//
// Gradients flow to the parameters only (the encoded inputs are data).
// Reference semantics: autograd through NeRF2D.forward,
// /root/reference/src/run_nerf_helpers.py:106-135.
#include "mlp_common.cuh"
#include <string.h>
#include <stdlib.h>

// 2-CTA dgrad (mlp_bwd2.cu)
int ctx_launch_dgrad2(const CtxMlpNet& net, const void* wtpacked, const float* fparams, const float* g_out,
                      const void* acts, void* dacts, int64_t P, cudaStream_t st);

namespace ctx {

// ============================== dgrad ======================================
struct DgradArgs {
  CtxMlpNet net;
  const uint8_t* wtpacked;
  const float* fparams;
  const float* g_out;   // [P, out_ch]
  const uint8_t* acts;  // forward records (ReLU masks)
  uint8_t* dacts;       // dZ records (same slot offsets)
  int64_t P;
  int n_steps;
  int step_src[CTX_MLP_MAX_LAYERS];  // layer whose W^T is applied (A operand = dZ of this layer)
  int step_dst[CTX_MLP_MAX_LAYERS];  // layer whose dZ the step produces
};

__device__ __forceinline__ float mask_apply(float v, uint32_t neg, int j) {
  return ((neg >> (31 - j)) & 1u) ? 0.f : v;
}

__global__ void __launch_bounds__(kMlpThreads, 1) mlp_dgrad_kernel(const __grid_constant__ DgradArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* h_buf = smem;
  uint8_t* x_buf = smem + kTiles * kHBytes;     // staging of the 16-channel g_out image
  uint8_t* w_buf = x_buf + kTiles * kXBytes;
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
    if (lane == 0) {
      uint32_t g = 0;
      for (int64_t it = blockIdx.x; it < n_iters_total; it += gridDim.x) {
        for (int si = 0; si < a.n_steps; ++si) {
          const CtxMlpLayer& S = net.L[a.step_src[si]];
          const int nchunks = S.N / CTX_MLP_KC;
          const uint32_t bytes = 256 * CTX_MLP_KC * 2;
          for (int c = 0; c < nchunks; ++c, ++g) {
            const int s = g % kStages;
            tc::mbar_wait(&ctl->empty[s], ((g / kStages) & 1) ^ 1);
            tc::mbar_arrive_expect_tx(&ctl->full[s], bytes);
            tc::bulk_g2s(w_buf + s * kStageBytes, a.wtpacked + S.wt_off + (size_t)c * bytes, bytes, &ctl->full[s]);
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      uint32_t g = 0, act_phase = 0;
      const uint32_t idesc = tc::make_idesc_bf16(kTileM, 256, 0, 0);
      for (int64_t it = blockIdx.x; it < n_iters_total; it += gridDim.x) {
        for (int si = 0; si < a.n_steps; ++si) {
          const CtxMlpLayer& S = net.L[a.step_src[si]];
          const int nchunks = S.N / CTX_MLP_KC;
          for (int c = 0; c < nchunks; ++c, ++g) {
            const int s = g % kStages;
            tc::mbar_wait(&ctl->full[s], (g / kStages) & 1);
            tc::tc_fence_after();
            const uint32_t b_base = tc::smem_u32(w_buf + s * kStageBytes);
#pragma unroll
            for (int t = 0; t < kTiles; ++t) {
              if (c == 0) {
                tc::mbar_wait(&ctl->act_ready[t], act_phase);
                tc::tc_fence_after();
              }
              const uint32_t a_base = tc::smem_u32(h_buf + t * kHBytes) + c * 4 * kK8Stride;
#pragma unroll
              for (int kk = 0; kk < 2; ++kk) {
                const uint64_t da = tc::make_smem_desc(a_base + kk * 2 * kK8Stride, kK8Stride, 128);
                const uint64_t db = tc::make_smem_desc(b_base + kk * 2 * 4096, 4096, 128);
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
    const int t = (warp - 2) >> 2;
    const int q = warp & 3;
    const int row = q * 32 + lane;
    uint8_t* my_h = h_buf + t * kHBytes;
    const uint32_t my_acc = tmem + t * CTX_MLP_W + ((uint32_t)(q * 32) << 16);
    const float* hw = a.fparams + net.head_off;
    const bool has_views = net.in_views > 0;
    uint32_t acc_phase = 0;
    const CtxMlpLayer& Last = net.L[net.n_layers - 1];

    for (int64_t it = blockIdx.x; it < n_iters_total; it += gridDim.x) {
      const int64_t tile_idx = it * kTiles + t;
      const int64_t p = tile_idx * kTileM + row;
      const bool valid = p < a.P;
      const uint8_t* rec = a.acts + (size_t)tile_idx * net.act_tile_bytes;
      uint8_t* drec = a.dacts + (size_t)tile_idx * net.act_tile_bytes;
      // ---- head: dZ of the last GEMM layer from g_out, on the CUDA cores ----
      float g[4] = {0.f, 0.f, 0.f, 0.f};
      if (valid) {
        if (net.out_ch == 4) {
          const float4 g4 = *reinterpret_cast<const float4*>(a.g_out + p * 4);
          g[0] = g4.x; g[1] = g4.y; g[2] = g4.z; g[3] = g4.w;
        } else {
          for (int o = 0; o < net.out_ch; ++o) g[o] = a.g_out[p * net.out_ch + o];
        }
      }
      const float d_alpha = g[3];
      {
        float gv[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) gv[i] = i < 4 ? g[i] : 0.f;
        store_row8(nullptr, row, 0, gv, false, drec + net.gout_slot);
        store_row8(nullptr, row, 8, gv + 8, false, drec + net.gout_slot);
      }
      {
        const int nw = Last.N / 32;
        const uint32_t* mrow = reinterpret_cast<const uint32_t*>(rec + Last.mask_slot) + row * nw;
        for (int cb = 0; cb < nw; ++cb) {
          const uint32_t neg = __ldg(mrow + cb);
          float v[32];
          if (has_views) {
            const float* wr = hw + 260;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const int c = cb * 32 + j;
              v[j] = g[0] * __ldg(wr + c) + g[1] * __ldg(wr + 128 + c) + g[2] * __ldg(wr + 256 + c);
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const int c = cb * 32 + j;
              v[j] = g[0] * __ldg(hw + c) + g[1] * __ldg(hw + 256 + c) + g[2] * __ldg(hw + 512 + c) +
                     g[3] * __ldg(hw + 768 + c);
            }
          }
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = mask_apply(v[j], neg, j);
#pragma unroll
          for (int j = 0; j < 32; j += 8) store_row8(my_h, row, cb * 32 + j, v + j, false, drec + Last.act_slot);
        }
      }
      tc::fence_proxy_async_smem();
      tc::mbar_arrive(&ctl->act_ready[t]);

      for (int si = 0; si < a.n_steps; ++si) {
        const CtxMlpLayer& Dst = net.L[a.step_dst[si]];
        // prefetch this row's ReLU mask words while the MMAs run
        uint32_t mw[8];
        if (Dst.relu) {
          const uint4* m4 = reinterpret_cast<const uint4*>(rec + Dst.mask_slot) + row * 2;
          const uint4 m0 = __ldg(m4), m1 = __ldg(m4 + 1);
          mw[0] = m0.x; mw[1] = m0.y; mw[2] = m0.z; mw[3] = m0.w;
          mw[4] = m1.x; mw[5] = m1.y; mw[6] = m1.z; mw[7] = m1.w;
        } else {
#pragma unroll
          for (int i = 0; i < 8; ++i) mw[i] = 0u;
        }
        tc::mbar_wait(&ctl->acc_full[t], acc_phase);
        acc_phase ^= 1;
        tc::tc_fence_after();
        const bool add_alpha = (Dst.epi == CTX_EPI_HIDDEN_ALPHA);
#pragma unroll
        for (int cb = 0; cb < 8; ++cb) {
          uint32_t vr[32];
          tc::tmem_ld32(my_acc + cb * 32, vr);
          tc::tmem_wait_ld();
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            float x = __uint_as_float(vr[j]);
            if (add_alpha) x = fmaf(d_alpha, __ldg(hw + cb * 32 + j), x);
            v[j] = mask_apply(x, mw[cb], j);
          }
#pragma unroll
          for (int j = 0; j < 32; j += 8)
            store_row8(si + 1 < a.n_steps ? my_h : nullptr, row, cb * 32 + j, v + j, false, drec + Dst.act_slot);
        }
        tc::fence_proxy_async_smem();
        tc::tc_fence_before();
        if (si + 1 < a.n_steps) tc::mbar_arrive(&ctl->act_ready[t]);
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

// ============================== wgrad ======================================
constexpr int kWgUnitBytes = 65536;
constexpr int kWgUnits = 3;
constexpr int kWgThreads = 192;     // warp 0 producer, warp 1 MMA, warps 2-5 column sums + flush
constexpr int kWgMaxJobs = 48;
constexpr size_t kWgSmemBytes = (size_t)kWgUnits * kWgUnitBytes + 256;

struct WgJob {
  int a_slot, a_ch, a_dz;       // A operand tile: record offset, channels (M, multiple of 128), from dZ records?
  int b_slot, b_ch, b_dz;       // B operand tile: channels = N (multiple of 16)
  float* out; int ld_out; int col_off;
  int transposed;               // 1: out[(n-n_lo)*ld + col_off + m]   0: out[m*ld + col_off + (n-n_lo)]
  int m_valid, n_lo, n_hi;
  float* bias_out; int bias_from_b; int bias_lo, bias_hi;   // column sums of the dZ operand -> bias_out[c - bias_lo]
  int cta_begin, cta_count;
};
struct WgradArgs {
  int n_jobs;
  WgJob job[kWgMaxJobs];
  const uint8_t* acts; const uint8_t* dacts;
  int tile_bytes;
  int64_t n_tiles;
};
struct __align__(8) WgCtl {
  uint64_t full[kWgUnits], empty[kWgUnits], done;
  uint32_t tmem_base;
};

__global__ void __launch_bounds__(kWgThreads, 1) mlp_wgrad_kernel(const __grid_constant__ WgradArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  WgCtl* ctl = reinterpret_cast<WgCtl*>(smem + kWgUnits * kWgUnitBytes);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // which job does this CTA serve?
  int ji = 0;
  for (int j = 0; j < a.n_jobs; ++j)
    if ((int)blockIdx.x >= a.job[j].cta_begin && (int)blockIdx.x < a.job[j].cta_begin + a.job[j].cta_count) ji = j;
  const WgJob& J = a.job[ji];
  const int split = blockIdx.x - J.cta_begin;
  const int m_halves = J.a_ch / 128;
  const uint32_t a_bytes = 128u * J.a_ch * 2, b_bytes = 128u * J.b_ch * 2;
  int64_t my_tiles = 0;
  if (split < a.n_tiles) my_tiles = (a.n_tiles - split + J.cta_count - 1) / J.cta_count;

  if (tid == 0) {
    for (int s = 0; s < kWgUnits; ++s) { tc::mbar_init(&ctl->full[s], 1); tc::mbar_init(&ctl->empty[s], 5); }
    tc::mbar_init(&ctl->done, 1);
    tc::mbar_fence_init();
  }
  if (warp == 1) tc::tmem_alloc(&ctl->tmem_base, 512);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = ctl->tmem_base;

  if (warp == 0) {
    if (lane == 0) {
      uint32_t u = 0;
      for (int64_t i = 0; i < my_tiles; ++i) {
        const int64_t tile = split + i * J.cta_count;
        const size_t base = (size_t)tile * a.tile_bytes;
        for (int which = 0; which < 2; ++which, ++u) {
          const int s = u % kWgUnits;
          tc::mbar_wait(&ctl->empty[s], ((u / kWgUnits) & 1) ^ 1);
          const uint8_t* src = which == 0 ? ((J.a_dz ? a.dacts : a.acts) + base + J.a_slot)
                                          : ((J.b_dz ? a.dacts : a.acts) + base + J.b_slot);
          const uint32_t bytes = which == 0 ? a_bytes : b_bytes;
          tc::mbar_arrive_expect_tx(&ctl->full[s], bytes);
          tc::bulk_g2s(smem + s * kWgUnitBytes, src, bytes, &ctl->full[s]);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = tc::make_idesc_bf16(128, J.b_ch, 1, 1);
      uint32_t u = 0;
      for (int64_t i = 0; i < my_tiles; ++i, u += 2) {
        const int sa = u % kWgUnits, sb = (u + 1) % kWgUnits;
        tc::mbar_wait(&ctl->full[sa], (u / kWgUnits) & 1);
        tc::mbar_wait(&ctl->full[sb], ((u + 1) / kWgUnits) & 1);
        tc::tc_fence_after();
        const uint32_t a_base = tc::smem_u32(smem + sa * kWgUnitBytes);
        const uint32_t b_base = tc::smem_u32(smem + sb * kWgUnitBytes);
        for (int mh = 0; mh < m_halves; ++mh) {
#pragma unroll
          for (int k16 = 0; k16 < 8; ++k16) {
            // MN-major: SBO = 2048 (next 8 channels), LBO = 128 (next 8 points); 16 points = 256 B
            const uint64_t da = tc::make_smem_desc(a_base + mh * 16 * 2048 + k16 * 256, 128, 2048);
            const uint64_t db = tc::make_smem_desc(b_base + k16 * 256, 128, 2048);
            tc::mma_bf16_ss(tmem + mh * 256, da, db, idesc, (i > 0 || k16 > 0) ? 1u : 0u);
          }
        }
        tc::mma_commit(&ctl->empty[sa]);
        tc::mma_commit(&ctl->empty[sb]);
      }
      tc::mma_commit(&ctl->done);
    }
  } else {
    // ---- column sums of the dZ operand (bias gradients), then the flush ----
    const int cw = warp - 2;  // 0..3
    const bool do_bias = J.bias_out != nullptr;
    const int dz_ch = J.bias_from_b ? J.b_ch : J.a_ch;
    const int n_chunks = dz_ch / 8;
    float s_lo[8], s_hi[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { s_lo[i] = 0.f; s_hi[i] = 0.f; }
    uint32_t u = 0;
    for (int64_t i = 0; i < my_tiles; ++i, u += 2) {
      const int sa = u % kWgUnits, sb = (u + 1) % kWgUnits;
      // always wait for both units: keeps these warps within one ring phase of the MMA issuer
      tc::mbar_wait(&ctl->full[sa], (u / kWgUnits) & 1);
      tc::mbar_wait(&ctl->full[sb], ((u + 1) / kWgUnits) & 1);
      if (do_bias) {
        const int sd = J.bias_from_b ? sb : sa;
        const uint8_t* tile = smem + sd * kWgUnitBytes;
#pragma unroll
        for (int ci = 0; ci < 8; ++ci) {
          const int c = cw + ci * 4;
          if (c < n_chunks) {
#pragma unroll
            for (int gp = 0; gp < 16; ++gp) {
              const uint32_t w2 = *reinterpret_cast<const uint32_t*>(tile + c * 2048 + gp * 128 + lane * 4);
              s_lo[ci] += __uint_as_float(w2 << 16);
              s_hi[ci] += __uint_as_float(w2 & 0xffff0000u);
            }
          }
        }
      }
      __syncwarp();
      if (lane == 0) { tc::mbar_arrive(&ctl->empty[sa]); tc::mbar_arrive(&ctl->empty[sb]); }
    }
    if (do_bias) {
#pragma unroll
      for (int ci = 0; ci < 8; ++ci) {
        float lo = s_lo[ci], hi = s_hi[ci];
#pragma unroll
        for (int o = 4; o < 32; o <<= 1) {
          lo += __shfl_xor_sync(CTX_FULL_MASK, lo, o);
          hi += __shfl_xor_sync(CTX_FULL_MASK, hi, o);
        }
        const int c = cw + ci * 4;
        if (c < n_chunks && lane < 4 && my_tiles > 0) {
          const int ch = c * 8 + lane * 2;
          if (ch >= J.bias_lo && ch < J.bias_hi) atomicAdd(J.bias_out + ch - J.bias_lo, lo);
          if (ch + 1 >= J.bias_lo && ch + 1 < J.bias_hi) atomicAdd(J.bias_out + ch + 1 - J.bias_lo, hi);
        }
      }
    }
    // ---- flush the TMEM accumulators: warp%4 selects the lane quarter ----
    if (my_tiles > 0) {
      tc::mbar_wait(&ctl->done, 0);
      tc::tc_fence_after();
      const int q = warp & 3;
      for (int mh = 0; mh < m_halves; ++mh) {
        const int m = mh * 128 + q * 32 + lane;
        for (int cb = 0; cb < (J.b_ch + 31) / 32; ++cb) {
          uint32_t vr[32];
          if (J.b_ch - cb * 32 >= 32) {
            tc::tmem_ld32(tmem + mh * 256 + cb * 32 + ((uint32_t)(q * 32) << 16), vr);
          } else {  // N = 16 heads: load 16 columns
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
                "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                : "=r"(vr[0]), "=r"(vr[1]), "=r"(vr[2]), "=r"(vr[3]), "=r"(vr[4]), "=r"(vr[5]), "=r"(vr[6]),
                  "=r"(vr[7]), "=r"(vr[8]), "=r"(vr[9]), "=r"(vr[10]), "=r"(vr[11]), "=r"(vr[12]), "=r"(vr[13]),
                  "=r"(vr[14]), "=r"(vr[15])
                : "r"(tmem + mh * 256 + cb * 32 + ((uint32_t)(q * 32) << 16))
                : "memory");
#pragma unroll
            for (int j = 16; j < 32; ++j) vr[j] = 0u;
          }
          tc::tmem_wait_ld();
          if (m < J.m_valid) {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const int n = cb * 32 + j;
              if (n >= J.n_lo && n < J.n_hi) {
                float* dst = J.transposed ? (J.out + (size_t)(n - J.n_lo) * J.ld_out + J.col_off + m)
                                          : (J.out + (size_t)m * J.ld_out + J.col_off + (n - J.n_lo));
                atomicAdd(dst, __uint_as_float(vr[j]));
              }
            }
          }
        }
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

// grads: HOST array of DEVICE pointers in the order of ctx_mlp_pack's `params`
// (gradients are ACCUMULATED into them: zero them first for a fresh gradient).
extern "C" int ctx_mlp_bwd(const void* net_host, const void* wtpacked, const float* fparams,
                           const float* g_out, const void* acts, void* dacts, int64_t P, float* const* grads,
                           int n_grads, void* stream) {
  if (!net_host || !wtpacked || !fparams || !g_out || !acts || !dacts || !grads || P < 0) return CTX_ERR_BAD_ARG;
  if (P == 0) return 0;
  const CtxMlpNet& net = *reinterpret_cast<const CtxMlpNet*>(net_host);
  const bool views = net.in_views > 0;
  const int D = views ? net.n_layers - 2 : net.n_layers;
  if (n_grads != 2 * D + (views ? 8 : 2)) return CTX_ERR_BAD_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(ctx::mlp_dgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)ctx::kMlpSmemBytes);
    if (e != cudaSuccess) return (int)e;
    e = cudaFuncSetAttribute(ctx::mlp_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)ctx::kWgSmemBytes);
    if (e != cudaSuccess) return (int)e;
    attr_set = true;
  }
  // ---------------- dgrad ----------------
  static const bool use_v1 = [] { const char* e = getenv("CTXNERF_MLP_KERNEL"); return e && e[0] == '1'; }();
  if (!use_v1) {
    const int rc = ctx_launch_dgrad2(net, wtpacked, fparams, g_out, acts, dacts, P, st);
    if (rc != 0) return rc;
  } else {
    ctx::DgradArgs a;
    a.net = net; a.wtpacked = (const uint8_t*)wtpacked; a.fparams = fparams; a.g_out = g_out;
    a.acts = (const uint8_t*)acts; a.dacts = (uint8_t*)dacts; a.P = P;
    int n = 0;
    for (int l = net.n_layers - 1; l >= 1; --l) {   // layer l's W^T produces dZ of layer l-1
      a.step_src[n] = l; a.step_dst[n] = l - 1; ++n;
    }
    a.n_steps = n;
    const int64_t iters = ctx::ceil_div(P, (int64_t)ctx::kTileM * ctx::kTiles);
    const int grid = (int)(iters < ctx::kNumSMs ? iters : ctx::kNumSMs);
    ctx::mlp_dgrad_kernel<<<grid, ctx::kMlpThreads, ctx::kMlpSmemBytes, st>>>(a);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return (int)e;
  }
  // ---------------- wgrad ----------------
  {
    ctx::WgradArgs w;
    memset(&w, 0, sizeof(w));
    w.acts = (const uint8_t*)acts; w.dacts = (const uint8_t*)dacts; w.tile_bytes = net.act_tile_bytes;
    const int64_t n_tiles = ctx::ceil_div(P, 128);
    w.n_tiles = n_tiles;
    int nj = 0;
    float cost[ctx::kWgMaxJobs];
    auto add = [&](int a_slot, int a_ch, int a_dz, int b_slot, int b_ch, int b_dz, float* out, int ld, int col_off,
                   int transposed, int m_valid, int n_lo, int n_hi, float* bias, int bias_from_b, int blo, int bhi) {
      ctx::WgJob& J = w.job[nj];
      J.a_slot = a_slot; J.a_ch = a_ch; J.a_dz = a_dz; J.b_slot = b_slot; J.b_ch = b_ch; J.b_dz = b_dz;
      J.out = out; J.ld_out = ld; J.col_off = col_off; J.transposed = transposed; J.m_valid = m_valid;
      J.n_lo = n_lo; J.n_hi = n_hi; J.bias_out = bias; J.bias_from_b = bias_from_b; J.bias_lo = blo; J.bias_hi = bhi;
      cost[nj] = (float)(a_ch + b_ch);   // HBM bytes per point decide the split, the kernel is bandwidth-bound
      ++nj;
    };
    for (int l = 0; l < net.n_layers; ++l) {
      const CtxMlpLayer& L = net.L[l];
      int pi;
      if (l < D) pi = 2 * l; else if (l == D) pi = 2 * D; else pi = 2 * D + 4;
      float* gW = grads[pi];
      float* gb = grads[pi + 1];
      int ld = 0;
      if (L.n_x_pre) ld += net.in_pts;
      const int h_col = ld;
      if (L.n_h) ld += 256;
      const int xd_col = ld;
      if (L.n_x_post) ld += net.in_views;
      bool bias_done = false;
      if (L.n_h) {  // h segment: transposed job, A = input activations (M = in), B = dZ (N = out)
        add(L.in_slot, 256, 0, L.act_slot, L.N, 1, gW, ld, h_col, 1, 256, 0, L.N, gb, 1, 0, L.N);
        bias_done = true;
      }
      if (L.n_x_pre) {  // point-encoding segment: A = dZ (M = out), B = x_p tile (N = 64, 63 real)
        add(L.act_slot, L.N, 1, net.xp_slot, CTX_MLP_XP_PAD, 0, gW, ld, 0, 0, L.N, 0, net.in_pts,
            bias_done ? nullptr : gb, 0, 0, L.N);
        bias_done = true;
      }
      if (L.n_x_post) {  // view-encoding segment
        add(L.act_slot, L.N, 1, net.xd_slot, CTX_MLP_XD_PAD, 0, gW, ld, xd_col, 0, L.N, 0, net.in_views, nullptr, 0,
            0, 0);
      }
    }
    if (views) {
      const CtxMlpLayer& H = net.L[D - 1];           // alpha_linear reads h_{D-1}
      const CtxMlpLayer& V = net.L[net.n_layers - 1];
      add(H.act_slot, 256, 0, net.gout_slot, 16, 1, grads[2 * D + 2], 256, 0, 1, 256, 3, 4, grads[2 * D + 3], 1, 3, 4);
      add(V.act_slot, 128, 0, net.gout_slot, 16, 1, grads[2 * D + 6], 128, 0, 1, 128, 0, 3, grads[2 * D + 7], 1, 0, 3);
    } else {
      const CtxMlpLayer& H = net.L[D - 1];
      add(H.act_slot, 256, 0, net.gout_slot, 16, 1, grads[2 * D], 256, 0, 1, 256, 0, net.out_ch, grads[2 * D + 1], 1,
          0, net.out_ch);
    }
    // distribute the 148 CTAs over the jobs proportionally to their HBM traffic
    float total = 0.f;
    for (int j = 0; j < nj; ++j) total += cost[j];
    int budget = ctx::kNumSMs, begin = 0;
    if (budget < nj) return CTX_ERR_UNSUPPORTED;
    int given[ctx::kWgMaxJobs];
    int used = 0;
    for (int j = 0; j < nj; ++j) {
      int c = (int)(cost[j] / total * (budget - nj)) + 1;
      if ((int64_t)c > n_tiles) c = (int)n_tiles;
      if (c < 1) c = 1;
      given[j] = c; used += c;
    }
    for (int j = 0; used < budget && j < 4 * nj; ++j) {   // hand out the remainder to the big jobs
      const int k = j % nj;
      if (cost[k] >= 384.f && (int64_t)given[k] < n_tiles) { ++given[k]; ++used; }
    }
    for (int j = 0; j < nj; ++j) { w.job[j].cta_begin = begin; w.job[j].cta_count = given[j]; begin += given[j]; }
    w.n_jobs = nj;
    ctx::mlp_wgrad_kernel<<<begin, ctx::kWgThreads, ctx::kWgSmemBytes, st>>>(w);
  }
  CTX_RETURN_LAST();
}
