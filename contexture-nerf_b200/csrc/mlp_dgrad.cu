// dgrad of the coordinate MLP on tcgen05 / TMEM: 2-CTA MMAs (cta_group::2) + ping-pong.
//
// Same machinery as mlp_fwd.cu run backwards: a cluster of two CTAs owns 512
// points per iteration as two 256-row tile pairs; for every step (layer l, last
// to first) dH = dZ_l * W_l is one chain of M=256 MMAs over the half-split
// transposed weight stream, and the epilogue of one tile pair (ReLU sign mask
// from the forward's bit record, alpha-head term, bf16 pack, store to shared
// memory as the next A operand and to the dZ record for wgrad) runs under the
// MMAs of the other pair.  The heads (rgb / alpha / output_linear) are
// differentiated on the CUDA cores when a tile is (re)initialised.
//
// Reference semantics: autograd through NeRF2D.forward,
// /root/reference/src/run_nerf_helpers.py:106-135.
#include "mlp_common.cuh"

namespace ctx {

constexpr int kDgStages = 8;
constexpr int kDgPairs = kDgStages / 2;
constexpr int kDgStageBytes = 8192;          // 128 in-feature rows x 32 K x 2 B
constexpr int kDgHeadFloats = 1028;

struct DgradArgs {
  CtxMlpNet net;
  const uint8_t* wtstream;   // half-split transposed stream
  const float* fparams;
  const float* g_out;        // [P, out_ch]
  const uint8_t* acts;       // forward records (ReLU masks)
  uint8_t* dacts;            // dZ records
  int64_t P;
  int n_steps;
  int step_src[CTX_MLP_MAX_LAYERS];
  int step_dst[CTX_MLP_MAX_LAYERS];
};

struct __align__(8) DgSmemCtl {
  uint64_t full[kDgPairs], empty[kDgPairs];
  uint64_t acc_full[kTiles], act_ready[kTiles];
  uint32_t tmem_base;
};
// two 64 KB dZ tiles + weight ring + head weights (fp32) + barriers
constexpr size_t kDgSmemBytes = (size_t)kTiles * kHBytes + (size_t)kDgStages * kDgStageBytes + kDgHeadFloats * 4 + 256;

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kMlpThreads, 1)
mlp_dgrad_kernel(const __grid_constant__ DgradArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* h_buf = smem;
  uint8_t* w_buf = smem + kTiles * kHBytes;
  float* s_head = reinterpret_cast<float*>(w_buf + kDgStages * kDgStageBytes);
  DgSmemCtl* ctl = reinterpret_cast<DgSmemCtl*>(s_head + kDgHeadFloats);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t r = tc::cluster_ctarank();
  const CtxMlpNet& net = a.net;
  const int64_t n_citers = ceil_div(a.P, (int64_t)kTileM * 4);
  const int64_t cid = blockIdx.x >> 1, ncl = gridDim.x >> 1;
  const int n_steps = a.n_steps;

  if (tid == 0) {
    for (int s = 0; s < kDgPairs; ++s) {
      tc::mbar_init(&ctl->full[s], r == 0 ? 2 : 1); tc::mbar_init(&ctl->empty[s], 1);   // leader: + peer relay
    }
    for (int t = 0; t < kTiles; ++t) { tc::mbar_init(&ctl->acc_full[t], 1); tc::mbar_init(&ctl->act_ready[t], 16); }
    tc::mbar_fence_init();
  }
  if (warp == 1) tc::tmem_alloc2(&ctl->tmem_base, 512);
  {
    const int nhead = net.in_views > 0 ? 648 : 1028;
    for (int i = tid; i < nhead; i += kMlpThreads) s_head[i] = a.fparams[net.head_off + i];
  }
  tc::tc_fence_before();
  tc::cluster_sync_all();
  tc::tc_fence_after();
  const uint32_t tmem = ctl->tmem_base;
  // registers: from the control warpgroup to the two epilogue warpgroups (every warp of a warpgroup executes its call)
  // (the calls sit INSIDE the role branches: ptxas budgets each region by the setmaxnreg that dominates it)
  if (warp < 4) {
  reg_dealloc<kCtlRegs>();
  if (warp == 0) {
    // ============ transposed-weight producer (this CTA's half of every chunk, twice per step) ============
    uint32_t g = 0;
    for (int64_t it = cid; it < n_citers; it += ncl) {
      for (int si = 0; si < n_steps; ++si) {
        const int src = a.step_src[si];
        const int nchunks = net.L[src].N / CTX_MLP_KC;
        const uint8_t* lsrc = a.wtstream + net.L[src].wt_off + r * kDgStageBytes;
        for (int ph = 0; ph < 2; ++ph) {
          for (int c = 0; c < nchunks; ++c, ++g) {
            const int s = g % kDgStages, pr = s >> 1;
            if (!(g & 1)) tc::mbar_wait(&ctl->empty[pr], ((g / kDgStages) & 1) ^ 1);
            if (tc::elect_one()) {
              tc::mbar_expect_tx(&ctl->full[pr], kDgStageBytes);
              tc::bulk_g2s(w_buf + s * kDgStageBytes, lsrc + (size_t)c * 2 * kDgStageBytes, kDgStageBytes, &ctl->full[pr]);
              if (g & 1) tc::mbar_arrive(&ctl->full[pr]);
            }
            __syncwarp();
          }
        }
      }
    }
  } else if (warp == 1 || warp == 2) {
    if (r != 0) {
      if (warp == 1) {
      // ============ peer CTA: relay "my half landed" to the leader, once per chunk pair ============
      uint32_t g = 0;
      for (int64_t it = cid; it < n_citers; it += ncl)
        for (int si = 0; si < n_steps; ++si) {
          const int nchunks = net.L[a.step_src[si]].N / CTX_MLP_KC;
          for (int c = 0; c < 2 * nchunks; ++c, ++g) {
            if (g & 1) {
              const int pr = (g % kDgStages) >> 1;
              tc::mbar_wait(&ctl->full[pr], (g / kDgStages) & 1);
              if (tc::elect_one()) tc::mbar_arrive_remote(&ctl->full[pr], 0);
              __syncwarp();
            }
          }
        }
      }
    } else {
      // ============ leader CTA: two MMA issuers, warp 1 -> tile pair A, warp 2 -> tile pair B ============
      // (one thread cannot issue fast enough to keep the tensor pipe fed; see mlp_fwd.cu)
      const int ph = warp == 1 ? 0 : 1;
      if (tc::elect_one()) {
        // single issuing thread, descriptor lo words advanced by adds, one barrier wait + one commit per ring
        // pair (4 MMAs) -- see mlp_fwd.cu
        constexpr uint32_t kDescHi = (128u >> 4) | (1u << 14);
        constexpr uint32_t kALbo = (uint32_t)(kK8Stride >> 4) << 16, kBLbo = (2048u >> 4) << 16;
        constexpr uint32_t kAStep = 4 * kK8Stride >> 4;
        const uint32_t idesc = tc::make_idesc_bf16(256, 256, 0, 0);
        const uint32_t h_lo = kALbo | (tc::smem_u32(h_buf + ph * kHBytes) >> 4);
        const uint32_t w_lo = kBLbo | (tc::smem_u32(w_buf) >> 4);
        const uint32_t full0 = tc::smem_u32(&ctl->full[0]), empty0 = tc::smem_u32(&ctl->empty[0]);
        const uint32_t accf = tc::smem_u32(&ctl->acc_full[ph]), actr = tc::smem_u32(&ctl->act_ready[ph]);
        const uint32_t acc = tmem + ph * CTX_MLP_W;
        uint32_t gl = 0, act_phase = 0;
        for (int64_t it = cid; it < n_citers; it += ncl) {
          for (int si = 0; si < n_steps; ++si) {
            const int nchunks = net.L[a.step_src[si]].N / CTX_MLP_KC;   // 4 or 8: phases are whole chunk pairs
            uint32_t g = gl + ph * nchunks;
            gl += 2 * nchunks;
            tc::mbar_wait_addr(actr, act_phase);
            act_phase ^= 1;
            tc::tc_fence_after();
            for (int c = 0; c < nchunks; c += 2, g += 2) {
              const uint32_t s = g & (kDgStages - 1);
              // the previous use of this ring pair may be the other issuer's: it must have been consumed before a
              // parity wait on the full barrier means this use (see mlp_fwd.cu)
              if (c < kDgStages && g >= kDgStages)
                tc::mbar_wait_addr(empty0 + (s >> 1) * 8, ((g - kDgStages) / kDgStages) & 1);
              tc::mbar_wait_addr(full0 + (s >> 1) * 8, (g / kDgStages) & 1);   // both halves landed
#pragma unroll
              for (int j = 0; j < 2; ++j) {
                const uint32_t a_lo = h_lo + (c + j) * kAStep;
                const uint32_t b_lo = w_lo + (s + j) * (kDgStageBytes >> 4);
                tc::mma2_bf16_ss_w(acc, a_lo, kDescHi, b_lo, kDescHi, idesc, (c + j) > 0 ? 1u : 0u);
                tc::mma2_bf16_ss_w(acc, a_lo + (2 * kK8Stride >> 4), kDescHi, b_lo + (2 * 2048 >> 4), kDescHi, idesc,
                                   1u);
              }
              if (c + 2 >= nchunks) tc::mma2_commit_addr(accf);
              tc::mma2_commit_addr(empty0 + (s >> 1) * 8);
            }
          }
        }
      }
      __syncwarp();
    }
  }
  } else {
    reg_alloc<kEpiRegs>();
    // ============ head-init + epilogue warps ============
    const int q = warp & 3;                 // TMEM lane quarter
    const int hi = (warp - 4) >> 2;         // epilogue: column half ; head init: tile (0 = A, 1 = B)
    const int row = q * 32 + lane;
    const float* hw = s_head;
    const bool has_views = net.in_views > 0;
    const int out_ch = net.out_ch, act_tile_bytes = net.act_tile_bytes, gout_slot = net.gout_slot;
    const int last = net.n_layers - 1;
    const int lastN = net.L[last].N, last_mask = net.L[last].mask_slot, last_act = net.L[last].act_slot;
    const int last_rec = net.L[last].rec_ch, gout_ch0 = net.gout_ch0, gout_rec = net.gout_rec_ch;
    const int64_t nP = a.P;
    uint32_t acc_phase[2] = {0, 0};

    auto arrive_act = [&](int ph) {
      tc::fence_proxy_async_smem();
      tc::tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (r == 0) tc::mbar_arrive(&ctl->act_ready[ph]);
        else tc::mbar_arrive_remote(&ctl->act_ready[ph], 0);
      }
    };
    // dZ of the last GEMM layer from g_out (through the rgb / output head and that layer's ReLU mask)
    auto head_init = [&](int64_t it, int te) {
      const int64_t tile_idx = it * 4 + te * 2 + r;
      const int64_t p = tile_idx * kTileM + row;
      const bool valid = p < nP;
      const uint8_t* rec = a.acts + (size_t)tile_idx * act_tile_bytes;
      uint8_t* drec = a.dacts + (size_t)tile_idx * act_tile_bytes;
      uint8_t* my_h = h_buf + te * kHBytes;
      float g[4] = {0.f, 0.f, 0.f, 0.f};
      if (valid) {
        if (out_ch == 4) {
          const float4 g4 = *reinterpret_cast<const float4*>(a.g_out + p * 4);
          g[0] = g4.x; g[1] = g4.y; g[2] = g4.z; g[3] = g4.w;
        } else {
#pragma unroll
          for (int o = 0; o < 3; ++o) if (o < out_ch) g[o] = a.g_out[p * out_ch + o];
        }
      }
      {
        float gv[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) gv[i] = i < 4 ? g[i] : 0.f;
        store_row8(nullptr, row, gout_ch0, gv, false, drec + gout_slot, gout_rec);
        store_row8(nullptr, row, gout_ch0 + 8, gv + 8, false, drec + gout_slot, gout_rec);
      }
      const int nw = lastN / 32;
      const uint32_t* mrow = reinterpret_cast<const uint32_t*>(rec + last_mask) + row;   // [column block][row]
      for (int cb = 0; cb < nw; ++cb) {
        const uint32_t neg = __ldcs(mrow + cb * kTileM);
        float v[32];
        if (has_views) {
          const float* wr = hw + 260;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int c = cb * 32 + j;
            v[j] = g[0] * wr[c] + g[1] * wr[128 + c] + g[2] * wr[256 + c];
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int c = cb * 32 + j;
            v[j] = g[0] * hw[c] + g[1] * hw[256 + c] + g[2] * hw[512 + c] + g[3] * hw[768 + c];
          }
        }
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = ((neg >> (31 - j)) & 1u) ? 0.f : v[j];
#pragma unroll
        for (int j = 0; j < 32; j += 8) store_row8(my_h, row, cb * 32 + j, v + j, false, drec + last_act, last_rec);
      }
    };

    if (cid < n_citers) head_init(cid, hi);
    arrive_act(0);
    arrive_act(1);

    for (int64_t it = cid; it < n_citers; it += ncl) {
      for (int si = 0; si < n_steps; ++si) {
        const int dst = a.step_dst[si];
        const int Drelu = net.L[dst].relu, Dmask = net.L[dst].mask_slot, Dact = net.L[dst].act_slot;
        const bool add_alpha = net.L[dst].epi == CTX_EPI_HIDDEN_ALPHA;
        const bool has_next = si + 1 < n_steps;
#pragma unroll 1
        for (int ph = 0; ph < 2; ++ph) {
          const int64_t tile_idx = it * 4 + ph * 2 + r;
          const int64_t p = tile_idx * kTileM + row;
          const uint8_t* rec = a.acts + (size_t)tile_idx * act_tile_bytes;
          uint8_t* drec = a.dacts + (size_t)tile_idx * act_tile_bytes;
          uint8_t* my_h = h_buf + ph * kHBytes;
          const uint32_t my_acc = tmem + ph * CTX_MLP_W + ((uint32_t)(q * 32) << 16) + hi * 128;
          // this warp's 128 columns: 4 mask words, fetched while the MMAs run
          uint4 mw = make_uint4(0u, 0u, 0u, 0u);
          if (Drelu) {
            const uint32_t* mp = reinterpret_cast<const uint32_t*>(rec + Dmask) + hi * 4 * kTileM + row;
            mw = make_uint4(__ldcs(mp), __ldcs(mp + kTileM), __ldcs(mp + 2 * kTileM), __ldcs(mp + 3 * kTileM));
          }
          float d_alpha = 0.f;
          if (add_alpha && p < nP) d_alpha = __ldg(a.g_out + p * 4 + 3);
          const uint32_t par = ph ? acc_phase[1] : acc_phase[0];
          tc::mbar_wait(&ctl->acc_full[ph], par);
          if (ph) acc_phase[1] ^= 1; else acc_phase[0] ^= 1;
          tc::tc_fence_after();
          // straight-line instantiations (alpha-head term / last step)
          auto run = [&](auto alpha_c, auto next_c) {
            constexpr bool ALPHA = decltype(alpha_c)::value, NEXT = decltype(next_c)::value;
            auto process = [&](const uint32_t (&vr)[32], int cbi) {
              const uint32_t neg = cbi == 0 ? mw.x : cbi == 1 ? mw.y : cbi == 2 ? mw.z : mw.w;
              const int cb = hi * 4 + cbi;
              float v[32];
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                float x = __uint_as_float(vr[j]);
                if constexpr (ALPHA) x = fmaf(d_alpha, hw[cb * 32 + j], x);
                v[j] = ((neg >> (31 - j)) & 1u) ? 0.f : x;
              }
#pragma unroll
              for (int j = 0; j < 32; j += 8)
                store_row8(NEXT ? my_h : nullptr, row, cb * 32 + j, v + j, false, Dact >= 0 ? drec + Dact : nullptr, 256);
            };
            uint32_t va[32];
#pragma unroll
            for (int cbi = 0; cbi < 4; ++cbi) {   // serial load -> wait -> process: running loads ahead measured slower
              tc::tmem_ld32(my_acc + cbi * 32, va);
              tc::tmem_wait_ld(va);
              process(va, cbi);
            }
          };
          {
            using std::integral_constant;
            if (add_alpha) {
              if (has_next) run(integral_constant<bool, true>{}, integral_constant<bool, true>{});
              else run(integral_constant<bool, true>{}, integral_constant<bool, false>{});
            } else {
              if (has_next) run(integral_constant<bool, false>{}, integral_constant<bool, true>{});
              else run(integral_constant<bool, false>{}, integral_constant<bool, false>{});
            }
          }
          if (has_next) {
            arrive_act(ph);
          } else {
            // the tile is finished: start the next iteration's head for it as soon as possible
            const int64_t nit = it + ncl;
            if (nit < n_citers && hi == ph) head_init(nit, ph);
            if (nit < n_citers) arrive_act(ph);
          }
        }
      }
    }
  }

  tc::tc_fence_before();
  tc::cluster_sync_all();
  if (warp == 1) {
    __syncwarp();
    tc::tmem_dealloc2(tmem, 512);
  }
}

}  // namespace ctx

// host launcher used by ctx_mlp_bwd (mlp_wgrad.cu)
int ctx_launch_dgrad(const CtxMlpNet& net, const void* wtpacked, const float* fparams, const float* g_out,
                      const void* acts, void* dacts, int64_t P, int max_sms, cudaStream_t st) {
  ctx::DgradArgs a;
  a.net = net;
  a.wtstream = (const uint8_t*)wtpacked;
  a.fparams = fparams; a.g_out = g_out; a.acts = (const uint8_t*)acts; a.dacts = (uint8_t*)dacts; a.P = P;
  int n = 0;
  for (int l = net.n_layers - 1; l >= 1; --l) { a.step_src[n] = l; a.step_dst[n] = l - 1; ++n; }
  a.n_steps = n;
  static ctx::DeviceOnce attr_once;
  if (attr_once.needed()) {
    cudaError_t e = cudaFuncSetAttribute(ctx::mlp_dgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)ctx::kDgSmemBytes);
    if (e != cudaSuccess) return (int)e;
  }
  const int64_t citers = ctx::ceil_div(P, (int64_t)ctx::kTileM * 4);
  int cap = ctx::num_sms() / 2;   // one cluster per SM pair; a smaller SM budget leaves room for a concurrent kernel
  if (max_sms > 0 && max_sms / 2 < cap) cap = max_sms / 2;
  if (cap < 1) cap = 1;
  const int ncl = (int)(citers < cap ? citers : cap);
  ctx::mlp_dgrad_kernel<<<2 * ncl, ctx::kMlpThreads, ctx::kDgSmemBytes, st>>>(a);
  return (int)cudaGetLastError();
}
