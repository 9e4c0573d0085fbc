// Network description (host) and weight packing (device) for the tcgen05 MLP.
//
// ctx_mlp_describe builds the layer chain for NeRF2D(D, W=256, input_ch, output_ch,
// skips) -- /root/reference/src/run_nerf_helpers.py:70-104 -- or, with
// in_views > 0, the upstream view-direction variant (commented :86-95).
// ctx_mlp_pack converts the fp32 nn.Linear weights ([out,in] row-major, parameter
// order of :81-97) into (a) the bf16 forward stream: per layer, per 32-wide K
// chunk, the B operand image [half][k8][N/2][8] (canonical no-swizzle K-major
// layout; each CTA of an MMA pair stages one half), with the bias riding on the
// constant-1 input channel, (b) the transposed stream used by the dgrad kernel,
// (c) the fp32 block of biases and head weights.  Runs once per optimizer step.
#include "ctx_common.cuh"
#include "mlp_desc.h"
#include <string.h>

namespace ctx {

struct PackSeg { int src, len, pad; };   // source column start, real length, padded length
struct PackLayer {
  const float* W; const float* b;
  int out_rows, ld;        // real rows, row stride of W
  int N;                   // padded rows of the B image
  int nseg; PackSeg seg[3];
  int w_off, wt_off, bias_off;
  int h_src;               // source column of the h part (transposed stream), -1: none
  int bias_mma;            // split copy: extra 16-wide bias chunk after the regular chunks
  int Kt;                  // K extent (= out features padded to 32) of the transposed stream
};
struct PackArgs {
  int n_layers;
  PackLayer L[CTX_MLP_MAX_LAYERS];
  uint16_t* w; uint16_t* wt; float* fparams;
  // heads
  int has_views, out_ch, head_off;
  const float* w_alpha; const float* b_alpha; const float* w_rgb; const float* b_rgb;
  const float* w_out; const float* b_out;
};

__device__ __forceinline__ uint16_t f2bf(float v) { return __bfloat16_as_ushort(__float2bfloat16_rn(v)); }

__global__ void mlp_pack_kernel(const __grid_constant__ PackArgs a) {
  const int l = blockIdx.y;
  const int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
  if (l == a.n_layers) {  // head block
    float* hp = a.fparams + a.head_off;
    if (a.has_views) {
      for (int i = tid; i < 256; i += nth) hp[i] = a.w_alpha[i];
      for (int i = tid; i < 4; i += nth) hp[256 + i] = i == 0 ? a.b_alpha[0] : 0.f;
      for (int i = tid; i < 384; i += nth) hp[260 + i] = a.w_rgb[i];
      for (int i = tid; i < 4; i += nth) hp[644 + i] = i < 3 ? a.b_rgb[i] : 0.f;
    } else {
      for (int i = tid; i < 1024; i += nth) hp[i] = (i / 256 < a.out_ch) ? a.w_out[i] : 0.f;
      for (int i = tid; i < 4; i += nth) hp[1024 + i] = i < a.out_ch ? a.b_out[i] : 0.f;
    }
    return;
  }
  const PackLayer& L = a.L[l];
  int Kpad = 0;
  for (int s = 0; s < L.nseg; ++s) Kpad += L.seg[s].pad;
  // forward stream: element e -> (chunk c, k8, n, kk)
  uint16_t* dst = a.w + L.w_off / 2;
  const int total = Kpad * L.N;
  for (int e = tid; e < total; e += nth) {
    const int c = e / (L.N * 32), r0 = e - c * (L.N * 32);
    const int k8 = r0 / (L.N * 8), r1 = r0 - k8 * (L.N * 8);
    const int n = r1 >> 3, kk = r1 & 7;
    int k = c * 32 + k8 * 8 + kk, col = -1;
    bool one_channel = false;   // last padded channel of an x segment: constant 1 in the 2-CTA kernels
    for (int s = 0; s < L.nseg; ++s) {
      if (k < L.seg[s].pad) {
        if (k < L.seg[s].len) col = L.seg[s].src + k;
        else if (k == L.seg[s].pad - 1) one_channel = true;
        break;
      }
      k -= L.seg[s].pad;
    }
    uint16_t val = (col >= 0 && n < L.out_rows) ? f2bf(L.W[(size_t)n * L.ld + col]) : (uint16_t)0;
    if (one_channel && !L.bias_mma && n < L.out_rows) val = f2bf(L.b[n]);   // bias rides on the constant-1 channel
    // half-split chunk for the 2-CTA MMAs: [half][k8][N/2][8] (each CTA of the pair stages one half)
    const int Nh = L.N >> 1, hf = n / Nh, nl = n - hf * Nh;
    dst[c * (L.N * 32) + hf * (Nh * 32) + k8 * (Nh * 8) + nl * 8 + kk] = val;
  }
  for (int i = tid; i < L.N; i += nth) a.fparams[L.bias_off + i] = i < L.out_rows ? L.b[i] : 0.f;
  if (L.bias_mma) {   // [half][2 k8][N/2][8]: bias at k = 15 (the constant-1 channel), zeros elsewhere
    uint16_t* bc = a.w + L.w_off / 2 + Kpad * L.N;
    const int Nh = L.N >> 1;
    for (int e = tid; e < 16 * L.N; e += nth) {
      const int hf = e / (16 * Nh), r0 = e - hf * 16 * Nh;
      const int k8 = r0 / (Nh * 8), r1 = r0 - k8 * Nh * 8;
      const int nl = r1 >> 3, kk = r1 & 7, n = hf * Nh + nl;
      bc[e] = (k8 == 1 && kk == 7 && n < L.out_rows) ? f2bf(L.b[n]) : (uint16_t)0;
    }
  }
  // transposed stream (dgrad): B_T[n' = h input feature (256)][k = output feature], chunks over k
  if (L.wt_off >= 0 && a.wt != nullptr) {
    uint16_t* dt = a.wt + L.wt_off / 2;
    const int totalT = L.Kt * 256;
    for (int e = tid; e < totalT; e += nth) {
      const int c = e / (256 * 32), r0 = e - c * (256 * 32);
      const int k8 = r0 / (256 * 8), r1 = r0 - k8 * (256 * 8);
      const int n = r1 >> 3, kk = r1 & 7;
      const int k = c * 32 + k8 * 8 + kk;  // output feature
      const uint16_t val = (k < L.out_rows) ? f2bf(L.W[(size_t)k * L.ld + L.h_src + n]) : (uint16_t)0;
      const int hf = n >> 7, nl = n & 127;
      dt[c * (256 * 32) + hf * (128 * 32) + k8 * (128 * 8) + nl * 8 + kk] = val;
    }
  }
}

}  // namespace ctx

#include <cuda.h>
namespace ctx {
int make_record_tensor_map(CUtensorMap* out, void* base, int tile_bytes, int64_t n_tiles) {
  typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static EncodeFn encode = nullptr;
  if (!encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn ||
        qres != cudaDriverEntryPointSuccess)
      return CTX_ERR_UNSUPPORTED;
    encode = (EncodeFn)fn;
  }
  if (tile_bytes % 1024 != 0 || n_tiles < 1 || (reinterpret_cast<uintptr_t>(base) & 15) != 0) return CTX_ERR_BAD_ARG;
  const cuuint64_t gdim[4] = {128, 2, (cuuint64_t)(tile_bytes / 1024), (cuuint64_t)n_tiles};
  const cuuint64_t gstride[3] = {32768, 1024, (cuuint64_t)tile_bytes};      // bytes, dims 1..3
  const cuuint32_t box[4] = {128, 2, 32, 1};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  const CUresult rc = encode(out, CU_TENSOR_MAP_DATA_TYPE_UINT64, 4, base, gdim, gstride, box, estr,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return rc == CUDA_SUCCESS ? 0 : CTX_ERR_UNSUPPORTED;
}
}  // namespace ctx

extern "C" int ctx_mlp_net_bytes(void) { return (int)sizeof(CtxMlpNet); }

extern "C" int ctx_mlp_describe(int D, uint32_t skip_mask, int in_pts, int in_views, int out_ch, void* net_out) {
  // (the last channel of each padded encoding is the constant 1 that carries the bias through the GEMM)
  if (!net_out || D < 2 || D > 16 || in_pts < 1 || in_pts > CTX_MLP_XP_PAD - 1 || in_views < 0 ||
      in_views > CTX_MLP_XD_PAD - 1 || out_ch < 1 || out_ch > 4)
    return CTX_ERR_BAD_ARG;
  if (in_views > 0 && out_ch != 4) return CTX_ERR_BAD_ARG;
  CtxMlpNet net;
  memset(&net, 0, sizeof(net));
  net.in_pts = in_pts; net.in_views = in_views; net.out_ch = out_ch;
  const int xp_chunks = CTX_MLP_XP_PAD / CTX_MLP_KC, h_chunks = CTX_MLP_W / CTX_MLP_KC;
  int w_off = 0, wt_off = 0, f_off = 0, slot = 0;
  net.xp_slot = slot; slot += 128 * CTX_MLP_XP_PAD * 2;
  net.xd_slot = slot; slot += 128 * CTX_MLP_XD_PAD * 2;
  if (in_views == 0) { net.gout_slot = slot; slot += 128 * 16 * 2; net.gout_ch0 = 0; net.gout_rec_ch = 16; }
  int n = 0, prev_slot = -1;
  for (int l = 0; l < D; ++l, ++n) {
    CtxMlpLayer& L = net.L[n];
    const bool skip_in = l > 0 && ((skip_mask >> (l - 1)) & 1u);
    L.n_x_pre = (l == 0 || skip_in) ? xp_chunks : 0;
    L.n_h = l == 0 ? 0 : h_chunks;
    L.n_x_post = 0; L.N = CTX_MLP_W; L.relu = 1; L.epi = CTX_EPI_HIDDEN;
    L.bias_off = f_off; f_off += L.N;
    L.w_off = w_off; w_off += (L.n_x_pre + L.n_h) * CTX_MLP_KC * L.N * 2;
    L.bias_mma = L.n_x_pre ? 0 : 1; L.bias_a_off = (CTX_MLP_XP_PAD - 16) / 8 * 128 * 16;
    if (L.bias_mma) w_off += 16 * L.N * 2;
    if (l > 0) { L.wt_off = wt_off; wt_off += CTX_MLP_W * 256 * 2; } else L.wt_off = -1;
    L.act_slot = slot; slot += 128 * L.N * 2; L.rec_ch = L.N;
    L.mask_slot = slot; slot += 128 * (L.N / 32) * 4;
    L.in_slot = prev_slot; prev_slot = L.act_slot;
  }
  if (in_views > 0) {
    net.L[n - 1].epi = CTX_EPI_HIDDEN_ALPHA;
    CtxMlpLayer& F = net.L[n];  // feature_linear: 256 -> 256, no activation
    F.n_x_pre = 0; F.n_h = h_chunks; F.n_x_post = 0; F.N = CTX_MLP_W; F.relu = 0; F.epi = CTX_EPI_HIDDEN;
    F.bias_off = f_off; f_off += F.N;
    F.w_off = w_off; w_off += F.n_h * CTX_MLP_KC * F.N * 2;
    F.bias_mma = 1; F.bias_a_off = (CTX_MLP_XD_PAD - 16) / 8 * 128 * 16; w_off += 16 * F.N * 2;
    F.wt_off = wt_off; wt_off += CTX_MLP_W * 256 * 2;
    // no record of the feature layer: it is linear, so dW_feature, dW_views[:, :256] and their biases follow from
    // G = dZ_views^T h_{D-1} (one wgrad job) and the weights themselves (mlp_wgrad.cu, wgrad_post_kernel)
    F.act_slot = -1; F.rec_ch = F.N;
    F.mask_slot = -1;
    F.in_slot = prev_slot; prev_slot = -1;
    ++n;
    CtxMlpLayer& V = net.L[n];  // views_linears.0: [feature | dirs] -> 128, relu ; rgb head folded in
    V.n_x_pre = 0; V.n_h = h_chunks; V.n_x_post = CTX_MLP_XD_PAD / CTX_MLP_KC; V.N = CTX_MLP_W / 2; V.relu = 1;
    V.epi = CTX_EPI_FINAL_VIEWS;
    V.bias_off = f_off; f_off += V.N;
    V.w_off = w_off; w_off += (V.n_h + V.n_x_post) * CTX_MLP_KC * V.N * 2;
    V.bias_mma = 0; V.bias_a_off = 0;
    V.wt_off = wt_off; wt_off += (CTX_MLP_W / 2) * 256 * 2;
    V.rec_ch = V.N + 16;   // [v or dZ_views (128) | g_out (16)]
    V.act_slot = slot; slot += 128 * V.rec_ch * 2;
    net.gout_slot = V.act_slot; net.gout_ch0 = V.N; net.gout_rec_ch = V.rec_ch;
    V.mask_slot = slot; slot += 128 * (V.N / 32) * 4;
    V.in_slot = prev_slot;
    ++n;
    net.head_off = f_off; f_off += 648;
  } else {
    net.L[n - 1].epi = CTX_EPI_FINAL_OUT;
    net.head_off = f_off; f_off += 1028;
  }
  net.n_layers = n;
  net.w_bytes = w_off; net.wt_bytes = wt_off; net.n_fparams = f_off; net.act_tile_bytes = slot;
  memcpy(net_out, &net, sizeof(net));
  return 0;
}

// params: host array of device pointers in module order:
//   pts_linears.{0..D-1}.{weight,bias}, then
//   no views: output_linear.{weight,bias}
//   views   : feature_linear.{w,b}, alpha_linear.{w,b}, views_linears.0.{w,b}, rgb_linear.{w,b}
extern "C" int ctx_mlp_pack(const void* net_host, const float* const* params, int n_params, void* wpacked,
                            void* wtpacked, float* fparams, void* stream) {
  if (!net_host || !params || !wpacked || !fparams) return CTX_ERR_BAD_ARG;
  const CtxMlpNet& net = *reinterpret_cast<const CtxMlpNet*>(net_host);
  const bool views = net.in_views > 0;
  const int D = views ? net.n_layers - 2 : net.n_layers;
  if (n_params != 2 * D + (views ? 8 : 2)) return CTX_ERR_BAD_ARG;
  for (int i = 0; i < n_params; ++i) if (!params[i]) return CTX_ERR_BAD_ARG;
  ctx::PackArgs a;
  memset(&a, 0, sizeof(a));
  a.n_layers = net.n_layers; a.w = (uint16_t*)wpacked; a.wt = (uint16_t*)wtpacked; a.fparams = fparams;
  a.has_views = views; a.out_ch = net.out_ch; a.head_off = net.head_off;
  for (int l = 0; l < net.n_layers; ++l) {
    const CtxMlpLayer& L = net.L[l];
    ctx::PackLayer& P = a.L[l];
    P.N = L.N; P.w_off = L.w_off; P.wt_off = L.wt_off; P.bias_off = L.bias_off; P.bias_mma = L.bias_mma;
    int pi;
    if (l < D) pi = 2 * l;
    else if (l == D) pi = 2 * D;          // feature_linear
    else pi = 2 * D + 4;                  // views_linears.0
    P.W = params[pi]; P.b = params[pi + 1];
    P.nseg = 0;
    int col = 0;
    if (L.n_x_pre) { P.seg[P.nseg++] = {col, net.in_pts, CTX_MLP_XP_PAD}; col += net.in_pts; }
    P.h_src = -1;
    if (L.n_h) { P.h_src = col; P.seg[P.nseg++] = {col, CTX_MLP_W, CTX_MLP_W}; col += CTX_MLP_W; }
    if (L.n_x_post) { P.seg[P.nseg++] = {col, net.in_views, CTX_MLP_XD_PAD}; col += net.in_views; }
    P.ld = col;
    P.out_rows = L.N;
    P.Kt = L.N;
  }
  if (views) {
    a.w_alpha = params[2 * D + 2]; a.b_alpha = params[2 * D + 3];
    a.w_rgb = params[2 * D + 6]; a.b_rgb = params[2 * D + 7];
  } else {
    a.w_out = params[2 * D]; a.b_out = params[2 * D + 1];
  }
  dim3 grid(32, net.n_layers + 1);
  ctx::mlp_pack_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(a);
  CTX_RETURN_LAST();
}
