// Host/device description of the coordinate MLP as a chain of tcgen05 GEMM
// layers (shared by the pack, forward, dgrad and wgrad kernels).
//
// Network (reference: NeRF2D, src/run_nerf_helpers.py:68-135, plus the upstream
// view-direction head kept there as comments :86-95, :117-127):
//   x_p = enc(points)  (in_pts ch, zero-padded to 64)      x_d = enc(dirs) (27 -> 32)
//   h = relu(W_l [x_p |] h + b_l)   l = 0..D-1, input skip-concatenated (x first, :115)
//   no views : out = W_o h + b_o                                     (:129)
//   views    : alpha = w_a.h + b_a ; f = W_f h + b_f ; v = relu(W_v [f | x_d] + b_v) ; rgb = W_r v + b_r
// Hidden width is fixed at W = 256 (the reference's value, trainer.py:133).
#pragma once
#include <stdint.h>

#define CTX_MLP_W 256
#define CTX_MLP_KC 32           /* K extent of one streamed weight chunk */
#define CTX_MLP_MAX_LAYERS 20
#define CTX_MLP_XP_PAD 64       /* padded channels of the point encoding */
#define CTX_MLP_XD_PAD 32       /* padded channels of the direction encoding */

enum {
  CTX_EPI_HIDDEN = 0,        /* h = act(acc + b) -> shared memory */
  CTX_EPI_HIDDEN_ALPHA = 1,  /* same, and alpha = w_a . h + b_a kept in a register */
  CTX_EPI_FINAL_VIEWS = 2,   /* v = relu(acc + b); rgb = W_r v + b_r; store [rgb, alpha] */
  CTX_EPI_FINAL_OUT = 3      /* h = relu(acc + b); out = W_o h + b_o; store out */
};

typedef struct {
  int32_t n_x_pre;   /* K32 chunks read from the x buffer before the h chunks */
  int32_t n_h;       /* K32 chunks read from the h buffer */
  int32_t n_x_post;  /* K32 chunks read from the x buffer after the h chunks */
  int32_t N;         /* output features of the GEMM: 256 or 128 */
  int32_t relu;
  int32_t epi;
  int32_t bias_off;  /* float offset into fparams */
  int32_t w_off;     /* byte offset of this layer's first chunk in the packed stream */
  int32_t wt_off;    /* byte offset in the transposed (dgrad) stream, -1 if none */
  int32_t act_slot;  /* byte offset, within a tile's activation record, of this layer's OUTPUT; -1: the layer keeps no
                        record (feature_linear: its weight gradients are rebuilt algebraically, see mlp_wgrad.cu) */
  int32_t in_slot;   /* byte offset of the record holding this layer's h INPUT (-1: none) */
  int32_t mask_slot; /* byte offset of the ReLU sign-bit record (128 rows x N/32 words), -1: none */
  int32_t bias_mma;  /* 2-CTA kernels fold the bias into the GEMM through the constant-1 pad channel of the x
                        buffer: 0 = it rides in an x chunk this layer reads anyway, 1 = one extra K=16 MMA whose
                        B operand is the 16-wide "bias chunk" stored after the layer's regular chunks */
  int32_t bias_a_off;/* byte offset inside the x tile of the 16 channels ending in the constant-1 channel */
  int32_t rec_ch;    /* channel width of the record slot (>= N: the views layer's slot also holds the 16 g_out channels,
                        so that [dZ_views | g_out] is ONE 144-wide wgrad operand) */
} CtxMlpLayer;

typedef struct {
  int32_t n_layers;
  int32_t in_pts;        /* real point-encoding channels (63 or 42) */
  int32_t in_views;      /* 27 or 0 */
  int32_t out_ch;        /* 4 or 3 */
  int32_t head_off;      /* float offset into fparams of the head block */
  int32_t w_bytes;       /* size of the packed forward stream */
  int32_t wt_bytes;      /* size of the packed transposed stream */
  int32_t n_fparams;     /* floats in fparams */
  int32_t act_tile_bytes;/* bytes of one 128-point activation record */
  int32_t xp_slot, xd_slot, gout_slot;
  int32_t gout_ch0;      /* first channel of the 16 g_out channels inside the slot at gout_slot */
  int32_t gout_rec_ch;   /* channel width of that slot (16, or 144 when it is the views layer's) */
  CtxMlpLayer L[CTX_MLP_MAX_LAYERS];
} CtxMlpNet;

/* fparams head block, floats from head_off:
 *   views   : w_alpha[256] | b_alpha,0,0,0 | W_rgb[3][128] | b_rgb[3],0
 *   no views: W_out[4][256] (rows >= out_ch zero) | b_out[4]                */
