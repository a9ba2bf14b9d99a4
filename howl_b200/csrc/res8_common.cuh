// Declarations shared by the fp32 (res8.cu) and tensor-core (res8_tc.cu) Res8 kernels.
#pragma once
#include <cuda_bf16.h>

#include "common.cuh"

#define R8_C 45
#define R8_W 10          // pooled width (n_mels 40 / 4)
#define R8_LAYERS 6
#define R8_MELS 40
#define R8_WPAD 12       // padded row (1 + 10 + 1)
#define R8_KW (R8_C * R8_C * 9)   // 18225 weights per 45->45 layer
#define R8_BN_EPS 1e-5
#define R8_BN_MOM 0.1


struct ConvParams {
  const float* in;        // [B,45,H,10]
  const float* in_mean;   // [45] or null (identity)
  const float* in_rstd;
  const float* w;         // [45 out][45 in][3][3]
  const float* res;       // residual added after ReLU, or null
  float* out;
  double* stats;          // [2][45] or null
  const float* aux;       // STATS == 2: tensor whose normalised value multiplies the output in the 2nd statistic
  const float* aux_mean;
  const float* aux_rstd;
  int64_t B;
  int H;
};

struct WgradParams {
  const float* dc;       // [B,45,H,10]  gradient at the conv output (ReLU mask applied); fp32 engine
  const __nv_bfloat16* dc_op;   // tensor-core engine: the same gradient in operand format (see r8tc_dcop_bytes)
  const float* x;        // [B,45,H,10]  conv input before normalisation
  const float* x_mean;   // or null
  const float* x_rstd;
  float* dw;             // [45][45][3][3], accumulated with atomics
  int64_t B;
  int H;
};

// tensor-core weight operands: per layer and direction (0 = forward with the producer's BatchNorm folded in, 1 = data
// gradient) one bf16 block [9 taps][6 chunks][96 n: 48 hi | 48 lo][8 k]
#define R8TC_WBLOCK (54 * 48 * 8)
#define R8TC_WELEMS ((size_t)R8_LAYERS * 2 * 2 * R8TC_WBLOCK)

bool r8tc_supported(int H);

// Operand format of every tensor the tensor-core convolutions read (activations "u_op", conv-output gradients "dc_op"),
// landed by TMA straight into the UMMA operand tiles:
//   per utterance [part: hi, lo][chunk of 8 channels: 6][raster row q: R][8 x bf16],  q = (y + 1) * 11 + (x + 1),
//   R = round_up((H + 2) * 11, 64); rows outside the image are zero.  Channel 45 of u_op is 1 at the image pixels.
size_t r8tc_dcop_bytes(int H);          // bytes per utterance
int r8tc_dcop_rows(int H);              // R

struct ApplyOpParams {
  const float* g;          // dgrad output [B,45,HW], or null when g_bcast is used
  const float* g_bcast;    // [B,45]: g = g_bcast / HW (layer 6)
  const float* u;
  const float* mean_rstd;  // [2][45] of this layer
  const double* stats;     // [2][45] sum(g), sum(g xhat)
  const float* gu_in;
  const float* mask_prev;
  float* gu_out;
  __nv_bfloat16* dc_op;
  __nv_bfloat16* dc_opT;   // the same gradient with rows = channels (weight-gradient A operand), r8tc_dcop_bytes per utterance
  int64_t B;
  int H;
  double count;
};
int r8tc_apply(howl_ctx_t* ctx, cudaStream_t st, const ApplyOpParams& p);
// data-gradient weight operands of all six layers (direction 1 blocks of wprep)
int r8tc_weight_prep(howl_ctx_t* ctx, cudaStream_t st, const float* w_layers, __nv_bfloat16* wprep);
// forward weight operand of one layer with BatchNorm(mean_rstd, or identity when null) folded in (the border-dependent bias
// rides on the ones channel)
int r8tc_fold(howl_ctx_t* ctx, cudaStream_t st, const float* w_layer, const float* mean_rstd, __nv_bfloat16* blk);
// forward (fwd, stats 0 / 1) or data gradient (!fwd, stats 0 / 2) of one layer; p.in is unused
int r8tc_conv(howl_ctx_t* ctx, cudaStream_t st, const ConvParams& p, const __nv_bfloat16* in_op, __nv_bfloat16* out_op,
              const __nv_bfloat16* w, bool fwd, int stats);
int r8tc_wgrad(howl_ctx_t* ctx, cudaStream_t st, const __nv_bfloat16* dc_opT, const __nv_bfloat16* x_op, const float* x_mean,
               const float* x_rstd, float* dw, int64_t B, int H);
