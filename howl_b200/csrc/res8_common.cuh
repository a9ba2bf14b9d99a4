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

// head of the backward (layer 6: upstream gradient = broadcast dh / HW), tensor-core engine
struct ApplyOpParams {
  const float* g_bcast;    // [B,45]: g = g_bcast / HW
  const __nv_bfloat16* u_op;   // u_6 in operand format
  const uint16_t* mask_bits;   // ReLU decisions of conv_6 stored by the forward: [B][3 channel groups][R]
  const float* mean_rstd;  // [2][45] of layer 6
  const double* stats;     // [2][45] sum(g), sum(g xhat)
  float* gu_out;           // planar G_6 (residual-path gradient of layer 4)
  __nv_bfloat16* dc_op;    // conv-output gradient of layer 6 in operand format
  int64_t B;
  int H;
  double count;
};
int r8tc_apply_head(howl_ctx_t* ctx, cudaStream_t st, const ApplyOpParams& p);

// one launch of the stream kernel (res8_tc.cu).  mode 0: forward (eval), 1: forward + BatchNorm statistics, 2: plain data gradient
// (planar fp32 out), 3: data gradient + BatchNorm backward / ReLU mask of the producer layer fused into the epilogue.
struct TcConvCall {
  int mode;
  int64_t B;
  int H;
  const __nv_bfloat16* in_op;    // A operand: activations (forward) or conv-output gradient (data gradient), operand format
  const __nv_bfloat16* w;        // weight operand block of this layer and direction
  // forward
  __nv_bfloat16* out_op;         // output in operand format (or null)
  const __nv_bfloat16* res_op;   // residual added after the ReLU (or null)
  uint16_t* mask_out;            // ReLU decisions [B][3][R] (or null)
  float* pooled_raw;             // [B][45] spatial sums of the output, accumulated (layer 6; or null)
  double* stats;                 // mode 1: [2][45] sum, sum of squares
  float* out_planar;             // planar fp32 copy of the output (mode 2: the data gradient; forward: optional)
  // mode 3
  const __nv_bfloat16* u_op;
  const float* bn_coef;          // [3][48] from r8tc_bn_bwd_coef
  const float* gu_in;
  float* gu_out;
  const uint16_t* mask_in;
  __nv_bfloat16* dc_out;
};
int r8tc_conv(howl_ctx_t* ctx, cudaStream_t st, const TcConvCall& c);
// BatchNorm-backward coefficients of layer j from the weight gradient of layer j + 1 (see bn_bwd_coef_kernel)
int r8tc_bn_bwd_coef(howl_ctx_t* ctx, cudaStream_t st, const float* w, const float* dw, const float* dones, const float* mean_rstd,
                     double count, float* coef);
int r8tc_debug_mask(howl_ctx_t* ctx, cudaStream_t st, const __nv_bfloat16* u_op, const uint16_t* bits, uint8_t* mask, int64_t B, int H);
// conv0 + ReLU + AvgPool(3,4) on the tensor cores: features -> a0 in operand format (+ the ReLU bits of every pooling window, or null)
int r8tc_conv0(howl_ctx_t* ctx, cudaStream_t st, const float* feats, const float* w0, __nv_bfloat16* a0_op, uint16_t* bits0, int64_t B, int F,
               int H);
// data-gradient weight operands of all six layers (direction 1 blocks of wprep)
int r8tc_weight_prep(howl_ctx_t* ctx, cudaStream_t st, const float* w_layers, __nv_bfloat16* wprep);
// forward weight operand of one layer with BatchNorm(mean_rstd, or identity when null) folded in (the border-dependent bias
// rides on the ones channel)
// BatchNorm finalisation of the producer layer done inside the fold (train mode): batch statistics -> mean / rstd (also written to
// mean_rstd_out for the backward), running statistics and num_batches_tracked updated once.  stats == null: mean_rstd is read as is.
struct TcFoldBn {
  const double* stats;      // [2][45] sum, sum of squares of the producer's output
  double count;
  float* mean_rstd_out;     // [2][45]
  float* running;           // [2][45] running mean / var, or null
  int64_t* nbt;             // or null
};
int r8tc_fold(howl_ctx_t* ctx, cudaStream_t st, const float* w_layer, const float* mean_rstd, __nv_bfloat16* blk, const TcFoldBn* bn = nullptr);
int r8tc_wgrad(howl_ctx_t* ctx, cudaStream_t st, const __nv_bfloat16* dc_op, const __nv_bfloat16* x_op, const float* x_mean,
               const float* x_rstd, float* dw, float* dones, int64_t B, int H);
