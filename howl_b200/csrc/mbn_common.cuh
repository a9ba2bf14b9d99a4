// MobileNetV2 path (howl/model/cnn.py:15-29 on torchvision's MobileNetV2) -- shared declarations.
//
// Every activation / gradient tensor of this path lives in HBM in ONE layout, the "tile-major operand format" (TMO):
//     [T = ceil(M / 128) row tiles][C8 = Cp / 8 channel chunks][128 rows][8 channels]  bf16,     Cp = channels padded to 16
// rows = NHWC pixels (row = (b * H + y) * W + x).  A 128-row tile with all its channel chunks is contiguous and IS the no-swizzle
// K-major UMMA operand (8-row x 16-byte core matrices, chunk stride 2048 B) of the pointwise-convolution GEMMs, and at the same
// time the MN-major operand (K = rows) of their weight gradients -- so one TMA bulk copy lands a tile and tcgen05.mma reads it with
// no staging pass.  Elementwise / depthwise kernels address (row, 8-channel chunk) as one 16-byte vector; consecutive threads take
// consecutive rows, so their loads and stores are contiguous 512-byte segments.  Pad rows and pad channels are kept at zero.
#pragma once
#include <cuda_bf16.h>

#include "common.cuh"

#define MBN_TILE 128

__host__ __device__ static inline int mbn_pad16(int c) { return (c + 15) & ~15; }
__host__ __device__ static inline int64_t mbn_tiles(int64_t rows) { return (rows + MBN_TILE - 1) / MBN_TILE; }
// bytes of a TMO tensor of `rows` rows and `c` channels
__host__ __device__ static inline size_t mbn_tmo_bytes(int64_t rows, int c) { return (size_t)mbn_tiles(rows) * MBN_TILE * mbn_pad16(c) * 2; }
// index (in 16-byte vectors) of (row, chunk) in a TMO tensor with c8 chunks
__host__ __device__ static inline size_t mbn_vec(int64_t row, int chunk, int c8) {
  return ((size_t)(row >> 7) * c8 + chunk) * MBN_TILE + (size_t)(row & 127);
}

// ---- pointwise-convolution GEMMs (mbn_gemm.cu) -------------------------------------------------------------------------------
// N tile of the GEMMs for a (padded) channel count: the largest multiple of 16 that divides np and is <= 240
int mbn_ntile(int np);
// Weight operand of the NT GEMM from an fp32 [n][k] matrix (row stride ld, rows >= n / cols >= k read as zero):
//   out = [np / nt tiles][kp / 8][nt rows][8] bf16;  transpose != 0 takes the matrix as [k][n] (the data-gradient operand W^T)
int mbn_weight_operand(howl_ctx_t* ctx, cudaStream_t st, const float* w, int n, int k, int ld, int transpose, __nv_bfloat16* out);
size_t mbn_weight_operand_bytes(int n, int k);
// several operands in one launch (kernel-argument table)
#define MBN_WOP_BATCH 40
struct MbnWopDesc {
  const float* w;
  __nv_bfloat16* out;
  int n, k, ld, transpose, nt;   // nt is filled in by mbn_weight_operand_batch
};
struct MbnWopBatch {
  MbnWopDesc d[MBN_WOP_BATCH];
};
int mbn_weight_operand_batch(howl_ctx_t* ctx, cudaStream_t st, const MbnWopDesc* descs, int count);
// C[M x N] = A[M x K] * W[N x K]^T (+ Add):  A, C, Add in TMO (rows >= M are written as zero), fp32 accumulation in TMEM
int mbn_gemm_nt(howl_ctx_t* ctx, cudaStream_t st, const __nv_bfloat16* A, const __nv_bfloat16* wop, const __nv_bfloat16* add,
                __nv_bfloat16* C, int64_t M, int K, int N);
// dW[n][k] (fp32, row stride ld, only n < n_valid / k < k_valid written) += sum_rows dC[row][n] * A[row][k]
int mbn_gemm_wgrad(howl_ctx_t* ctx, cudaStream_t st, const __nv_bfloat16* dC, const __nv_bfloat16* A, float* dW, int64_t M, int N,
                   int K, int n_valid, int k_valid, int ld);

// ---- fp32 A^T B on the tensor cores (bf16 x 3 split), used by the LSTM / LAS weight gradients -----------------------------------
// hi = bf16(x), lo = bf16(x - hi) of the fp32 row-major matrix x [rows][ld] (columns 0 .. c-1), both in TMO (pad rows / channels zero);
// colsum / colsum2 != null: += the column sums of x (the bias gradients come for free with the pass)
int mbn_pack_split(howl_ctx_t* ctx, cudaStream_t st, const float* x, int64_t ld, int64_t rows, int c, __nv_bfloat16* hi, __nv_bfloat16* lo,
                   float* colsum = nullptr, float* colsum2 = nullptr);
// dW[n][k] (fp32, row stride ld) += sum_r X[r][n] Y[r][k] = Xhi^T Yhi + Xhi^T Ylo + Xlo^T Yhi  (|error| ~ 2^-16 per product, fp32 accumulate)
int mbn_atb3_packed(howl_ctx_t* ctx, cudaStream_t st, const __nv_bfloat16* xhi, const __nv_bfloat16* xlo, const __nv_bfloat16* yhi,
                    const __nv_bfloat16* ylo, float* dW, int64_t rows, int N, int K, int ld);

// ---- fp32 C = X W^T on the tensor cores (bf16 x 3 split folded into K), used by the LSTM heads --------------------------------
// X3 = [X_hi | X_hi | X_lo] in TMO with 3 * pad16(c) channels (mbn_tmo_bytes(rows, 3 * pad16(c)) bytes)
int mbn_pack3(howl_ctx_t* ctx, cudaStream_t st, const float* x, int64_t ld, int64_t rows, int c, __nv_bfloat16* out);
// W3 = [W_hi | W_lo | W_hi] operand of an fp32 [n][k] matrix (row stride ld; transpose != 0: stored [k][n])
size_t mbn_weight_operand3_bytes(int n, int k);
int mbn_weight_operand3(howl_ctx_t* ctx, cudaStream_t st, const float* w, int n, int k, int ld, int transpose, __nv_bfloat16* out);
// C[row * ldc + n] (+)= sum_k X[row][k] W[n][k] (+ bias[n]) (ReLU), fp32 row-major output
int mbn_gemm_nt3_f32(howl_ctx_t* ctx, cudaStream_t st, const __nv_bfloat16* x3, const __nv_bfloat16* wop3, float* C, int64_t ldc, int64_t M, int K,
                     int N, const float* bias, int relu, int accumulate = 0);
