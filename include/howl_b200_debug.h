/*
 * howl_b200_debug.h -- tuning aids and test hooks of libhowl_b200.so.  NOT part of the drop-in boundary
 * (include/howl_b200.h): nothing a howl maintainer binds lives here.  Used by tools/ (micro-benchmarks) and by
 * tests/ (the mask-forced gradient oracle).
 */
#ifndef HOWL_B200_DEBUG_H_
#define HOWL_B200_DEBUG_H_

#include "howl_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* One 128x48x32 GEMM through the library's UMMA descriptor helpers (A, B fp32 device arrays, bf16-rounded inside;
 * mn_major selects the operand layout of the weight-gradient GEMM); D fp32 [128][48]. */
int howl_b200_selftest_umma(howl_ctx_t* ctx, void* stream, const float* A, const float* B, float* D, int32_t mn_major,
                            int32_t variant);

/* The tensor-core forward (kind 1) or data-gradient (kind 2) kernels write per-CTA cycle counters of their pipeline
 * waits to buf[sm_count][16] (uint64, device memory); buf = NULL switches it off.  See tools/profile_stream.py. */
int howl_b200_debug_stream_profile(howl_ctx_t* ctx, void* buf, int32_t kind);

/* Cycles (device int64) that `iters` back-to-back M=128 (mode bit 2: 64) x N x 16 bf16 tcgen05.mma take on one SM.
 * mode bit 0: A operand from tensor memory, bit 1: B operand MN-major. */
int howl_b200_debug_umma_bench(howl_ctx_t* ctx, void* stream, int32_t mode, int32_t N, int32_t iters, long long* cycles);

/*
 * Test hook for the mask-forced gradient oracle (tests/test_gpu_parity.py): the ReLU decisions the backward of the
 * forward kept in `workspace` takes, as bytes (1 = gradient passes).
 *   mask0   [B, 45, 3*H, n_mels]  conv0 pre-activation > 0 at the pixels the (3,4) average pooling keeps (H = frames / 3)
 *   masks16 [6, B, 45, H, 10]     layers 1..6: relu(conv_i) > 0, i.e. u_i > 0 (odd layers) or u_i > residual (even layers)
 * Either pointer may be NULL.  feats / params as given to howl_b200_res8_fwd.
 */
int howl_b200_res8_debug_masks(howl_ctx_t* ctx, void* stream, const float* feats, const float* params, int64_t B,
                               int32_t frames, int32_t n_mels, int32_t num_labels, const void* workspace,
                               size_t workspace_bytes, uint8_t* mask0, uint8_t* masks16);

/* Test hooks for the tensor-core GEMMs of the MobileNetV2 path on plain fp32 row-major matrices (converted to the bf16 tile-major
 * operand format inside):  C[M,N] = A[M,K] * W[N,K]^T (+ add[M,N]);   dW[N,K] = dC[M,N]^T * A[M,K]. */
int64_t howl_b200_debug_mbn_workspace_bytes(int64_t M, int K, int N);
int howl_b200_debug_mbn_gemm(howl_ctx_t* ctx, void* stream, const float* A, const float* W, const float* add, float* C, int64_t M,
                             int32_t K, int32_t N, void* workspace, size_t workspace_bytes);
int howl_b200_debug_mbn_wgrad(howl_ctx_t* ctx, void* stream, const float* dC, const float* A, float* dW, int64_t M, int32_t N,
                              int32_t K, void* workspace, size_t workspace_bytes);

/* Test hook for the mask-forced gradient oracle of the MobileNetV2 path: the activation decisions the backward of the forward kept in
 * `workspace` takes, as bytes (1 = gradient passes), concatenated:
 *   stem      [B, 3, n_mels, frames + 4]   conv output that wins its (1,2) max-pooling pair AND is positive (ReLU)
 *   per conv with ReLU6, in network order: [B * hout * wout, cout] (NHWC)   0 < BatchNorm output < 6 */
int64_t howl_b200_mobilenet_debug_mask_bytes(int64_t B, int32_t frames, int32_t n_mels);
int howl_b200_mobilenet_debug_masks(howl_ctx_t* ctx, void* stream, const float* feats, const float* params, int64_t B, int32_t frames,
                                    int32_t n_mels, int32_t num_labels, const void* workspace, size_t workspace_bytes, uint8_t* out);

#ifdef __cplusplus
}
#endif
#endif /* HOWL_B200_DEBUG_H_ */
