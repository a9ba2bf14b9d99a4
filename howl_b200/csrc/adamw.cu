// K4 -- fused AdamW over the flat parameter buffer, and the whole-step convenience entry point.
// torch.optim.AdamW semantics (training/run/train.py:256,302):
//   p *= 1 - lr*wd;  m = b1 m + (1-b1) g;  v = b2 v + (1-b2) g^2;  p -= lr/(1-b1^t) * m / (sqrt(v)/sqrt(1-b2^t) + eps)
#include <math.h>

#include "common.cuh"

__global__ void __launch_bounds__(256) adamw_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                    float* __restrict__ m, float* __restrict__ v, int64_t n,
                                                    float decay, float b1, float b2, float eps, float step_size,
                                                    float inv_sqrt_bc2) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float gi = g[i];
    float pi = p[i] * decay;
    const float mi = b1 * m[i] + (1.f - b1) * gi;        // lerp form == torch's mul_/add_ to 1 ulp
    const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    const float denom = sqrtf(vi) * inv_sqrt_bc2 + eps;
    pi -= step_size * __fdiv_rn(mi, denom);
    p[i] = pi;
  }
}

extern "C" int howl_b200_adamw(howl_ctx_t* ctx, void* stream, float* params, const float* grads, float* exp_avg,
                               float* exp_avg_sq, int64_t n, int64_t step, float lr, float beta1, float beta2,
                               float eps, float weight_decay) {
  if (!ctx) return HOWL_E_INVALID;
  HOWL_REQUIRE(ctx, params && grads && exp_avg && exp_avg_sq, HOWL_E_INVALID, "adamw: null pointer");
  HOWL_REQUIRE(ctx, n >= 0 && step >= 1, HOWL_E_INVALID, "adamw: n=%lld step=%lld", (long long)n, (long long)step);
  if (n == 0) return HOWL_OK;
  const double bc1 = 1.0 - pow((double)beta1, (double)step);
  const double bc2 = 1.0 - pow((double)beta2, (double)step);
  const float step_size = (float)((double)lr / bc1);
  const float inv_sqrt_bc2 = (float)(1.0 / sqrt(bc2));
  const float decay = (float)(1.0 - (double)lr * (double)weight_decay);
  int64_t blocks = howl_ceil_div(n, 256);
  if (blocks > (int64_t)ctx->sm_count * 8) blocks = (int64_t)ctx->sm_count * 8;
  adamw_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(params, grads, exp_avg, exp_avg_sq, n, decay, beta1,
                                                                   beta2, eps, step_size, inv_sqrt_bc2);
  HOWL_LAUNCHED(ctx, "adamw");
  return HOWL_OK;
}

extern "C" int howl_b200_res8_train_step(howl_ctx_t* ctx, void* stream, const float* pcm, const int64_t* labels,
                                         int64_t B, int64_t T, const float* fb, float zmuv_mean, float zmuv_std,
                                         const int32_t* rects, int32_t num_labels, float* params, float* bn_running,
                                         int64_t* num_batches_tracked, float* grads, float* exp_avg,
                                         float* exp_avg_sq, int64_t step, float lr, float weight_decay, float* loss,
                                         float* logits, void* workspace, size_t workspace_bytes) {
  if (!ctx) return HOWL_E_INVALID;
  HOWL_REQUIRE(ctx, workspace, HOWL_E_WORKSPACE, "train_step: null workspace");
  const int M = ctx->fe.n_mels;
  const int64_t F = howl_b200_num_frames(T, ctx->fe.hop);
  HOWL_REQUIRE(ctx, F > 0 && F < (1 << 20), HOWL_E_INVALID, "train_step: bad clip length %lld", (long long)T);
  // features live at the head of the workspace, the res8 workspace follows
  const size_t feat_bytes = howl_align_up(sizeof(float) * (size_t)B * F * M, 256);
  HOWL_REQUIRE(ctx, workspace_bytes > feat_bytes, HOWL_E_WORKSPACE, "train_step: workspace too small");
  float* feats = (float*)workspace;
  void* ws = (char*)workspace + feat_bytes;
  const size_t ws_bytes = workspace_bytes - feat_bytes;
  int rc = howl_b200_frontend_fwd(ctx, stream, pcm, B, T, fb, zmuv_mean, zmuv_std, rects,
                                  HOWL_FE_TIME_MAJOR | HOWL_FE_ZMUV, feats);
  if (rc) return rc;
  rc = howl_b200_res8_fwd(ctx, stream, feats, B, (int)F, M, num_labels, params, bn_running, num_batches_tracked, 1,
                          logits, ws, ws_bytes);
  if (rc) return rc;
  rc = howl_b200_res8_bwd(ctx, stream, feats, labels, B, (int)F, M, num_labels, B, params, grads, loss, ws, ws_bytes);
  if (rc) return rc;
  return howl_b200_adamw(ctx, stream, params, grads, exp_avg, exp_avg_sq, howl_b200_res8_param_count(num_labels),
                         step, lr, 0.9f, 0.999f, 1e-8f, weight_decay);
}
