// Small element-wise operators of the howl transform API that are not always fused into K1:
//   ZmuvTransform.forward on an arbitrary tensor (howl/data/transform/operator.py:145-146),
//   SpecAugmentTransform masks on [B,C,M,F] (howl/data/transform/transform.py:310-326),
//   the x[:, :1].permute(0,1,3,2).contiguous() at the top of Res8.forward (howl/model/cnn.py:128-129).
#include "common.cuh"

__global__ void zmuv_kernel(const float* __restrict__ x, float* __restrict__ y, int64_t n, float mean, float std) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    y[i] = __fdiv_rn(x[i] - mean, std);
}

__global__ void spec_mask_kernel(float* __restrict__ x, int64_t B, int C, int M, int F, const int32_t* __restrict__ rects) {
  const int64_t n = B * C * M * (int64_t)F;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int f = (int)(i % F);
    const int m = (int)((i / F) % M);
    const int64_t b = i / ((int64_t)F * M * C);
    const int f0 = rects[b * 4], fl = rects[b * 4 + 1], t0 = rects[b * 4 + 2], tl = rects[b * 4 + 3];
    if ((m >= f0 && m < f0 + fl) || (f >= t0 && f < t0 + tl)) x[i] = 0.f;
  }
}

// out[b][f][m] = x[b][0][m][f]   (tile transpose through shared memory; x has C channels)
__global__ void to_time_major_kernel(const float* __restrict__ x, float* __restrict__ out, int C, int M, int F) {
  __shared__ float tile[32][33];
  const int64_t b = blockIdx.z;
  const int f0 = blockIdx.x * 32, m0 = blockIdx.y * 32;
  const float* src = x + b * (int64_t)C * M * F;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int m = m0 + r, f = f0 + threadIdx.x;
    if (m < M && f < F) tile[r][threadIdx.x] = src[(int64_t)m * F + f];
  }
  __syncthreads();
  float* dst = out + b * (int64_t)F * M;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int f = f0 + r, m = m0 + threadIdx.x;
    if (m < M && f < F) dst[(int64_t)f * M + m] = tile[threadIdx.x][r];
  }
}

static unsigned ew_blocks(howl_ctx_t* ctx, int64_t n) {
  int64_t b = howl_ceil_div(n, 256);
  const int64_t cap = (int64_t)ctx->sm_count * 16;
  return (unsigned)(b < cap ? (b < 1 ? 1 : b) : cap);
}

extern "C" int howl_b200_zmuv_fwd(howl_ctx_t* ctx, void* stream, const float* x, int64_t n, float mean, float std,
                                  float* out) {
  if (!ctx) return HOWL_E_INVALID;
  HOWL_REQUIRE(ctx, x && out && n >= 0, HOWL_E_INVALID, "zmuv_fwd: bad argument");
  HOWL_REQUIRE(ctx, std > 0.f, HOWL_E_INVALID, "zmuv_fwd: std must be > 0");
  if (n == 0) return HOWL_OK;
  zmuv_kernel<<<ew_blocks(ctx, n), 256, 0, (cudaStream_t)stream>>>(x, out, n, mean, std);
  HOWL_LAUNCHED(ctx, "zmuv");
  return HOWL_OK;
}

extern "C" int howl_b200_spec_mask(howl_ctx_t* ctx, void* stream, float* x, int64_t B, int32_t C, int32_t M, int32_t F,
                                   const int32_t* rects) {
  if (!ctx) return HOWL_E_INVALID;
  HOWL_REQUIRE(ctx, x && rects && B >= 0 && C > 0 && M > 0 && F > 0, HOWL_E_INVALID, "spec_mask: bad argument");
  if (B == 0) return HOWL_OK;
  spec_mask_kernel<<<ew_blocks(ctx, B * C * M * (int64_t)F), 256, 0, (cudaStream_t)stream>>>(x, B, C, M, F, rects);
  HOWL_LAUNCHED(ctx, "spec_mask");
  return HOWL_OK;
}

extern "C" int howl_b200_to_time_major(howl_ctx_t* ctx, void* stream, const float* x, int64_t B, int32_t C, int32_t M,
                                       int32_t F, float* out) {
  if (!ctx) return HOWL_E_INVALID;
  HOWL_REQUIRE(ctx, x && out && B >= 0 && C > 0 && M > 0 && F > 0, HOWL_E_INVALID, "to_time_major: bad argument");
  HOWL_REQUIRE(ctx, B <= 65535, HOWL_E_UNSUPPORTED, "to_time_major: batch > 65535");
  if (B == 0) return HOWL_OK;
  dim3 grid((unsigned)howl_ceil_div(F, 32), (unsigned)howl_ceil_div(M, 32), (unsigned)B), block(32, 8);
  to_time_major_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(x, out, C, M, F);
  HOWL_LAUNCHED(ctx, "to_time_major");
  return HOWL_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// batch gather: the data movement of WakeWordFrameBatchifier.__call__ + tensorize_audio_data
// (howl/data/transform/batchifier.py:56-118, operator.py:89-109) once the host has drawn the plan: row r of the batch is
// counts[r] samples starting at clips[starts[r]], placed at column dst_off[r] of a zero row of max_length samples.
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) batch_gather_kernel(const float* __restrict__ clips, const int64_t* __restrict__ starts,
                                                           const int64_t* __restrict__ counts,
                                                           const int64_t* __restrict__ dst_off, int64_t max_length,
                                                           float* __restrict__ out) {
  const int64_t r = blockIdx.y;
  const int64_t n = counts[r], d0 = dst_off[r];
  const float* src = clips + starts[r];
  float* dst = out + r * max_length;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < max_length; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t j = i - d0;
    dst[i] = (j >= 0 && j < n) ? __ldg(src + j) : 0.f;
  }
}

extern "C" int howl_b200_batch_gather(howl_ctx_t* ctx, void* stream, const float* clips, const int64_t* starts,
                                      const int64_t* counts, const int64_t* dst_off, int64_t B, int64_t max_length,
                                      float* out) {
  if (!ctx) return HOWL_E_INVALID;
  HOWL_REQUIRE(ctx, clips && starts && counts && dst_off && out, HOWL_E_INVALID, "batch_gather: null pointer");
  HOWL_REQUIRE(ctx, B >= 0 && B <= 65535 && max_length >= 0, HOWL_E_INVALID, "batch_gather: bad shape");
  if (B == 0 || max_length == 0) return HOWL_OK;
  int64_t bx = howl_ceil_div(max_length, 256 * 4);
  if (bx > 64) bx = 64;
  batch_gather_kernel<<<dim3((unsigned)bx, (unsigned)B), 256, 0, (cudaStream_t)stream>>>(clips, starts, counts, dst_off,
                                                                                        max_length, out);
  HOWL_LAUNCHED(ctx, "batch_gather");
  return HOWL_OK;
}
