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

// ---------------------------------------------------------------------------------------------------------------
// device-side waveform augmentation fused into the batch gather.  The reference augments whole clips on the host
// before batching (training/run/train.py:202-221: DatasetMixer -> TimeshiftTransform -> NoiseTransform -> batchifier,
// howl/data/transform/transform.py:120-231); here the host only replays the draws and row r of the batch is produced in ONE pass:
//   x[j]   = clips[starts[r] + j]                                   (the time shift is a start offset / a shorter count)
//   x[j]   = x[j] * (1 - alpha) + bg[bg_starts[r] + j] * alpha       (DatasetMixer; two roundings and an add, as torch computes it)
//   x[j]   = clamp(x[j] + clamp(N(0, sigma)), -1, 1)                 (NoiseTransform "white")
//   x[j]   = clamp(x[j] + (Bern(p/2) - Bern(p/2)), -1, 1)            (NoiseTransform "salt_pepper")
// with the noise drawn in the kernel from Philox4x32-10 keyed by (seed; row, sample) -- the same distributions as torch's host
// generator, not the same stream.
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, ctr.x), lo0 = 0xD2511F53u * ctr.x;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, ctr.z), lo1 = 0xCD9E8D57u * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += 0x9E3779B9u;
    key.y += 0xBB67AE85u;
  }
  return ctr;
}

struct GatherAugArgs {
  const float* clips;
  const int64_t* starts;
  const int64_t* counts;
  const int64_t* dst_off;
  const float* bg;            // concatenated background clips, or null
  const int64_t* bg_starts;   // [B], < 0: no mixing for the row
  const double* alpha;        // [B] (double: the reference forms 1 - alpha in double before torch rounds both weights to fp32)
  const float* sigma;         // [B] white-noise strength, 0: off
  const float* sp_prob;       // [B] salt-and-pepper probability, 0: off
  unsigned long long seed;
  int64_t max_length;
  float* out;
};

__global__ void __launch_bounds__(256) batch_gather_aug_kernel(const GatherAugArgs a) {
  const int64_t r = blockIdx.y;
  const int64_t n = a.counts[r], d0 = a.dst_off[r];
  const float* src = a.clips + a.starts[r];
  const int64_t bgs = (a.bg && a.bg_starts) ? a.bg_starts[r] : -1;
  const double al64 = a.alpha ? a.alpha[r] : 0.0;
  const float al = (float)al64, one_m = (float)(1.0 - al64);      // python computes 1 - alpha in double, torch rounds both weights to fp32
  const float sg = a.sigma ? a.sigma[r] : 0.f, sp = a.sp_prob ? a.sp_prob[r] : 0.f;
  float* dst = a.out + r * a.max_length;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.max_length; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t j = i - d0;
    float x = 0.f;
    if (j >= 0 && j < n) {
      x = __ldg(src + j);
      if (bgs >= 0) x = __fadd_rn(__fmul_rn(x, one_m), __fmul_rn(__ldg(a.bg + bgs + j), al));
      if (sg > 0.f || sp > 0.f) {
        const uint4 rnd = philox4x32_10(make_uint4((uint32_t)j, (uint32_t)(j >> 32), (uint32_t)r, 0u),
                                        make_uint2((uint32_t)a.seed, (uint32_t)(a.seed >> 32)));
        if (sg > 0.f) {
          const float u1 = ((float)(rnd.x >> 8) + 0.5f) * (1.f / 16777216.f), u2 = ((float)(rnd.y >> 8) + 0.5f) * (1.f / 16777216.f);
          const float nz = sqrtf(-2.f * logf(u1)) * cospif(2.f * u2) * sg;
          x = fminf(fmaxf(x + fminf(fmaxf(nz, -1.f), 1.f), -1.f), 1.f);
        }
        if (sp > 0.f) {
          const float half = 0.5f * sp;
          const float m = (((float)(rnd.z >> 8) * (1.f / 16777216.f)) < half ? 1.f : 0.f) - (((float)(rnd.w >> 8) * (1.f / 16777216.f)) < half ? 1.f : 0.f);
          x = fminf(fmaxf(x + m, -1.f), 1.f);
        }
      }
    }
    dst[i] = x;
  }
}

extern "C" int howl_b200_batch_gather_aug(howl_ctx_t* ctx, void* stream, const float* clips, const int64_t* starts, const int64_t* counts,
                                          const int64_t* dst_off, int64_t B, int64_t max_length, const float* bg, const int64_t* bg_starts,
                                          const double* alpha, const float* sigma, const float* sp_prob, uint64_t seed, float* out) {
  if (!ctx) return HOWL_E_INVALID;
  HOWL_REQUIRE(ctx, clips && starts && counts && dst_off && out, HOWL_E_INVALID, "batch_gather_aug: null pointer");
  HOWL_REQUIRE(ctx, B >= 0 && B <= 65535 && max_length >= 0, HOWL_E_INVALID, "batch_gather_aug: bad shape");
  HOWL_REQUIRE(ctx, (bg != nullptr) == (bg_starts != nullptr) && (!bg || alpha), HOWL_E_INVALID, "batch_gather_aug: bg, bg_starts and alpha go together");
  if (B == 0 || max_length == 0) return HOWL_OK;
  GatherAugArgs a;
  a.clips = clips; a.starts = starts; a.counts = counts; a.dst_off = dst_off; a.bg = bg; a.bg_starts = bg_starts; a.alpha = alpha;
  a.sigma = sigma; a.sp_prob = sp_prob; a.seed = seed; a.max_length = max_length; a.out = out;
  int64_t bx = howl_ceil_div(max_length, 256 * 4);
  if (bx > 64) bx = 64;
  batch_gather_aug_kernel<<<dim3((unsigned)bx, (unsigned)B), 256, 0, (cudaStream_t)stream>>>(a);
  HOWL_LAUNCHED(ctx, "batch_gather_aug");
  return HOWL_OK;
}
