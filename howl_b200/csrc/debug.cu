// Tuning / self-test kernels of the tensor-core path.  NOT part of the drop-in surface: declared in
// include/howl_b200_debug.h, used only by tools/ and tests/.
#include "common.cuh"
#include "tc_common.cuh"
#include "../../include/howl_b200_debug.h"

// =============================================================================================
// tuning aid: issue-to-retire cost of one tcgen05.mma shape on an otherwise idle SM.  One CTA issues `iters` M = 128, K = 16
// bf16 MMAs into one accumulator (operands are whatever shared memory holds -- only the timing matters) and reports the
// cycles from first issue to the commit's arrival.  mode bit 0: A from tensor memory (TS) instead of shared memory (SS);
// bit 1: B MN-major instead of K-major; bit 2: M = 64.
// =============================================================================================
__global__ void __launch_bounds__(128, 1) umma_bench_kernel(int mode, int N, int iters, long long* __restrict__ out) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t s_tmem;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp == 0) tc::tmem_alloc<512>(&s_tmem);
  if (tid == 32) {
    tc::mbar_init(&bar, 1);
    tc::fence_barrier_init();
  }
  for (int i = tid; i < (128 * 2 + 256 * 2 + 64) * 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0u;
  tc::fence_proxy_async();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = s_tmem;
  if (warp == 0 && tc::elect_one()) {
    const int M = (mode & 4) ? 64 : 128;
    const uint32_t sa = tc::smem_u32(smem), sb = sa + 128 * 32;
    const uint64_t ad = tc::smem_desc(sa, 128 * 16, 128);
    const uint64_t bd = (mode & 2) ? tc::smem_desc(sb, 128, 16 * 16) : tc::smem_desc(sb, (uint32_t)N * 16, 128);
    const uint32_t idesc = tc::instr_desc_bf16(M, N, 0, (mode & 2) ? 1 : 0);
    const uint32_t a_t = tmem + 256u;
    if (mode & 1) {
      tc::tmem_cp_128x256b(a_t, ad);
      if (mode & 32) tc::tmem_cp_128x256b(tmem + 432u, ad);
    }
    // bit 3: rotate over 9 accumulators (48 columns apart, two MMAs each); bit 4: operands start one 16-byte row off the
    // 128-byte core-matrix alignment (what a 3x3 tap shift does); bit 5: a tcgen05.cp of a fresh A tile every 18 MMAs
    const uint32_t mis = (mode & 16) ? 1u : 0u;
    const uint64_t ad2 = ad + mis, bd2 = bd + mis;
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      const uint32_t d = (mode & 8) ? tmem + 48u * (uint32_t)((i >> 1) % 9) : tmem;
      const uint32_t acc = (mode & 8) ? (i >= 18 ? 1u : (uint32_t)(i & 1)) : (i ? 1u : 0u);
      const uint32_t at = (mode & 32) ? 432u + 8u * (uint32_t)((i / 18) & 7) : 256u;
      if ((mode & 32) && i % 18 == 0) tc::tmem_cp_128x256b(tmem + 432u + 8u * (uint32_t)(((i / 18) + 1) & 7), ad);
      if (mode & 1) tc::umma_bf16_ts(d, tmem + at, bd2, idesc, acc);
      else tc::umma_bf16(d, ad2, bd2, idesc, acc);
    }
    tc::umma_commit(&bar);
    tc::mbar_wait(&bar, 0);
    out[0] = clock64() - t0;
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc<512>(tmem);
}

extern "C" int howl_b200_debug_umma_bench(howl_ctx_t* ctx, void* stream, int32_t mode, int32_t N, int32_t iters, long long* cycles) {
  if (!ctx) return HOWL_E_INVALID;
  HOWL_REQUIRE(ctx, cycles && iters > 0 && N >= 16 && N <= 256 && N % 16 == 0, HOWL_E_INVALID, "umma_bench: bad arguments");
  const size_t smem = (128 * 2 + 256 * 2 + 64) * 16;   // + slack for the misaligned variants
  umma_bench_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(mode, N, iters, cycles);
  HOWL_LAUNCHED(ctx, "umma_bench");
  return HOWL_OK;
}

// =============================================================================================
// descriptor self-test: two small GEMMs through exactly the helpers above
//   test 0 (K-major):  D[128][48] = A[128][32] * B[48][32]^T       rows at 16 B, chunk stride = rows * 16
//   test 1 (MN-major): D[128][48] = A[32][128]^T * B[32][48]       K rows at 16 B, 8-wide MN groups at chunk stride
// inputs are fp32 [M][K] / [N][K] (test 0) or [K][M] / [K][N] (test 1); bf16-rounded inside; D fp32 [128][48].
// =============================================================================================
__global__ void __launch_bounds__(128, 1) umma_selftest_kernel(const float* __restrict__ A, const float* __restrict__ B,
                                                               float* __restrict__ D, int mn_major, int variant) {
  extern __shared__ __align__(128) unsigned char smem[];
  uint4* sa = reinterpret_cast<uint4*>(smem);   // K-major: [4 chunks][128 rows]; MN-major: [16 chunks of 8 m][32 k]
  uint4* sb = sa + 4 * 128;                     // K-major: [4 chunks][48 rows];  MN-major: [6 chunks of 8 n][32 k]
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t s_tmem;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (warp == 0) tc::tmem_alloc<64>(&s_tmem);
  if (tid == 32) {
    tc::mbar_init(&bar, 1);
    tc::fence_barrier_init();
  }
  auto pack = [](const float* v) {
    uint32_t h[4];
    for (int i = 0; i < 4; ++i)
      h[i] = (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(v[2 * i])) |
             ((uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(v[2 * i + 1])) << 16);
    return make_uint4(h[0], h[1], h[2], h[3]);
  };
  if (mn_major == 2) {   // test 2: A K-major (copied to tensor memory with tcgen05.cp), B MN-major
    for (int i = tid; i < 4 * 128; i += 128) {
      const int chunk = i / 128, row = i % 128;
      float v[8];
      for (int j = 0; j < 8; ++j) v[j] = A[row * 32 + chunk * 8 + j];
      sa[chunk * 128 + row] = pack(v);
    }
    for (int i = tid; i < 6 * 32; i += 128) {
      const int chunk = i / 32, k = i % 32;
      float v[8];
      for (int j = 0; j < 8; ++j) v[j] = B[k * 48 + chunk * 8 + j];
      sb[chunk * 32 + k] = pack(v);
    }
  } else if (!mn_major) {
    for (int i = tid; i < 4 * 128; i += 128) {
      const int chunk = i / 128, row = i % 128;
      float v[8];
      for (int j = 0; j < 8; ++j) v[j] = A[row * 32 + chunk * 8 + j];
      sa[chunk * 128 + row] = pack(v);
    }
    for (int i = tid; i < 4 * 48; i += 128) {
      const int chunk = i / 48, row = i % 48;
      float v[8];
      for (int j = 0; j < 8; ++j) v[j] = B[row * 32 + chunk * 8 + j];
      sb[chunk * 48 + row] = pack(v);
    }
  } else {
    for (int i = tid; i < 16 * 32; i += 128) {
      const int chunk = i / 32, k = i % 32;
      float v[8];
      for (int j = 0; j < 8; ++j) v[j] = A[k * 128 + chunk * 8 + j];
      sa[chunk * 32 + k] = pack(v);
    }
    for (int i = tid; i < 6 * 32; i += 128) {
      const int chunk = i / 32, k = i % 32;
      float v[8];
      for (int j = 0; j < 8; ++j) v[j] = B[k * 48 + chunk * 8 + j];
      sb[chunk * 32 + k] = pack(v);
    }
  }
  tc::fence_proxy_async();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = s_tmem;
  if (tid == 0) {
    const uint32_t sa_s = tc::smem_u32(sa), sb_s = tc::smem_u32(sb);
    for (int ks = 0; ks < 2; ++ks) {
      uint64_t ad, bd;
      uint32_t idesc;
      if (mn_major == 2) {
        const uint32_t a_tmem = tmem + 48u + 8u * (uint32_t)ks;
        tc::tmem_cp_128x256b(a_tmem, tc::smem_desc(sa_s + (2 * ks) * 128 * 16, 128 * 16, 128));
        bd = tc::smem_desc(sb_s + ks * 16 * 16, 128, 32 * 16);
        tc::umma_bf16_ts(tmem, a_tmem, bd, tc::instr_desc_bf16(128, 48, 0, 1), ks ? 1u : 0u);
        continue;
      }
      if (!mn_major) {
        uint32_t lbo_a = 128 * 16, sbo_a = 128, lbo_b = 48 * 16, sbo_b = 128;
        if (variant & 1) { uint32_t t = lbo_a; lbo_a = sbo_a; sbo_a = t; t = lbo_b; lbo_b = sbo_b; sbo_b = t; }
        ad = tc::smem_desc(sa_s + (2 * ks) * 128 * 16, lbo_a, sbo_a);
        bd = tc::smem_desc(sb_s + (2 * ks) * 48 * 16, lbo_b, sbo_b);
        idesc = tc::instr_desc_bf16(128, 48, 0, 0);
      } else {
        uint32_t lbo = 128, sbo_a = 32 * 16, sbo_b = 32 * 16;
        if (variant & 1) {
          ad = tc::smem_desc(sa_s + ks * 16 * 16, sbo_a, lbo);
          bd = tc::smem_desc(sb_s + ks * 16 * 16, sbo_b, lbo);
        } else {
          ad = tc::smem_desc(sa_s + ks * 16 * 16, lbo, sbo_a);
          bd = tc::smem_desc(sb_s + ks * 16 * 16, lbo, sbo_b);
        }
        idesc = tc::instr_desc_bf16(128, 48, 1, 1);
      }
      tc::umma_bf16(tmem, ad, bd, idesc, ks ? 1u : 0u);
    }
    tc::umma_commit(&bar);
  }
  tc::mbar_wait(&bar, 0);
  tc::fence_after_sync();
  float v[48];
  const uint32_t taddr = tmem + ((uint32_t)(32 * warp) << 16);
  tc::tmem_ld16(taddr, v);
  tc::tmem_ld16(taddr + 16, v + 16);
  tc::tmem_ld16(taddr + 32, v + 32);
  for (int n = 0; n < 48; ++n) D[(32 * warp + lane) * 48 + n] = v[n];
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc<64>(tmem);
}

extern "C" int howl_b200_selftest_umma(howl_ctx_t* ctx, void* stream, const float* A, const float* B, float* D,
                                       int32_t mn_major, int32_t variant) {
  if (!ctx) return HOWL_E_INVALID;
  HOWL_REQUIRE(ctx, A && B && D, HOWL_E_INVALID, "selftest_umma: null pointer");
  const size_t smem = (4 * 128 + 16 * 32 + 6 * 32 + 4 * 48) * 16 + 1024;
  HOWL_CUDA(ctx, cudaFuncSetAttribute(umma_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  umma_selftest_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(A, B, D, mn_major, variant);
  HOWL_LAUNCHED(ctx, "umma_selftest");
  return HOWL_OK;
}

