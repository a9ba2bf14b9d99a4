// tcgen05 GEMMs of the MobileNetV2 path (pointwise convolutions, their data and weight gradients) over the tile-major operand format
// (mbn_common.cuh).  sm_100a only.
//
//  * mbn_gemm_nt_kernel: C = A * W^T.  Persistent CTAs walk (row tile, N tile) items.  Warp roles: one TMA loader (a row tile's K
//    range and the matching slice of the weight operand are each ONE contiguous bulk copy per 64-wide K stage), one MMA issuer
//    (M = 128, N = the N tile <= 240, K = 16 per tcgen05.mma, two accumulators alternating in TMEM), four epilogue warps (thread =
//    one row: tcgen05.ld 8 columns -> (+ Add) -> bf16 -> one 16-byte store straight into the output tile, which the layout makes
//    a contiguous 512 bytes per warp).
//  * mbn_gemm_wgrad_kernel: dW = dC^T * A, the reduction runs over the ROWS, so both operands are read MN-major straight from the
//    same tiles (K = 16 rows per MMA); the [128 x Kt] accumulator stays in TMEM across all row tiles of a CTA and is added to the
//    fp32 gradient with atomics once.
//  * fp32 products for the LSTM / LAS paths reuse both kernels with the operands split into (hi, lo) bf16: A^T B as three operand pairs
//    streaming through one accumulator (mbn_atb3_packed), X W^T with the split folded into K ([hi|hi|lo] x [hi|lo|hi], mbn_gemm_nt3_f32)
//    and an fp32 row-major epilogue (+ bias, ReLU, accumulate).
#include <algorithm>

#include "mbn_common.cuh"
#include "tc_common.cuh"
#include "../../include/howl_b200_debug.h"

#define MG_THREADS 192           // warps 0-3 epilogue, warp 4 MMA issue, warp 5 TMA loader
#define MG_STAGES 4
#define MG_KSTAGE 8              // chunks (of 8 channels) per pipeline stage = 64 K
#define MG_NT_MAX 240

int mbn_ntile(int np) {
  for (int nt = MG_NT_MAX; nt >= 16; nt -= 16)
    if (np % nt == 0) return nt;
  return 16;
}

size_t mbn_weight_operand_bytes(int n, int k) { return (size_t)mbn_pad16(n) * mbn_pad16(k) * 2; }

__global__ void mbn_weight_operand_kernel(const float* __restrict__ w, int n, int k, int ld, int transpose, int np, int kp, int nt,
                                          __nv_bfloat16* __restrict__ out) {
  const int total = np * kp;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    // i -> (tile, chunk, row, j)
    const int j = i & 7, row = (i >> 3) % nt, chunk = ((i >> 3) / nt) % (kp / 8), tile = (i >> 3) / (nt * (kp / 8));
    const int nn = tile * nt + row, kk = chunk * 8 + j;
    float v = 0.f;
    if (nn < n && kk < k) v = transpose ? w[(size_t)kk * ld + nn] : w[(size_t)nn * ld + kk];
    out[i] = __float2bfloat16_rn(v);
  }
}

// The same conversion for up to MBN_WOP_BATCH matrices in ONE launch (grid.y = matrix): a MobileNetV2 step prepares 35 + 35 operands of a
// few KB each, i.e. 70 launches that are pure launch latency when issued one by one.
__global__ void mbn_weight_operand_batch_kernel(const MbnWopBatch batch) {
  const MbnWopDesc d = batch.d[blockIdx.y];
  const int np = mbn_pad16(d.n), kp = mbn_pad16(d.k), nt = d.nt;
  const int total = np * kp;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int j = i & 7, row = (i >> 3) % nt, chunk = ((i >> 3) / nt) % (kp / 8), tile = (i >> 3) / (nt * (kp / 8));
    const int nn = tile * nt + row, kk = chunk * 8 + j;
    float v = 0.f;
    if (nn < d.n && kk < d.k) v = d.transpose ? d.w[(size_t)kk * d.ld + nn] : d.w[(size_t)nn * d.ld + kk];
    d.out[i] = __float2bfloat16_rn(v);
  }
}

int mbn_weight_operand_batch(howl_ctx_t* ctx, cudaStream_t st, const MbnWopDesc* descs, int count) {
  for (int first = 0; first < count; first += MBN_WOP_BATCH) {
    MbnWopBatch batch;
    const int m = count - first < MBN_WOP_BATCH ? count - first : MBN_WOP_BATCH;
    for (int i = 0; i < m; ++i) {
      batch.d[i] = descs[first + i];
      HOWL_REQUIRE(ctx, batch.d[i].w && batch.d[i].out && batch.d[i].n > 0 && batch.d[i].k > 0, HOWL_E_INVALID, "mbn_weight_operand_batch: bad matrix %d", first + i);
      batch.d[i].nt = mbn_ntile(mbn_pad16(batch.d[i].n));
    }
    mbn_weight_operand_batch_kernel<<<dim3(48, (unsigned)m), 256, 0, st>>>(batch);
    HOWL_LAUNCHED(ctx, "mbn_weight_operand");
  }
  return HOWL_OK;
}

int mbn_weight_operand(howl_ctx_t* ctx, cudaStream_t st, const float* w, int n, int k, int ld, int transpose, __nv_bfloat16* out) {
  const int np = mbn_pad16(n), kp = mbn_pad16(k), nt = mbn_ntile(np);
  const int total = np * kp;
  mbn_weight_operand_kernel<<<(total + 255) / 256, 256, 0, st>>>(w, n, k, ld, transpose, np, kp, nt, out);
  HOWL_LAUNCHED(ctx, "mbn_weight_operand");
  return HOWL_OK;
}

// =============================================================================================
struct MgArgs {
  const __nv_bfloat16* A;
  const __nv_bfloat16* W;
  const __nv_bfloat16* add;
  __nv_bfloat16* C;
  int64_t M, m_tiles;
  int k8, n8, nt, n_tiles;
  // fp32 row-major output (C32 != null): C32[row * ldc + n] = acc (+ bias[n]) (ReLU), n < n_valid
  float* C32;
  int64_t ldc;
  const float* bias;
  int n_valid, relu, accumulate;     // accumulate: C32 += (the second direction of a bidirectional product)
};

template <bool F32>
__global__ void __launch_bounds__(MG_THREADS, 1) mbn_gemm_nt_kernel(const MgArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(8) uint64_t bar_full[MG_STAGES], bar_empty[MG_STAGES], bar_acc[2], bar_accfree[2];
  __shared__ uint32_t s_tmem;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nt = a.nt, k8 = a.k8;
  const uint32_t a_stage = MG_KSTAGE * 2048u, b_stage = (uint32_t)MG_KSTAGE * nt * 16u, stage_bytes = a_stage + b_stage;
  const int kst = (k8 + MG_KSTAGE - 1) / MG_KSTAGE;
  const int64_t items = a.m_tiles * a.n_tiles;
  if (warp == 4) {
    tc::tmem_alloc<512>(&s_tmem);
    if (lane == 0) {
      for (int i = 0; i < MG_STAGES; ++i) {
        tc::mbar_init(&bar_full[i], 1);
        tc::mbar_init(&bar_empty[i], 1);
      }
      for (int i = 0; i < 2; ++i) {
        tc::mbar_init(&bar_acc[i], 1);
        tc::mbar_init(&bar_accfree[i], 4);
      }
      tc::fence_barrier_init();
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = s_tmem;

  if (warp == 5) {
    // ================= TMA loader =================
    if (tc::elect_one()) {
      uint32_t it = 0;
      for (int64_t item = blockIdx.x; item < items; item += gridDim.x) {
        const int64_t mt = item / a.n_tiles;
        const int ntile = (int)(item - mt * a.n_tiles);
        const unsigned char* asrc = reinterpret_cast<const unsigned char*>(a.A) + (size_t)mt * k8 * 2048;
        const unsigned char* bsrc = reinterpret_cast<const unsigned char*>(a.W) + (size_t)ntile * k8 * nt * 16;
        for (int ks = 0; ks < kst; ++ks, ++it) {
          const int s = it % MG_STAGES;
          const int nch = min(MG_KSTAGE, k8 - ks * MG_KSTAGE);
          if (it >= MG_STAGES) tc::mbar_wait(&bar_empty[s], ((it / MG_STAGES) - 1) & 1);
          unsigned char* dst = smem + (size_t)s * stage_bytes;
          tc::mbar_expect_tx(&bar_full[s], (uint32_t)nch * (2048u + (uint32_t)nt * 16u));
          tc::tma_bulk_g2s(dst, asrc + (size_t)ks * MG_KSTAGE * 2048, (uint32_t)nch * 2048u, &bar_full[s]);
          tc::tma_bulk_g2s(dst + a_stage, bsrc + (size_t)ks * MG_KSTAGE * nt * 16, (uint32_t)nch * nt * 16u, &bar_full[s]);
        }
      }
    }
    __syncwarp();
  } else if (warp == 4) {
    // ================= MMA issuer =================
    if (tc::elect_one()) {
      const uint32_t idesc = tc::instr_desc_bf16(128, nt, 0, 0);
      const uint32_t hi128 = tc::desc_hi(128u);
      uint32_t it = 0, n_item = 0;
      for (int64_t item = blockIdx.x; item < items; item += gridDim.x, ++n_item) {
        const uint32_t acc = n_item & 1;
        if (n_item >= 2) tc::mbar_wait(&bar_accfree[acc], ((n_item >> 1) - 1) & 1);
        tc::fence_after_sync();
        const uint32_t d = tmem + acc * 256u;
        for (int ks = 0; ks < kst; ++ks, ++it) {
          const int s = it % MG_STAGES;
          const int nch = min(MG_KSTAGE, k8 - ks * MG_KSTAGE);
          tc::mbar_wait(&bar_full[s], (it / MG_STAGES) & 1);
          tc::fence_after_sync();
          const uint32_t sa = tc::smem_u32(smem + (size_t)s * stage_bytes), sb = sa + a_stage;
          for (int k16 = 0; k16 < nch / 2; ++k16) {
            const uint64_t ad = tc::desc_make(tc::desc_lo(sa + (uint32_t)k16 * 4096u, 2048u), hi128);
            const uint64_t bd = tc::desc_make(tc::desc_lo(sb + (uint32_t)k16 * 2u * nt * 16u, (uint32_t)nt * 16u), hi128);
            tc::umma_bf16(d, ad, bd, idesc, (ks | k16) ? 1u : 0u);
          }
          tc::umma_commit(&bar_empty[s]);
        }
        tc::umma_commit(&bar_acc[acc]);
      }
    }
    __syncwarp();
  } else {
    // ================= epilogue: thread = one row of the tile =================
    uint32_t n_item = 0;
    const int r = 32 * warp + lane;
    for (int64_t item = blockIdx.x; item < items; item += gridDim.x, ++n_item) {
      const uint32_t acc = n_item & 1;
      const int64_t mt = item / a.n_tiles;
      const int ntile = (int)(item - mt * a.n_tiles);
      const bool valid = mt * MBN_TILE + r < a.M;
      const size_t vec0 = ((size_t)mt * a.n8 + (size_t)ntile * (nt / 8)) * MBN_TILE + r;     // 16-byte vector of (row, first chunk of the N tile)
      uint4* out = reinterpret_cast<uint4*>(a.C) + vec0;
      const uint4* add = a.add ? reinterpret_cast<const uint4*>(a.add) + vec0 : nullptr;
      tc::mbar_wait(&bar_acc[acc], (n_item >> 1) & 1);
      tc::fence_after_sync();
      const uint32_t taddr = tmem + ((uint32_t)(32 * warp) << 16) + acc * 256u;
      for (int c = 0; c < nt / 8; ++c) {
        float v[8];
        tc::tmem_ld8(taddr + 8 * c, v);
        if (F32) {
          if (valid) {
            const int n0 = ntile * nt + c * 8;
            float* dst = a.C32 + (mt * MBN_TILE + r) * a.ldc + n0;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              if (a.bias && n0 + j < a.n_valid) v[j] += __ldg(a.bias + n0 + j);
              if (a.relu) v[j] = fmaxf(v[j], 0.f);
            }
            if (a.accumulate) {
#pragma unroll
              for (int j = 0; j < 8; ++j)
                if (n0 + j < a.n_valid) v[j] += dst[j];
            }
            if (n0 + 8 <= a.n_valid && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
              reinterpret_cast<float4*>(dst)[0] = make_float4(v[0], v[1], v[2], v[3]);
              reinterpret_cast<float4*>(dst)[1] = make_float4(v[4], v[5], v[6], v[7]);
            } else {
#pragma unroll
              for (int j = 0; j < 8; ++j)
                if (n0 + j < a.n_valid) dst[j] = v[j];
            }
          }
          continue;
        }
        if (add && valid) {
          const uint4 av = __ldg(add + (size_t)c * MBN_TILE);
          const uint32_t w[4] = {av.x, av.y, av.z, av.w};
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            v[2 * i] += __uint_as_float(w[i] << 16);
            v[2 * i + 1] += __uint_as_float(w[i] & 0xffff0000u);
          }
        }
        uint32_t o[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const __nv_bfloat162 h = __floats2bfloat162_rn(valid ? v[2 * i] : 0.f, valid ? v[2 * i + 1] : 0.f);
          o[i] = *reinterpret_cast<const uint32_t*>(&h);
        }
        out[(size_t)c * MBN_TILE] = make_uint4(o[0], o[1], o[2], o[3]);
      }
      tc::fence_before_sync();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&bar_accfree[acc]);
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 4) tc::tmem_dealloc<512>(tmem);
}

int mbn_gemm_nt(howl_ctx_t* ctx, cudaStream_t st, const __nv_bfloat16* A, const __nv_bfloat16* wop, const __nv_bfloat16* add,
                __nv_bfloat16* C, int64_t M, int K, int N) {
  MgArgs a;
  a.A = A; a.W = wop; a.add = add; a.C = C; a.M = M; a.m_tiles = mbn_tiles(M);
  a.C32 = nullptr; a.ldc = 0; a.bias = nullptr; a.n_valid = N; a.relu = 0; a.accumulate = 0;
  const int kp = mbn_pad16(K), np = mbn_pad16(N);
  a.k8 = kp / 8; a.n8 = np / 8; a.nt = mbn_ntile(np); a.n_tiles = np / a.nt;
  HOWL_REQUIRE(ctx, A && wop && C && M > 0, HOWL_E_INVALID, "mbn_gemm: bad argument");
  const size_t smem = (size_t)MG_STAGES * (MG_KSTAGE * 2048 + (size_t)MG_KSTAGE * a.nt * 16);
  HOWL_CUDA(ctx, cudaFuncSetAttribute(mbn_gemm_nt_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int64_t items = a.m_tiles * a.n_tiles;
  const int grid = (int)(items < ctx->sm_count ? items : ctx->sm_count);
  mbn_gemm_nt_kernel<false><<<grid, MG_THREADS, smem, st>>>(a);
  HOWL_LAUNCHED(ctx, "mbn_gemm");
  return HOWL_OK;
}

// =============================================================================================
// weight gradient: dW[n][k] += sum_rows dC[row][n] * A[row][k]
// =============================================================================================
struct MwArgs {
  const __nv_bfloat16* dC[3];   // `products` operand pairs accumulated into the same tile (1: plain bf16; 3: the hi/lo split of an fp32 product)
  const __nv_bfloat16* A[3];
  int products;
  float* dW;
  int64_t m_tiles;
  int n8, k8;          // chunks of dC / A
  int kt, k_tiles;     // A-channel tile (MMA N) and count
  int n_tiles;         // tiles of 128 dC channels (MMA M)
  int slices;          // row-tile slices
  int n_valid, k_valid, ld;
};
#define MW_THREADS 192
#define MW_STAGES 2

template <int PRODUCTS>
__global__ void __launch_bounds__(MW_THREADS, 1) mbn_gemm_wgrad_kernel(const MwArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(8) uint64_t bar_full[MW_STAGES], bar_empty[MW_STAGES], bar_done;
  __shared__ uint32_t s_tmem;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // blockIdx.x -> (n tile, k tile, slice)
  const int slice = blockIdx.x % a.slices, pair = blockIdx.x / a.slices;
  const int ktile = pair % a.k_tiles, ntile = pair / a.k_tiles;
  const int nch = min(16, a.n8 - ntile * 16);          // dC chunks of this n tile (<= 16 = 128 channels)
  const int kch = a.kt / 8;
  const uint32_t d_bytes = 16u * 2048u, a_bytes = (uint32_t)kch * 2048u, stage_bytes = d_bytes + a_bytes;
  const int64_t t0 = a.m_tiles * slice / a.slices, t1 = a.m_tiles * (slice + 1) / a.slices;
  if (warp == 4) {
    tc::tmem_alloc<256>(&s_tmem);
    if (lane == 0) {
      for (int i = 0; i < MW_STAGES; ++i) {
        tc::mbar_init(&bar_full[i], 1);
        tc::mbar_init(&bar_empty[i], 1);
      }
      tc::mbar_init(&bar_done, 1);
      tc::fence_barrier_init();
    }
  }
  // the MMA reads 16 dC chunks whatever nch is: chunks beyond the tensor's are whatever shared memory holds -- zero them once so the
  // unused accumulator rows stay finite
  for (int i = tid; i < (int)(MW_STAGES * stage_bytes / 16); i += MW_THREADS) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  tc::fence_proxy_async();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = s_tmem;
  if (warp == 5) {
    if (tc::elect_one()) {
      uint32_t it = 0;
#pragma unroll
      for (int p = 0; p < PRODUCTS; ++p)
        for (int64_t t = t0; t < t1; ++t, ++it) {
          const int s = it % MW_STAGES;
          if (it >= MW_STAGES) tc::mbar_wait(&bar_empty[s], ((it / MW_STAGES) - 1) & 1);
          unsigned char* dst = smem + (size_t)s * stage_bytes;
          tc::mbar_expect_tx(&bar_full[s], (uint32_t)nch * 2048u + a_bytes);
          tc::tma_bulk_g2s(dst, reinterpret_cast<const unsigned char*>(a.dC[p]) + ((size_t)t * a.n8 + (size_t)ntile * 16) * 2048,
                           (uint32_t)nch * 2048u, &bar_full[s]);
          tc::tma_bulk_g2s(dst + d_bytes, reinterpret_cast<const unsigned char*>(a.A[p]) + ((size_t)t * a.k8 + (size_t)ktile * kch) * 2048, a_bytes,
                           &bar_full[s]);
        }
    }
    __syncwarp();
  } else if (warp == 4) {
    if (tc::elect_one()) {
      const uint32_t idesc = tc::instr_desc_bf16(128, a.kt, 1, 1);       // both operands MN-major: K = rows of the tile
      const uint32_t hi = tc::desc_hi(2048u);                            // stride between 8-channel groups
      uint32_t it = 0;
      const int64_t n_it = (t1 - t0) * PRODUCTS;
      for (int64_t tt = 0; tt < n_it; ++tt, ++it) {
        const int s = it % MW_STAGES;
        tc::mbar_wait(&bar_full[s], (it / MW_STAGES) & 1);
        tc::fence_after_sync();
        const uint32_t sd = tc::smem_u32(smem + (size_t)s * stage_bytes), sa = sd + d_bytes;
#pragma unroll
        for (int k16 = 0; k16 < 8; ++k16) {
          const uint64_t ad = tc::desc_make(tc::desc_lo(sd + (uint32_t)k16 * 256u, 128u), hi);
          const uint64_t bd = tc::desc_make(tc::desc_lo(sa + (uint32_t)k16 * 256u, 128u), hi);
          tc::umma_bf16(tmem, ad, bd, idesc, (it | (uint32_t)k16) ? 1u : 0u);
        }
        tc::umma_commit(&bar_empty[s]);
      }
      tc::umma_commit(&bar_done);
    }
    __syncwarp();
  } else if (t1 > t0) {
    tc::mbar_wait(&bar_done, 0);
    tc::fence_after_sync();
    const int n = ntile * 128 + 32 * warp + lane;
    const uint32_t taddr = tmem + ((uint32_t)(32 * warp) << 16);
    for (int c = 0; c < kch; ++c) {
      float v[8];
      tc::tmem_ld8(taddr + 8 * c, v);
      if (n < a.n_valid) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int k = ktile * a.kt + c * 8 + j;
          if (k < a.k_valid) atomicAdd(a.dW + (size_t)n * a.ld + k, v[j]);
        }
      }
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 4) tc::tmem_dealloc<256>(tmem);
}

static int mbn_gemm_wgrad_n(howl_ctx_t* ctx, cudaStream_t st, int products, const __nv_bfloat16* const* dCs, const __nv_bfloat16* const* As, float* dW,
                            int64_t M, int N, int K, int n_valid, int k_valid, int ld);
int mbn_gemm_wgrad(howl_ctx_t* ctx, cudaStream_t st, const __nv_bfloat16* dC, const __nv_bfloat16* A, float* dW, int64_t M, int N,
                   int K, int n_valid, int k_valid, int ld) {
  return mbn_gemm_wgrad_n(ctx, st, 1, &dC, &A, dW, M, N, K, n_valid, k_valid, ld);
}
static int mbn_gemm_wgrad_n(howl_ctx_t* ctx, cudaStream_t st, int products, const __nv_bfloat16* const* dCs, const __nv_bfloat16* const* As, float* dW,
                            int64_t M, int N, int K, int n_valid, int k_valid, int ld) {
  MwArgs a;
  const __nv_bfloat16* dC = dCs[0];
  const __nv_bfloat16* A = As[0];
  for (int p = 0; p < 3; ++p) {
    a.dC[p] = dCs[p < products ? p : 0];
    a.A[p] = As[p < products ? p : 0];
  }
  a.products = products;
  a.dW = dW; a.m_tiles = mbn_tiles(M);
  const int np = mbn_pad16(N), kp = mbn_pad16(K);
  a.n8 = np / 8; a.k8 = kp / 8;
  a.kt = mbn_ntile(kp); a.k_tiles = kp / a.kt;
  a.n_tiles = (np + 127) / 128;
  a.n_valid = n_valid; a.k_valid = k_valid; a.ld = ld;
  HOWL_REQUIRE(ctx, dC && A && dW && M > 0, HOWL_E_INVALID, "mbn_wgrad: bad argument");
  const int pairs = a.n_tiles * a.k_tiles;
  int slices = (2 * ctx->sm_count + pairs - 1) / pairs;
  if (slices > a.m_tiles) slices = (int)a.m_tiles;
  if (slices < 1) slices = 1;
  a.slices = slices;
  const size_t smem = (size_t)MW_STAGES * (16 * 2048 + (size_t)(a.kt / 8) * 2048) + 2048;
  if (products == 3) {
    HOWL_CUDA(ctx, cudaFuncSetAttribute(mbn_gemm_wgrad_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    mbn_gemm_wgrad_kernel<3><<<pairs * slices, MW_THREADS, smem, st>>>(a);
    HOWL_LAUNCHED(ctx, "mbn_wgrad3");
  } else {
    HOWL_CUDA(ctx, cudaFuncSetAttribute(mbn_gemm_wgrad_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    mbn_gemm_wgrad_kernel<1><<<pairs * slices, MW_THREADS, smem, st>>>(a);
    HOWL_LAUNCHED(ctx, "mbn_wgrad");
  }
  return HOWL_OK;
}

// =============================================================================================
// fp32 C = X W^T on the tensor cores: the three bf16 products of the (hi, lo) split as ONE GEMM over a three times longer K,
//   X3 = [X_hi | X_hi | X_lo]  (TMO, mbn_pack3),   W3 = [W_hi | W_lo | W_hi]  (operand, mbn_weight_operand3),   fp32 accumulation in TMEM,
// with an fp32 row-major epilogue (+ bias, ReLU).  Used by the LSTM heads.
// =============================================================================================
__global__ void __launch_bounds__(256) mbn_pack3_kernel(const float* __restrict__ x, int64_t ld, int64_t rows, int c, int c8,
                                                        uint4* __restrict__ out) {
  const int64_t n = mbn_tiles(rows) * MBN_TILE * c8;
  for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < n; v += (int64_t)gridDim.x * blockDim.x) {
    const int r = (int)(v % MBN_TILE), chunk = (int)((v / MBN_TILE) % c8);
    const int64_t tile = v / MBN_TILE / c8, row = tile * MBN_TILE + r;
    float f[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] = 0.f;
    if (row < rows) {
      const float* src = x + row * ld + chunk * 8;
      if (chunk * 8 + 8 <= c && ((reinterpret_cast<uintptr_t>(src) & 15) == 0)) {
        const float4 a = *reinterpret_cast<const float4*>(src), b = *reinterpret_cast<const float4*>(src + 4);
        f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (chunk * 8 + j < c) f[j] = src[j];
      }
    }
    uint32_t h[4], l[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const __nv_bfloat16 h0 = __float2bfloat16_rn(f[2 * j]), h1 = __float2bfloat16_rn(f[2 * j + 1]);
      const __nv_bfloat16 l0 = __float2bfloat16_rn(f[2 * j] - __bfloat162float(h0)), l1 = __float2bfloat16_rn(f[2 * j + 1] - __bfloat162float(h1));
      h[j] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
      l[j] = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
    }
    // the output tensor has 3 * c8 chunks per tile: [hi | hi | lo]
    uint4* dst = out + (tile * 3 * c8 + chunk) * MBN_TILE + r;
    const uint4 hv = make_uint4(h[0], h[1], h[2], h[3]);
    dst[0] = hv;
    dst[(size_t)c8 * MBN_TILE] = hv;
    dst[(size_t)2 * c8 * MBN_TILE] = make_uint4(l[0], l[1], l[2], l[3]);
  }
}

int mbn_pack3(howl_ctx_t* ctx, cudaStream_t st, const float* x, int64_t ld, int64_t rows, int c, __nv_bfloat16* out) {
  HOWL_REQUIRE(ctx, x && out && rows > 0 && c > 0, HOWL_E_INVALID, "mbn_pack3: bad argument");
  const int c8 = mbn_pad16(c) / 8;
  const int64_t n = mbn_tiles(rows) * MBN_TILE * c8;
  const unsigned blocks = (unsigned)std::min<int64_t>((n + 255) / 256, (int64_t)ctx->sm_count * 16);
  mbn_pack3_kernel<<<blocks, 256, 0, st>>>(x, ld, rows, c, c8, reinterpret_cast<uint4*>(out));
  HOWL_LAUNCHED(ctx, "mbn_pack3");
  return HOWL_OK;
}

// operand of W3 = [W_hi | W_lo | W_hi] for an fp32 [n][k] matrix (transpose != 0: stored [k][n]); K3 = 3 * pad16(k)
__global__ void mbn_weight_operand3_kernel(const float* __restrict__ w, int n, int k, int ld, int transpose, int np, int kp, int nt,
                                           __nv_bfloat16* __restrict__ out) {
  const int k3 = 3 * kp, total = np * k3;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int j = i & 7, row = (i >> 3) % nt, chunk = ((i >> 3) / nt) % (k3 / 8), tile = (i >> 3) / (nt * (k3 / 8));
    const int nn = tile * nt + row, kk3 = chunk * 8 + j, part = kk3 / kp, kk = kk3 - part * kp;
    float v = 0.f;
    if (nn < n && kk < k) v = transpose ? w[(size_t)kk * ld + nn] : w[(size_t)nn * ld + kk];
    const __nv_bfloat16 hi = __float2bfloat16_rn(v);
    out[i] = part == 1 ? __float2bfloat16_rn(v - __bfloat162float(hi)) : hi;
  }
}

size_t mbn_weight_operand3_bytes(int n, int k) { return (size_t)mbn_pad16(n) * 3 * mbn_pad16(k) * 2; }

int mbn_weight_operand3(howl_ctx_t* ctx, cudaStream_t st, const float* w, int n, int k, int ld, int transpose, __nv_bfloat16* out) {
  const int np = mbn_pad16(n), kp = mbn_pad16(k), nt = mbn_ntile(np);
  const int total = np * 3 * kp;
  mbn_weight_operand3_kernel<<<(total + 255) / 256, 256, 0, st>>>(w, n, k, ld, transpose, np, kp, nt, out);
  HOWL_LAUNCHED(ctx, "mbn_weight_operand3");
  return HOWL_OK;
}

int mbn_gemm_nt3_f32(howl_ctx_t* ctx, cudaStream_t st, const __nv_bfloat16* x3, const __nv_bfloat16* wop3, float* C, int64_t ldc, int64_t M, int K,
                     int N, const float* bias, int relu, int accumulate) {
  MgArgs a;
  a.A = x3; a.W = wop3; a.add = nullptr; a.C = nullptr; a.M = M; a.m_tiles = mbn_tiles(M);
  a.C32 = C; a.ldc = ldc; a.bias = bias; a.n_valid = N; a.relu = relu; a.accumulate = accumulate;
  const int kp3 = 3 * mbn_pad16(K), np = mbn_pad16(N);
  a.k8 = kp3 / 8; a.n8 = np / 8; a.nt = mbn_ntile(np); a.n_tiles = np / a.nt;
  HOWL_REQUIRE(ctx, x3 && wop3 && C && M > 0, HOWL_E_INVALID, "mbn_gemm_nt3_f32: bad argument");
  const size_t smem = (size_t)MG_STAGES * (MG_KSTAGE * 2048 + (size_t)MG_KSTAGE * a.nt * 16);
  HOWL_CUDA(ctx, cudaFuncSetAttribute(mbn_gemm_nt_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int64_t items = a.m_tiles * a.n_tiles;
  const int grid = (int)(items < ctx->sm_count ? items : ctx->sm_count);
  mbn_gemm_nt_kernel<true><<<grid, MG_THREADS, smem, st>>>(a);
  HOWL_LAUNCHED(ctx, "mbn_gemm_f32");
  return HOWL_OK;
}

// =============================================================================================
// fp32 A^T B as three bf16 products (LSTM / LAS weight gradients): split-pack + three weight-gradient GEMMs
// =============================================================================================
// thread = (row of a 128-row tile, 8-channel chunk = blockIdx.y), walking the tiles: stores are contiguous, loads are 32-byte row pieces.
// colsum != null: the column sums of x (bias gradients) are accumulated on the way (+= into colsum and colsum2).
__global__ void __launch_bounds__(256) mbn_pack_split_kernel(const float* __restrict__ x, int64_t ld, int64_t rows, int c, int c8,
                                                             uint4* __restrict__ hi, uint4* __restrict__ lo, float* __restrict__ colsum,
                                                             float* __restrict__ colsum2) {
  const int chunk = blockIdx.y, r = threadIdx.x & (MBN_TILE - 1);
  const int64_t tiles = mbn_tiles(rows);
  float cs[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) cs[j] = 0.f;
  for (int64_t tile = (int64_t)blockIdx.x * 2 + (threadIdx.x >> 7); tile < tiles; tile += (int64_t)gridDim.x * 2) {
    const int64_t row = tile * MBN_TILE + r;
    float f[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] = 0.f;
    if (row < rows) {
      const float* src = x + row * ld + chunk * 8;
      if (chunk * 8 + 8 <= c && ((reinterpret_cast<uintptr_t>(src) & 15) == 0)) {
        const float4 a = *reinterpret_cast<const float4*>(src), b = *reinterpret_cast<const float4*>(src + 4);
        f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (chunk * 8 + j < c) f[j] = src[j];
      }
    }
    uint32_t h[4], l[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const __nv_bfloat16 h0 = __float2bfloat16_rn(f[2 * j]), h1 = __float2bfloat16_rn(f[2 * j + 1]);
      const __nv_bfloat16 l0 = __float2bfloat16_rn(f[2 * j] - __bfloat162float(h0)), l1 = __float2bfloat16_rn(f[2 * j + 1] - __bfloat162float(h1));
      h[j] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
      l[j] = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
    }
    const size_t v = ((size_t)tile * c8 + chunk) * MBN_TILE + r;
    hi[v] = make_uint4(h[0], h[1], h[2], h[3]);
    lo[v] = make_uint4(l[0], l[1], l[2], l[3]);
#pragma unroll
    for (int j = 0; j < 8; ++j) cs[j] += f[j];
  }
  if (colsum) {
    __shared__ float s_cs[8][8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float t = warp_sum(cs[j]);
      if (lane == 0) s_cs[warp][j] = t;
    }
    __syncthreads();
    if (threadIdx.x < 8 && chunk * 8 + threadIdx.x < c) {
      float t = 0.f;
      for (int w = 0; w < 8; ++w) t += s_cs[w][threadIdx.x];
      atomicAdd(colsum + chunk * 8 + threadIdx.x, t);
      if (colsum2) atomicAdd(colsum2 + chunk * 8 + threadIdx.x, t);
    }
  }
}

int mbn_pack_split(howl_ctx_t* ctx, cudaStream_t st, const float* x, int64_t ld, int64_t rows, int c, __nv_bfloat16* hi, __nv_bfloat16* lo,
                   float* colsum, float* colsum2) {
  HOWL_REQUIRE(ctx, x && hi && lo && rows > 0 && c > 0, HOWL_E_INVALID, "mbn_pack_split: bad argument");
  const int c8 = mbn_pad16(c) / 8;
  const int64_t pairs = (mbn_tiles(rows) + 1) / 2;
  const unsigned gx = (unsigned)std::max<int64_t>(1, std::min<int64_t>(pairs, std::max<int64_t>(1, (int64_t)ctx->sm_count * 16 / c8)));
  mbn_pack_split_kernel<<<dim3(gx, c8), 256, 0, st>>>(x, ld, rows, c, c8, reinterpret_cast<uint4*>(hi), reinterpret_cast<uint4*>(lo), colsum,
                                                      colsum2);
  HOWL_LAUNCHED(ctx, "mbn_pack_split");
  return HOWL_OK;
}

int mbn_atb3_packed(howl_ctx_t* ctx, cudaStream_t st, const __nv_bfloat16* xhi, const __nv_bfloat16* xlo, const __nv_bfloat16* yhi,
                    const __nv_bfloat16* ylo, float* dW, int64_t rows, int N, int K, int ld) {
  // one launch: the three operand pairs stream through the same accumulator tile, one epilogue
  const __nv_bfloat16* xs[3] = {xhi, xhi, xlo};
  const __nv_bfloat16* ys[3] = {yhi, ylo, yhi};
  return mbn_gemm_wgrad_n(ctx, st, 3, xs, ys, dW, rows, N, K, N, K, ld);
}

// =============================================================================================
// test hooks (include/howl_b200_debug.h): the two GEMMs on plain fp32 row-major matrices (packed to the operand format inside)
// =============================================================================================
__global__ void mbn_pack_kernel(const float* __restrict__ x, int64_t rows, int c, int cp, __nv_bfloat16* __restrict__ out) {
  const int64_t rows_pad = mbn_tiles(rows) * MBN_TILE, n = rows_pad * cp;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int j = (int)(i & 7);
    const int64_t v = i >> 3;
    const int r = (int)(v % MBN_TILE), chunk = (int)((v / MBN_TILE) % (cp / 8));
    const int64_t row = (v / MBN_TILE / (cp / 8)) * MBN_TILE + r;
    const int ch = chunk * 8 + j;
    out[i] = __float2bfloat16_rn((row < rows && ch < c) ? x[row * c + ch] : 0.f);
  }
}
__global__ void mbn_unpack_kernel(const __nv_bfloat16* __restrict__ t, int64_t rows, int c, int cp, float* __restrict__ out) {
  const int64_t n = rows * c;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = i / c;
    const int ch = (int)(i - row * c);
    out[i] = __bfloat162float(t[(mbn_vec(row, ch >> 3, cp / 8) << 3) + (ch & 7)]);
  }
}

extern "C" int64_t howl_b200_debug_mbn_workspace_bytes(int64_t M, int K, int N) {
  return (int64_t)(mbn_tmo_bytes(M, K) + 2 * mbn_tmo_bytes(M, N) + mbn_weight_operand_bytes(N, K) + 1024);
}

extern "C" int howl_b200_debug_mbn_gemm(howl_ctx_t* ctx, void* stream, const float* A, const float* W, const float* add, float* C, int64_t M,
                                        int32_t K, int32_t N, void* workspace, size_t workspace_bytes) {
  if (!ctx) return HOWL_E_INVALID;
  HOWL_REQUIRE(ctx, A && W && C && workspace && (int64_t)workspace_bytes >= howl_b200_debug_mbn_workspace_bytes(M, K, N), HOWL_E_INVALID,
               "debug_mbn_gemm: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  char* p = (char*)workspace;
  __nv_bfloat16* a_t = (__nv_bfloat16*)p; p += howl_align_up(mbn_tmo_bytes(M, K), 256);
  __nv_bfloat16* c_t = (__nv_bfloat16*)p; p += howl_align_up(mbn_tmo_bytes(M, N), 256);
  __nv_bfloat16* add_t = (__nv_bfloat16*)p; p += howl_align_up(mbn_tmo_bytes(M, N), 256);
  __nv_bfloat16* w_t = (__nv_bfloat16*)p;
  mbn_pack_kernel<<<ctx->sm_count * 4, 256, 0, st>>>(A, M, K, mbn_pad16(K), a_t);
  HOWL_LAUNCHED(ctx, "mbn_pack");
  if (add) {
    mbn_pack_kernel<<<ctx->sm_count * 4, 256, 0, st>>>(add, M, N, mbn_pad16(N), add_t);
    HOWL_LAUNCHED(ctx, "mbn_pack");
  }
  int rc = mbn_weight_operand(ctx, st, W, N, K, K, 0, w_t);
  if (rc) return rc;
  rc = mbn_gemm_nt(ctx, st, a_t, w_t, add ? add_t : nullptr, c_t, M, K, N);
  if (rc) return rc;
  mbn_unpack_kernel<<<ctx->sm_count * 4, 256, 0, st>>>(c_t, M, N, mbn_pad16(N), C);
  HOWL_LAUNCHED(ctx, "mbn_unpack");
  return HOWL_OK;
}

extern "C" int howl_b200_debug_mbn_wgrad(howl_ctx_t* ctx, void* stream, const float* dC, const float* A, float* dW, int64_t M, int32_t N,
                                         int32_t K, void* workspace, size_t workspace_bytes) {
  if (!ctx) return HOWL_E_INVALID;
  HOWL_REQUIRE(ctx, dC && A && dW && workspace && (int64_t)workspace_bytes >= howl_b200_debug_mbn_workspace_bytes(M, K, N), HOWL_E_INVALID,
               "debug_mbn_wgrad: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  char* p = (char*)workspace;
  __nv_bfloat16* a_t = (__nv_bfloat16*)p; p += howl_align_up(mbn_tmo_bytes(M, K), 256);
  __nv_bfloat16* c_t = (__nv_bfloat16*)p;
  mbn_pack_kernel<<<ctx->sm_count * 4, 256, 0, st>>>(A, M, K, mbn_pad16(K), a_t);
  HOWL_LAUNCHED(ctx, "mbn_pack");
  mbn_pack_kernel<<<ctx->sm_count * 4, 256, 0, st>>>(dC, M, N, mbn_pad16(N), c_t);
  HOWL_LAUNCHED(ctx, "mbn_pack");
  HOWL_CUDA(ctx, cudaMemsetAsync(dW, 0, sizeof(float) * (size_t)N * K, st));
  return mbn_gemm_wgrad(ctx, st, c_t, a_t, dW, M, N, K, N, K, K);
}
