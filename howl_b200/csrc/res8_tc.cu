// Res8 45->45 3x3 convolutions on the 5th-generation tensor cores (tcgen05.mma + TMEM), sm_100a.
//
// Forward / data gradient as an implicit GEMM per utterance group:
//     D[pixel, out] += sum_{tap} A_tap[pixel, in] * W_tap[in, out]            M = 128-pixel tiles, N = 48, K = 9 x 48
// The (normalised) input tile of U stacked utterances lives in shared memory in the padded "pitch 11" raster
// (one shared zero column between image rows), so that a 3x3 tap is a shift of the operand start address:
//     q(u, y, x) = u * (H + 2) * 11 + (y + 1) * 11 + (x + 1),   shift(dy, dx) = (dy - 1) * 11 + (dx - 1).
// Weight gradient as 9 GEMMs  dW_tap[out, in] += sum_{pixel} dC[pixel, out] * X[pixel + shift, in]  with both operands
// MN-major (K = pixels) from the same shared-memory layout; the 9 x 48 accumulator columns stay in TMEM across all
// utterances a CTA owns.  Operands are bf16 (hi, lo) splits of the fp32 tensors, three MMAs per product, fp32
// accumulation (tc_common.cuh).  BatchNorm of the producer layer is applied while staging, ReLU / residual /
// statistics in the epilogue straight out of TMEM -- same dataflow and HBM tensors as the fp32 kernels in res8.cu.
#include "res8_common.cuh"
#include "tc_common.cuh"

#define TC_THREADS 256
#define TC_PITCH 11
#define TC_Q0 12                  // raster index of pixel (0, 0)
#define TC_N 48                   // channels padded to 48
#define TC_WBYTES (R8TC_WBLOCK * 2)

struct TcGeom {
  int H, U, Pu, tiles, R;        // R = operand rows kept in shared memory
};

static TcGeom tc_geom(int H) {
  TcGeom best{H, 0, (H + 2) * TC_PITCH, 0, 0};
  double best_eff = 0.0;
  for (int U = 1; U <= 4; ++U) {
    const int n_out = (U - 1) * best.Pu + TC_PITCH * H - 1;
    const int tiles = (n_out + 127) / 128;
    const int R = (TC_Q0 + tiles * 128 + 12 + 7) & ~7;
    const size_t smem = 2 * (size_t)TC_WBYTES + 2 * (size_t)R * 96 + 4096;
    if (tiles > 5 || smem > 227 * 1024) continue;
    const double eff = (double)U * H * R8_W / (tiles * 128.0);
    if (eff > best_eff + 1e-9) {
      best_eff = eff;
      best.U = U;
      best.tiles = tiles;
      best.R = R;
    }
  }
  return best;
}

static size_t tc_wgrad_smem(int H, int* Kp_out, int* Rx_out) {
  const int Pu = (H + 2) * TC_PITCH;
  const int Kp = (Pu + 15) & ~15;
  const int Rx = (12 + Kp + 12 + 7) & ~7;
  if (Kp_out) *Kp_out = Kp;
  if (Rx_out) *Rx_out = Rx;
  // the A operand addresses 16 groups of 8 output channels (M = 128): keep the 10 padding groups inside the buffer
  const size_t operand = (12 * (size_t)Kp + 12 * (size_t)Rx) * 16;
  const size_t need_a = ((size_t)(6 + 16) * Kp) * 16;
  return (operand > need_a ? operand : need_a) + 200 * 4;
}

bool r8tc_supported(int H) { return tc_geom(H).U > 0 && tc_wgrad_smem(H, nullptr, nullptr) <= 227 * 1024; }

// =============================================================================================
// weight operands: W fp32 [6][45 o][45 c][3][3] -> bf16 (hi, lo) [tap][chunk][48 n][8 k]
//   dir 0 (forward)      : n = o, k = c, tap as stored
//   dir 1 (data gradient): n = c, k = o, tap flipped (8 - tap)
// =============================================================================================
__global__ void tc_weight_prep_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ out, int dir) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= R8_LAYERS * R8TC_WBLOCK) return;
  const int layer = i / R8TC_WBLOCK, r = i - layer * R8TC_WBLOCK;
  const int j = r & 7, n = (r >> 3) % TC_N, chunk = ((r >> 3) / TC_N) % 6, tap = (r >> 3) / (TC_N * 6);
  const int k = chunk * 8 + j;
  float v = 0.f;
  if (n < R8_C && k < R8_C) {
    const int o = dir ? k : n, c = dir ? n : k, t = dir ? 8 - tap : tap;
    v = w[(size_t)layer * R8_KW + (o * R8_C + c) * 9 + t];
  }
  const __nv_bfloat16 hi = __float2bfloat16_rn(v);
  const __nv_bfloat16 lo = __float2bfloat16_rn(v - __bfloat162float(hi));
  __nv_bfloat16* base = out + ((size_t)(layer * 2 + dir) * 2) * R8TC_WBLOCK;
  base[r] = hi;
  base[R8TC_WBLOCK + r] = lo;
}

int r8tc_weight_prep(howl_ctx_t* ctx, cudaStream_t st, const float* w_layers, __nv_bfloat16* wprep, int dir) {
  const int n = R8_LAYERS * R8TC_WBLOCK;
  tc_weight_prep_kernel<<<(n + 255) / 256, 256, 0, st>>>(w_layers, wprep, dir);
  HOWL_LAUNCHED(ctx, "tc_weight_prep");
  return HOWL_OK;
}

// =============================================================================================
// forward / data-gradient kernel
// =============================================================================================
struct TcConvArgs {
  ConvParams p;
  const __nv_bfloat16* whi;
  const __nv_bfloat16* wlo;
  TcGeom g;
};

// stage one [45][H][10] fp32 plane set as (hi, lo) bf16 rows of the raster; BN applied on the fly
__device__ __forceinline__ void tc_stage_planes(const float* __restrict__ src, bool present, int H, int row_base,
                                                int R, const float* s_mean, const float* s_rstd, uint4* a_hi,
                                                uint4* a_lo, int tid) {
  const int HW = H * R8_W;
  for (int it = tid; it < 6 * HW; it += TC_THREADS) {
    const int chunk = it / HW, pp = it - chunk * HW;
    const int y = pp / R8_W, x = pp - y * R8_W;
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = chunk * 8 + j;
      v[j] = (present && c < R8_C) ? (__ldg(src + (size_t)c * HW + pp) - s_mean[c]) * s_rstd[c] : 0.f;
    }
    uint4 hi, lo;
    tc::split8(v, hi, lo);
    const int row = row_base + (y + 1) * TC_PITCH + (x + 1);
    a_hi[chunk * R + row] = hi;
    a_lo[chunk * R + row] = lo;
  }
}

template <bool RELU, int STATS>
__global__ void __launch_bounds__(TC_THREADS, 1) conv3x3_tc_kernel(const TcConvArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  const ConvParams& p = a.p;
  const int H = p.H, U = a.g.U, Pu = a.g.Pu, tiles = a.g.tiles, R = a.g.R, HW = H * R8_W;
  uint4* w_hi = reinterpret_cast<uint4*>(smem);
  uint4* w_lo = reinterpret_cast<uint4*>(smem + TC_WBYTES);
  uint4* a_hi = reinterpret_cast<uint4*>(smem + 2 * TC_WBYTES);
  uint4* a_lo = a_hi + 6 * R;
  float* s_f = reinterpret_cast<float*>(a_lo + 6 * R);
  float* s_mean = s_f;            // [48]
  float* s_rstd = s_f + 48;
  float* s_amean = s_f + 96;
  float* s_arstd = s_f + 144;
  float* s_red = s_f + 192;       // [8 warps][2][48]
  __shared__ __align__(8) uint64_t bar_w, bar_mma;
  __shared__ uint32_t s_tmem;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  if (warp == 0) tc::tmem_alloc<256>(&s_tmem);
  if (tid == 32) {
    tc::mbar_init(&bar_w, 1);
    tc::mbar_init(&bar_mma, 1);
    tc::fence_barrier_init();
  }
  for (int i = tid; i < 12 * R; i += TC_THREADS) a_hi[i] = make_uint4(0, 0, 0, 0);   // a_hi and a_lo are contiguous
  if (tid < 48) {
    const bool c_ok = tid < R8_C;
    s_mean[tid] = (c_ok && p.in_mean) ? p.in_mean[tid] : 0.f;
    s_rstd[tid] = (c_ok && p.in_rstd) ? p.in_rstd[tid] : 1.f;
    s_amean[tid] = (c_ok && STATS == 2) ? p.aux_mean[tid] : 0.f;
    s_arstd[tid] = (c_ok && STATS == 2) ? p.aux_rstd[tid] : 0.f;
  }
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  if (tid == 32) {
    tc::mbar_expect_tx(&bar_w, 2 * TC_WBYTES);
    tc::tma_bulk_g2s(w_hi, a.whi, TC_WBYTES, &bar_w);
    tc::tma_bulk_g2s(w_lo, a.wlo, TC_WBYTES, &bar_w);
  }
  const uint32_t tmem = s_tmem;
  const uint32_t idesc = tc::instr_desc_bf16(128, TC_N, 0, 0);
  const uint32_t a_hi_s = tc::smem_u32(a_hi), a_lo_s = tc::smem_u32(a_lo);
  const uint32_t w_hi_s = tc::smem_u32(w_hi), w_lo_s = tc::smem_u32(w_lo);
  tc::mbar_wait(&bar_w, 0);

  float st1[R8_C], st2[R8_C];
#pragma unroll
  for (int c = 0; c < R8_C; ++c) st1[c] = st2[c] = 0.f;

  const int64_t groups = (p.B + U - 1) / U;
  uint32_t phase = 0;
  for (int64_t g = blockIdx.x; g < groups; g += gridDim.x) {
    // ---- stage U utterances (generic proxy writes), then hand the tile to the async proxy
    for (int u = 0; u < U; ++u) {
      const int64_t b = g * U + u;
      tc_stage_planes(p.in + b * (int64_t)R8_C * HW, b < p.B, H, u * Pu, R, s_mean, s_rstd, a_hi, a_lo, tid);
    }
    tc::fence_proxy_async();
    __syncthreads();
    // ---- one thread issues every MMA of this group: tiles x 9 taps x 3 k-steps x 3 split terms
    if (tid == 0) {
      tc::fence_after_sync();
      for (int t = 0; t < tiles; ++t) {
        const uint32_t d = tmem + (uint32_t)(t * TC_N);
        uint32_t acc = 0;
#pragma unroll 1
        for (int tap = 0; tap < 9; ++tap) {
          const int shift = (tap / 3 - 1) * TC_PITCH + (tap % 3 - 1);
          const uint32_t row0 = (uint32_t)(TC_Q0 + 128 * t + shift);
#pragma unroll
          for (int ks = 0; ks < 3; ++ks) {
            const uint32_t aoff = ((uint32_t)(2 * ks) * R + row0) * 16u;
            const uint32_t boff = (uint32_t)((tap * 6 + 2 * ks) * TC_N) * 16u;
            const uint64_t ah = tc::smem_desc(a_hi_s + aoff, (uint32_t)R * 16u, 128u);
            const uint64_t al = tc::smem_desc(a_lo_s + aoff, (uint32_t)R * 16u, 128u);
            const uint64_t bh = tc::smem_desc(w_hi_s + boff, TC_N * 16u, 128u);
            const uint64_t bl = tc::smem_desc(w_lo_s + boff, TC_N * 16u, 128u);
            tc::umma_bf16(d, al, bh, idesc, acc);
            tc::umma_bf16(d, ah, bl, idesc, 1u);
            tc::umma_bf16(d, ah, bh, idesc, 1u);
            acc = 1u;
          }
        }
      }
      tc::umma_commit(&bar_mma);
    }
    tc::mbar_wait(&bar_mma, phase);
    phase ^= 1u;
    tc::fence_after_sync();
    // ---- epilogue out of TMEM: warps 0-3 take even tiles, warps 4-7 odd tiles; thread = one raster row
    for (int t = warp >> 2; t < tiles; t += 2) {
      const int q = TC_Q0 + 128 * t + 32 * (warp & 3) + lane;
      const int u = q / Pu, rem = q - u * Pu;
      const int y = rem / TC_PITCH - 1, x = rem % TC_PITCH - 1;
      const int64_t b = g * U + u;
      const bool valid = (u < U) && (b < p.B) && (y >= 0) && (y < H) && (x >= 0) && (x < R8_W);
      const uint32_t taddr = tmem + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)(t * TC_N);
      float v[48];
      tc::tmem_ld16(taddr, v);
      tc::tmem_ld16(taddr + 16, v + 16);
      tc::tmem_ld16(taddr + 32, v + 32);
      if (valid) {
        const int64_t base = b * (int64_t)R8_C * HW + y * R8_W + x;
#pragma unroll
        for (int c = 0; c < R8_C; ++c) {
          float o = v[c];
          if (RELU) o = fmaxf(o, 0.f);
          const int64_t idx = base + (int64_t)c * HW;
          if (p.res) o += __ldg(p.res + idx);
          p.out[idx] = o;
          if (STATS == 1) {
            st1[c] += o;
            st2[c] = fmaf(o, o, st2[c]);
          } else if (STATS == 2) {
            st1[c] += o;
            st2[c] = fmaf(o, (__ldg(p.aux + idx) - s_amean[c]) * s_arstd[c], st2[c]);
          }
        }
      }
    }
    tc::fence_before_sync();
    __syncthreads();
  }
  // ---- per-channel statistics: lanes -> warps -> one fp64 atomic per channel and CTA
  if (STATS) {
#pragma unroll
    for (int c = 0; c < R8_C; ++c) {
      const float a1 = warp_sum(st1[c]), a2 = warp_sum(st2[c]);
      if (lane == 0) {
        s_red[(warp * 2 + 0) * 48 + c] = a1;
        s_red[(warp * 2 + 1) * 48 + c] = a2;
      }
    }
    __syncthreads();
    if (tid < 2 * R8_C) {
      const int which = tid / R8_C, c = tid - which * R8_C;
      double s = 0.0;
      for (int w = 0; w < 8; ++w) s += (double)s_red[(w * 2 + which) * 48 + c];
      atomicAdd(&p.stats[tid], s);
    }
  }
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc<256>(tmem);
}

static size_t tc_conv_smem(const TcGeom& g) { return 2 * (size_t)TC_WBYTES + 12 * (size_t)g.R * 16 + (192 + 8 * 2 * 48) * 4; }

int r8tc_conv(howl_ctx_t* ctx, cudaStream_t st, const ConvParams& p, const __nv_bfloat16* whi, const __nv_bfloat16* wlo,
              bool relu, int stats) {
  TcConvArgs a;
  a.p = p;
  a.whi = whi;
  a.wlo = wlo;
  a.g = tc_geom(p.H);
  HOWL_REQUIRE(ctx, a.g.U > 0, HOWL_E_UNSUPPORTED, "tensor-core conv: H=%d does not fit", p.H);
  const size_t smem = tc_conv_smem(a.g);
  const int64_t groups = (p.B + a.g.U - 1) / a.g.U;
  const int grid = (int)(groups < ctx->sm_count ? groups : ctx->sm_count);
#define TC_LAUNCH(RELU_, STATS_)                                                                                  \
  do {                                                                                                            \
    HOWL_CUDA(ctx, cudaFuncSetAttribute(conv3x3_tc_kernel<RELU_, STATS_>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                        (int)smem));                                                              \
    conv3x3_tc_kernel<RELU_, STATS_><<<grid, TC_THREADS, smem, st>>>(a);                                          \
  } while (0)
  if (relu && stats == 1) TC_LAUNCH(true, 1);
  else if (relu && stats == 0) TC_LAUNCH(true, 0);
  else if (!relu && stats == 2) TC_LAUNCH(false, 2);
  else if (!relu && stats == 0) TC_LAUNCH(false, 0);
  else HOWL_REQUIRE(ctx, false, HOWL_E_INVALID, "tensor-core conv: unsupported mode");
#undef TC_LAUNCH
  HOWL_LAUNCHED(ctx, relu ? "conv3x3_fwd_tc" : "conv3x3_dgrad_tc");
  return HOWL_OK;
}

// =============================================================================================
// weight-gradient kernel: dW_tap[o][c] = sum_q dC[q][o] * X[q + shift][c], all 9 taps resident in TMEM
// =============================================================================================
struct TcWgradArgs {
  WgradParams p;
  int Kp;    // raster positions per utterance rounded up to 16 (the GEMM K extent)
  int Rx;    // rows of the X operand: 12 + Kp + 12 (+ pad)
};

__global__ void __launch_bounds__(TC_THREADS, 1) conv3x3_wgrad_tc_kernel(const TcWgradArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  const WgradParams& p = a.p;
  const int H = p.H, HW = H * R8_W, Kp = a.Kp, Rx = a.Rx;
  uint4* d_hi = reinterpret_cast<uint4*>(smem);     // [6][Kp]   dC, rows = raster position q
  uint4* d_lo = d_hi + 6 * Kp;
  uint4* x_hi = d_lo + 6 * Kp;                      // [6][Rx]   X,  rows = q + 12
  uint4* x_lo = x_hi + 6 * Rx;
  float* s_mean = reinterpret_cast<float*>(x_lo + 6 * Rx);
  float* s_rstd = s_mean + 48;
  __shared__ __align__(8) uint64_t bar_mma;
  __shared__ uint32_t s_tmem;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  if (warp == 0) tc::tmem_alloc<512>(&s_tmem);
  if (tid == 32) {
    tc::mbar_init(&bar_mma, 1);
    tc::fence_barrier_init();
  }
  for (int i = tid; i < 12 * Kp + 12 * Rx; i += TC_THREADS) d_hi[i] = make_uint4(0, 0, 0, 0);
  if (tid < 48) {
    const bool c_ok = tid < R8_C;
    s_mean[tid] = (c_ok && p.x_mean) ? p.x_mean[tid] : 0.f;
    s_rstd[tid] = (c_ok && p.x_rstd) ? p.x_rstd[tid] : 1.f;
  }
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = s_tmem;
  const uint32_t idesc = tc::instr_desc_bf16(128, TC_N, 1, 1);   // both operands MN-major (K = raster positions)
  const uint32_t d_hi_s = tc::smem_u32(d_hi), d_lo_s = tc::smem_u32(d_lo);
  const uint32_t x_hi_s = tc::smem_u32(x_hi), x_lo_s = tc::smem_u32(x_lo);
  uint32_t phase = 0, first = 1;
  for (int64_t b = blockIdx.x; b < p.B; b += gridDim.x) {
    // dC: no normalisation, rows q;  X: BN on the fly, rows q + 12
    {
      const float* src = p.dc + b * (int64_t)R8_C * HW;
      for (int it = tid; it < 6 * HW; it += TC_THREADS) {
        const int chunk = it / HW, pp = it - chunk * HW;
        const int y = pp / R8_W, x = pp - y * R8_W;
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int c = chunk * 8 + j;
          v[j] = (c < R8_C) ? __ldg(src + (size_t)c * HW + pp) : 0.f;
        }
        uint4 hi, lo;
        tc::split8(v, hi, lo);
        const int row = (y + 1) * TC_PITCH + (x + 1);
        d_hi[chunk * Kp + row] = hi;
        d_lo[chunk * Kp + row] = lo;
      }
    }
    tc_stage_planes(p.x + b * (int64_t)R8_C * HW, true, H, 12, Rx, s_mean, s_rstd, x_hi, x_lo, tid);
    tc::fence_proxy_async();
    __syncthreads();
    if (tid == 0) {
      tc::fence_after_sync();
#pragma unroll 1
      for (int tap = 0; tap < 9; ++tap) {
        const int shift = (tap / 3 - 1) * TC_PITCH + (tap % 3 - 1);
        const uint32_t d = tmem + (uint32_t)(tap * TC_N);
        uint32_t acc = first ? 0u : 1u;
        for (int k0 = 0; k0 < Kp; k0 += 16) {
          const uint32_t aoff = (uint32_t)k0 * 16u;
          const uint32_t boff = (uint32_t)(12 + shift + k0) * 16u;
          // MN-major: lbo = stride between 8-position K groups (128 B), sbo = stride between 8-channel groups
          const uint64_t ah = tc::smem_desc(d_hi_s + aoff, 128u, (uint32_t)Kp * 16u);
          const uint64_t al = tc::smem_desc(d_lo_s + aoff, 128u, (uint32_t)Kp * 16u);
          const uint64_t bh = tc::smem_desc(x_hi_s + boff, 128u, (uint32_t)Rx * 16u);
          const uint64_t bl = tc::smem_desc(x_lo_s + boff, 128u, (uint32_t)Rx * 16u);
          tc::umma_bf16(d, al, bh, idesc, acc);
          tc::umma_bf16(d, ah, bl, idesc, 1u);
          tc::umma_bf16(d, ah, bh, idesc, 1u);
          acc = 1u;
        }
      }
      tc::umma_commit(&bar_mma);
    }
    first = 0;
    tc::mbar_wait(&bar_mma, phase);   // operands may be overwritten once the MMAs have drained
    phase ^= 1u;
    tc::fence_after_sync();
  }
  // ---- epilogue: TMEM lane = output channel o (rows 45..127 are padding), column = tap * 48 + c
  if (first == 0) {
    const int o = 32 * (warp & 3) + lane;
    const int half = warp >> 2;                    // warps 0-3: taps 0..4, warps 4-7: taps 5..8
    for (int tap = half ? 5 : 0; tap < (half ? 9 : 5); ++tap) {
      const uint32_t taddr = tmem + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)(tap * TC_N);
      float v[48];
      tc::tmem_ld16(taddr, v);
      tc::tmem_ld16(taddr + 16, v + 16);
      tc::tmem_ld16(taddr + 32, v + 32);
      if (o < R8_C) {
#pragma unroll
        for (int c = 0; c < R8_C; ++c) atomicAdd(&p.dw[(o * R8_C + c) * 9 + tap], v[c]);
      }
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc<512>(tmem);
}

int r8tc_wgrad(howl_ctx_t* ctx, cudaStream_t st, const WgradParams& p) {
  TcWgradArgs a;
  a.p = p;
  const size_t smem = tc_wgrad_smem(p.H, &a.Kp, &a.Rx);
  HOWL_REQUIRE(ctx, smem <= 227 * 1024, HOWL_E_UNSUPPORTED, "tensor-core wgrad: H=%d does not fit", p.H);
  HOWL_CUDA(ctx, cudaFuncSetAttribute(conv3x3_wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int grid = (int)(p.B < ctx->sm_count ? p.B : ctx->sm_count);
  conv3x3_wgrad_tc_kernel<<<grid, TC_THREADS, smem, st>>>(a);
  HOWL_LAUNCHED(ctx, "conv3x3_wgrad_tc");
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
  if (!mn_major) {
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
