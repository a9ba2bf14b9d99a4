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

#define TC_THREADS 288            // weight-gradient kernel: warps 0-7 staging transform + epilogue, warp 8 TMA + MMA issue
#define TC_CONV_THREADS 256       // forward / data-gradient kernel: 8 worker warps, lane 0 of warp 0 also issues TMA + MMAs
#define TC_WORKERS 256
#define TC_PITCH 11
#define TC_Q0 12                  // raster index of pixel (0, 0)
#define TC_N 48                   // channels padded to 48
#define TC_WBYTES (R8TC_WBLOCK * 2)
#define TC_SMEM_LIMIT (227 * 1024 - 512)

struct TcGeom {
  int H, U, Pu, tiles, R;        // R = operand rows kept in shared memory
  size_t smem;
};

// staging slot of one utterance's raw planes: TMA needs 16-byte aligned source and size, an utterance block starts at a
// multiple of 8 bytes only, so the copy starts at the aligned-down address and the slot carries up to 16 bytes of slack
__host__ __device__ static inline size_t tc_slot_bytes(int H) { return ((size_t)R8_C * H * R8_W * 4 + 16 + 15) & ~(size_t)15; }

static size_t tc_conv_smem(int U, int R, int H) {
  return 2 * (size_t)TC_WBYTES + 12 * (size_t)R * 16 + (size_t)U * tc_slot_bytes(H) + (192 + 8 * 2 * 48) * 4;
}

static TcGeom tc_geom(int H) {
  TcGeom best{H, 0, (H + 2) * TC_PITCH, 0, 0, 0};
  double best_eff = 0.0;
  for (int U = 1; U <= 4; ++U) {
    const int n_out = (U - 1) * best.Pu + TC_PITCH * H - 1;
    const int tiles = (n_out + 127) / 128;
    const int R = (TC_Q0 + tiles * 128 + 12 + 7) & ~7;
    const size_t smem = tc_conv_smem(U, R, H);
    if (tiles > 5 || smem > TC_SMEM_LIMIT) continue;
    const double eff = (double)U * H * R8_W / (tiles * 128.0);
    if (eff > best_eff + 1e-9) {
      best_eff = eff;
      best.U = U;
      best.tiles = tiles;
      best.R = R;
      best.smem = smem;
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
  // the A operand addresses 16 groups of 8 rows (M = 128 = dC_hi | dC_lo | padding): groups 12..15 must stay inside
  const size_t operand = (12 * (size_t)Kp + 12 * (size_t)Rx) * 16;
  const size_t need_a = ((size_t)16 * Kp) * 16;
  return (operand > need_a ? operand : need_a) + tc_slot_bytes(H) + 128 * 4;
}

static size_t tc_fwd_op_smem(int H) {
  const int Kp = ((H + 2) * TC_PITCH + 15) & ~15;
  return 2 * (size_t)TC_WBYTES + 2 * (size_t)12 * (Kp + 24) * 16 + 128 * 16 + (48 + 8 * 2 * 24) * 4 + 12 * 16;
}
bool r8tc_supported(int H) {
  return tc_geom(H).U > 0 && tc_wgrad_smem(H, nullptr, nullptr) <= TC_SMEM_LIMIT && tc_fwd_op_smem(H) <= TC_SMEM_LIMIT &&
         TC_PITCH * H - 1 <= 3 * 128;
}

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

// raw fp32 planes [45][H][10] (shared-memory staging, landed by TMA) -> (hi, lo) bf16 raster rows; BN on the fly.
// 256 worker threads; one item = 8 channels of one pixel = one 16-byte operand row.
__device__ __forceinline__ void tc_transform(const float* __restrict__ stage, bool present, int H, int row_base, int R,
                                             const float* s_mean, const float* s_rstd, uint4* a_hi, uint4* a_lo, int tid) {
  const int HW = H * R8_W;
#pragma unroll 1
  for (int chunk = 0; chunk < 6; ++chunk) {
    float mu[8], rs[8];
    int coff[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = chunk * 8 + j;
      const int cc = c < R8_C ? c : R8_C - 1;                 // clamped address, zero scale: no guards around loads
      mu[j] = s_mean[cc];
      rs[j] = (present && c < R8_C) ? s_rstd[cc] : 0.f;
      coff[j] = cc * HW;
    }
    for (int pp = tid; pp < HW; pp += TC_WORKERS) {
      const int y = pp / R8_W, x = pp - y * R8_W;
      float v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = (stage[coff[j] + pp] - mu[j]) * rs[j];
      uint4 hi, lo;
      tc::split8(v, hi, lo);
      const int row = row_base + (y + 1) * TC_PITCH + (x + 1);
      a_hi[chunk * R + row] = hi;
      a_lo[chunk * R + row] = lo;
    }
  }
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

template <bool RELU, int STATS>
__global__ void __launch_bounds__(TC_CONV_THREADS, 1) conv3x3_tc_kernel(const TcConvArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  const ConvParams& p = a.p;
  const int H = p.H, U = a.g.U, Pu = a.g.Pu, tiles = a.g.tiles, R = a.g.R, HW = H * R8_W;
  const uint32_t plane_bytes = (uint32_t)(R8_C * HW * 4);
  uint4* w_hi = reinterpret_cast<uint4*>(smem);
  uint4* w_lo = reinterpret_cast<uint4*>(smem + TC_WBYTES);
  uint4* a_hi = reinterpret_cast<uint4*>(smem + 2 * TC_WBYTES);
  uint4* a_lo = a_hi + 6 * R;
  unsigned char* stage = reinterpret_cast<unsigned char*>(a_lo + 6 * R);   // [U] slots of raw fp32 planes [45][HW]
  const size_t slot = tc_slot_bytes(H);
  float* s_f = reinterpret_cast<float*>(stage + (size_t)U * slot);
  float* s_mean = s_f;            // [48]
  float* s_rstd = s_f + 48;
  float* s_amean = s_f + 96;
  float* s_arstd = s_f + 144;
  float* s_red = s_f + 192;       // [8 warps][2][48]
  __shared__ __align__(8) uint64_t bar_w, bar_stage, bar_tile[5];
  __shared__ uint32_t s_tmem;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t groups = (p.B + U - 1) / U;

  if (warp == 0) tc::tmem_alloc<256>(&s_tmem);
  if (tid == 32) {
    tc::mbar_init(&bar_w, 1);
    tc::mbar_init(&bar_stage, 1);
    for (int t = 0; t < 5; ++t) tc::mbar_init(&bar_tile[t], 1);
    tc::fence_barrier_init();
  }
  for (int i = tid; i < 12 * R; i += TC_CONV_THREADS) a_hi[i] = make_uint4(0, 0, 0, 0);   // a_hi and a_lo are contiguous
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
  const uint32_t tmem = s_tmem;

  // TMA producer for one group of U utterances (each utterance's 45 planes are one contiguous block in HBM)
  auto issue_stage = [&](int64_t g) {
    uint32_t bytes = 0;
    for (int u = 0; u < U; ++u) {
      const int64_t b = g * U + u;
      if (b < p.B) {
        const uint32_t mis = (uint32_t)(reinterpret_cast<uintptr_t>(p.in + b * (int64_t)R8_C * HW) & 15);
        bytes += (mis + plane_bytes + 15u) & ~15u;
      }
    }
    tc::mbar_expect_tx(&bar_stage, bytes);
    for (int u = 0; u < U; ++u) {
      const int64_t b = g * U + u;
      if (b < p.B) {
        const unsigned char* src = reinterpret_cast<const unsigned char*>(p.in + b * (int64_t)R8_C * HW);
        const uint32_t mis = (uint32_t)(reinterpret_cast<uintptr_t>(src) & 15);
        tc::tma_bulk_g2s(stage + (size_t)u * slot, src - mis, (mis + plane_bytes + 15u) & ~15u, &bar_stage);
      }
    }
  };
  if (tid == 0) {
    tc::mbar_expect_tx(&bar_w, 2 * TC_WBYTES);
    tc::tma_bulk_g2s(w_hi, a.whi, TC_WBYTES, &bar_w);
    tc::tma_bulk_g2s(w_lo, a.wlo, TC_WBYTES, &bar_w);
    if ((int64_t)blockIdx.x < groups) issue_stage(blockIdx.x);
  }
  const uint32_t idesc = tc::instr_desc_bf16(128, TC_N, 0, 0);
  const uint32_t a_hi_s = tc::smem_u32(a_hi), a_lo_s = tc::smem_u32(a_lo);
  const uint32_t w_hi_s = tc::smem_u32(w_hi), w_lo_s = tc::smem_u32(w_lo);

  float st1[24], st2[24];        // statistics of this thread's 24 channels (warp quad) over all its rows
#pragma unroll
  for (int c = 0; c < 24; ++c) st1[c] = st2[c] = 0.f;

  uint32_t phase = 0;
  for (int64_t g = blockIdx.x; g < groups; g += gridDim.x) {
    // ---- raw planes have landed: transform to bf16 (hi, lo) operand rows
    tc::mbar_wait(&bar_stage, phase);
    for (int u = 0; u < U; ++u) {
      const int64_t b = g * U + u;
      const uint32_t mis = (uint32_t)(reinterpret_cast<uintptr_t>(p.in + b * (int64_t)R8_C * HW) & 15);
      tc_transform(reinterpret_cast<const float*>(stage + (size_t)u * slot + mis), b < p.B, H, u * Pu, R, s_mean, s_rstd,
                   a_hi, a_lo, tid);
    }
    tc::fence_proxy_async();
    __syncthreads();            // operands complete, staging buffer free
    {
      if (tid == 0 && g + gridDim.x < groups) issue_stage(g + gridDim.x);   // prefetch the next group during the MMAs
      __syncwarp();
      if (warp == 0 && tc::elect_one()) {   // one elected lane of warp 0 feeds the tensor pipe, then the warp joins the epilogue
        if (g == (int64_t)blockIdx.x) tc::mbar_wait(&bar_w, 0);
        tc::fence_after_sync();
        const uint32_t ah_lo = tc::desc_lo(a_hi_s, (uint32_t)R * 16u), al_lo = tc::desc_lo(a_lo_s, (uint32_t)R * 16u);
        const uint32_t bh_lo = tc::desc_lo(w_hi_s, TC_N * 16u), bl_lo = tc::desc_lo(w_lo_s, TC_N * 16u);
        const uint32_t d_hi128 = tc::desc_hi(128u);
        for (int t = 0; t < tiles; ++t) {
          const uint32_t d = tmem + (uint32_t)(t * TC_N);
          const uint32_t rowb = (uint32_t)(TC_Q0 + 128 * t);
#pragma unroll 1
          for (int tap = 0; tap < 9; ++tap) {
            const int shift = (tap / 3 - 1) * TC_PITCH + (tap % 3 - 1);
#pragma unroll
            for (int ks = 0; ks < 3; ++ks) {
              const uint32_t aoff = (uint32_t)(2 * ks) * (uint32_t)R + rowb + (uint32_t)shift;   // 16-byte units
              const uint32_t boff = (uint32_t)((tap * 6 + 2 * ks) * TC_N);
              const uint64_t ah = tc::desc_make(ah_lo + aoff, d_hi128), al = tc::desc_make(al_lo + aoff, d_hi128);
              const uint64_t bh = tc::desc_make(bh_lo + boff, d_hi128), bl = tc::desc_make(bl_lo + boff, d_hi128);
              tc::umma_bf16(d, al, bh, idesc, (tap | ks) ? 1u : 0u);
              tc::umma_bf16(d, ah, bl, idesc, 1u);
              tc::umma_bf16(d, ah, bh, idesc, 1u);
            }
          }
          tc::umma_commit(&bar_tile[t]);   // tile t can be drained while later tiles are still in the tensor pipe
        }
      }
      __syncwarp();
    }
    {
      // ---- epilogue out of TMEM: thread = one raster row; warp quad h (warps 4h..4h+3) owns channels 24h..24h+23 of
      //      every tile, so both quads start on tile 0 as soon as it commits and the work is balanced
      const int half = warp >> 2;
      for (int t = 0; t < tiles; ++t) {
        const int q = TC_Q0 + 128 * t + 32 * (warp & 3) + lane;
        const int u = q / Pu, rem = q - u * Pu;
        const int y = rem / TC_PITCH - 1, x = rem % TC_PITCH - 1;
        const int64_t b = g * U + u;
        const bool valid = (u < U) && (b < p.B) && (y >= 0) && (y < H) && (x >= 0) && (x < R8_W);
        const int64_t base = valid ? b * (int64_t)R8_C * HW + y * R8_W + x : 0;
        const uint32_t taddr = tmem + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)(t * TC_N + 24 * half);
        // the one extra operand of this mode (residual in the forward, aux in the data gradient) is fetched while the
        // tile is still in the tensor pipe
        const float* extra = (STATS == 2) ? p.aux : p.res;
        float pre[24];
#pragma unroll
        for (int j = 0; j < 24; ++j) {
          const int c = 24 * half + j;
          pre[j] = (extra && c < R8_C && valid) ? __ldg(extra + base + (int64_t)c * HW) : 0.f;
        }
        tc::mbar_wait(&bar_tile[t], phase);
        tc::fence_after_sync();
#pragma unroll
        for (int cb = 0; cb < 3; ++cb) {
          float v[8];
          tc::tmem_ld8(taddr + 8 * cb, v);
          if (valid) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const int jj = cb * 8 + j, c = 24 * half + jj;
              if (c < R8_C) {
                float o = v[j];
                if (RELU) o = fmaxf(o, 0.f);
                if (STATS != 2) o += pre[jj];
                else if (p.res) o += __ldg(p.res + base + (int64_t)c * HW);
                p.out[base + (int64_t)c * HW] = o;
                if (STATS == 1) {
                  st1[jj] += o;
                  st2[jj] = fmaf(o, o, st2[jj]);
                } else if (STATS == 2) {
                  st1[jj] += o;
                  st2[jj] = fmaf(o, (pre[jj] - s_amean[c]) * s_arstd[c], st2[jj]);
                }
              }
            }
          }
        }
      }
      tc::fence_before_sync();
    }
    phase ^= 1u;
    __syncthreads();            // TMEM drained, operand rows reusable
  }
  // ---- per-channel statistics: lanes -> warps -> one fp64 atomic per channel and CTA
  if (STATS) {
    {
#pragma unroll
      for (int j = 0; j < 24; ++j) {
        const float a1 = warp_sum(st1[j]), a2 = warp_sum(st2[j]);
        if (lane == 0) {
          s_red[(warp * 2 + 0) * 24 + j] = a1;
          s_red[(warp * 2 + 1) * 24 + j] = a2;
        }
      }
    }
    __syncthreads();
    if (tid < 2 * R8_C) {
      const int which = tid / R8_C, c = tid - which * R8_C, half = c / 24, j = c - 24 * half;
      double s = 0.0;
      for (int w = 4 * half; w < 4 * half + 4; ++w) s += (double)s_red[(w * 2 + which) * 24 + j];
      atomicAdd(&p.stats[tid], s);
    }
  }
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc<256>(tmem);
}

int r8tc_conv(howl_ctx_t* ctx, cudaStream_t st, const ConvParams& p, const __nv_bfloat16* whi, const __nv_bfloat16* wlo,
              bool relu, int stats) {
  TcConvArgs a;
  a.p = p;
  a.whi = whi;
  a.wlo = wlo;
  a.g = tc_geom(p.H);
  HOWL_REQUIRE(ctx, a.g.U > 0, HOWL_E_UNSUPPORTED, "tensor-core conv: H=%d does not fit", p.H);
  const size_t smem = a.g.smem;
  const int64_t groups = (p.B + a.g.U - 1) / a.g.U;
  const int grid = (int)(groups < ctx->sm_count ? groups : ctx->sm_count);
#define TC_LAUNCH(RELU_, STATS_)                                                                                  \
  do {                                                                                                            \
    HOWL_CUDA(ctx, cudaFuncSetAttribute(conv3x3_tc_kernel<RELU_, STATS_>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                        (int)smem));                                                              \
    conv3x3_tc_kernel<RELU_, STATS_><<<grid, TC_CONV_THREADS, smem, st>>>(a);                                         \
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
// weight-gradient kernel: dW_tap[o][c] = sum_q dC[q][o] * X[q + shift][c], all 9 taps resident in TMEM.
// The M = 128 rows of the A operand are [dC_hi (48) | dC_lo (48) | padding (32)] -- dC_lo sits exactly six 8-channel
// groups behind dC_hi in shared memory -- so two MMAs per k-step (x X_hi, x X_lo) produce hi*hi, lo*hi, hi*lo and
// lo*lo; the epilogue adds TMEM rows o and 48 + o.
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
  const uint32_t plane_bytes = (uint32_t)(R8_C * HW * 4);
  uint4* d_hi = reinterpret_cast<uint4*>(smem);     // [6][Kp]   dC, rows = raster position q
  uint4* d_lo = d_hi + 6 * Kp;
  uint4* x_hi = d_lo + 6 * Kp;                      // [6][Rx]   X,  rows = q + 12
  uint4* x_lo = x_hi + 6 * Rx;
  const size_t operand = (12 * (size_t)Kp + 12 * (size_t)Rx) * 16, need_a = (size_t)16 * Kp * 16;
  unsigned char* stage_x = smem + (operand > need_a ? operand : need_a);
  float* s_mean = reinterpret_cast<float*>(stage_x + tc_slot_bytes(H));
  float* s_rstd = s_mean + 48;
  __shared__ __align__(8) uint64_t bar_stage, bar_d, bar_mma;
  __shared__ uint32_t s_tmem;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t dop_bytes = (uint32_t)(12 * Kp * 16);

  if (warp == 8) {
    tc::tmem_alloc<512>(&s_tmem);
    if (lane == 0) {
      tc::mbar_init(&bar_stage, 1);
      tc::mbar_init(&bar_d, 1);
      tc::mbar_init(&bar_mma, 1);
      tc::fence_barrier_init();
    }
  }
  for (int i = tid; i < 12 * Kp + 12 * Rx; i += TC_THREADS) d_hi[i] = make_uint4(0, 0, 0, 0);
  if (tid < 48) {
    const bool c_ok = tid < R8_C;
    s_mean[tid] = (c_ok && p.x_mean) ? p.x_mean[tid] : 0.f;
    s_rstd[tid] = (c_ok && p.x_rstd) ? p.x_rstd[tid] : 1.f;
  }
  tc::fence_proxy_async();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = s_tmem;
  // X: raw fp32 planes -> staging slot (prefetched one utterance ahead); dC: operand-format image straight into d_hi|d_lo
  auto issue_stage = [&](int64_t b) {
    const unsigned char* sx = reinterpret_cast<const unsigned char*>(p.x + b * (int64_t)R8_C * HW);
    const uint32_t mx = (uint32_t)(reinterpret_cast<uintptr_t>(sx) & 15);
    const uint32_t bx = (mx + plane_bytes + 15u) & ~15u;
    tc::mbar_expect_tx(&bar_stage, bx);
    tc::tma_bulk_g2s(stage_x, sx - mx, bx, &bar_stage);
  };
  auto issue_d = [&](int64_t b) {
    tc::mbar_expect_tx(&bar_d, dop_bytes);
    tc::tma_bulk_g2s(d_hi, reinterpret_cast<const unsigned char*>(p.dc_op) + (size_t)b * dop_bytes, dop_bytes, &bar_d);
  };
  if (tid == 256 && (int64_t)blockIdx.x < p.B) {
    issue_stage(blockIdx.x);
    issue_d(blockIdx.x);
  }
  const uint32_t idesc = tc::instr_desc_bf16(128, TC_N, 1, 1);   // both operands MN-major (K = raster positions)
  const uint32_t d_hi_s = tc::smem_u32(d_hi);
  const uint32_t x_hi_s = tc::smem_u32(x_hi), x_lo_s = tc::smem_u32(x_lo);
  uint32_t phase = 0, first = 1;
  for (int64_t b = blockIdx.x; b < p.B; b += gridDim.x) {
    if (warp < 8) {
      tc::mbar_wait(&bar_stage, phase);
      const uint32_t mx = (uint32_t)(reinterpret_cast<uintptr_t>(p.x + b * (int64_t)R8_C * HW) & 15);
      tc_transform(reinterpret_cast<const float*>(stage_x + mx), true, H, 12, Rx, s_mean, s_rstd, x_hi, x_lo, tid);  // X: rows q + 12
      tc::fence_proxy_async();
    }
    __syncthreads();
    if (tid == 256 && b + gridDim.x < p.B) issue_stage(b + gridDim.x);
    __syncwarp();
    if (warp == 8 && tc::elect_one()) {
      tc::mbar_wait(&bar_d, phase);       // dC operand image of this utterance has landed
      tc::fence_after_sync();
      // MN-major: lbo = stride between 8-position K groups (128 B), sbo = stride between 8-channel groups
      const uint32_t ad_lo = tc::desc_lo(d_hi_s, 128u), ad_hi = tc::desc_hi((uint32_t)Kp * 16u);
      const uint32_t bh_lo = tc::desc_lo(x_hi_s, 128u), bl_lo = tc::desc_lo(x_lo_s, 128u), b_hi = tc::desc_hi((uint32_t)Rx * 16u);
#pragma unroll
      for (int tap = 0; tap < 9; ++tap) {
        const int shift = (tap / 3 - 1) * TC_PITCH + (tap % 3 - 1);
        const uint32_t d = tmem + (uint32_t)(tap * TC_N);
        uint32_t acc = first ? 0u : 1u;
#pragma unroll 4
        for (int k0 = 0; k0 < Kp; k0 += 16) {
          const uint64_t ad = tc::desc_make(ad_lo + (uint32_t)k0, ad_hi);
          const uint32_t boff = (uint32_t)(12 + shift + k0);
          tc::umma_bf16(d, ad, tc::desc_make(bl_lo + boff, b_hi), idesc, acc);
          tc::umma_bf16(d, ad, tc::desc_make(bh_lo + boff, b_hi), idesc, 1u);
          acc = 1u;
        }
      }
      tc::umma_commit(&bar_mma);
    }
    first = 0;
    tc::mbar_wait(&bar_mma, phase);   // operands may be overwritten once the MMAs have drained
    phase ^= 1u;
    tc::fence_after_sync();
    if (tid == 256 && b + gridDim.x < p.B) issue_d(b + gridDim.x);
  }
  // ---- epilogue: TMEM lane r: r < 48 -> dC_hi row of channel r, 48 <= r < 96 -> dC_lo row of channel r - 48
  if (first == 0 && warp < 8) {
    const int r = 32 * (warp & 3) + lane;
    const int o = r < 48 ? r : r - 48;
    const bool ok = r < 96 && o < R8_C;
    const int half = warp >> 2;                    // warps 0-3: taps 0..4, warps 4-7: taps 5..8
    for (int tap = half ? 5 : 0; tap < (half ? 9 : 5); ++tap) {
      const uint32_t taddr = tmem + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)(tap * TC_N);
#pragma unroll
      for (int cb = 0; cb < 3; ++cb) {
        float v[16];
        tc::tmem_ld16(taddr + 16 * cb, v);
        if (ok) {
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int c = cb * 16 + j;
            if (c < R8_C) atomicAdd(&p.dw[(o * R8_C + c) * 9 + tap], v[j]);
          }
        }
      }
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 8) tc::tmem_dealloc<512>(tmem);
}

int r8tc_wgrad(howl_ctx_t* ctx, cudaStream_t st, const WgradParams& p) {
  TcWgradArgs a;
  a.p = p;
  const size_t smem = tc_wgrad_smem(p.H, &a.Kp, &a.Rx);
  HOWL_REQUIRE(ctx, smem <= TC_SMEM_LIMIT, HOWL_E_UNSUPPORTED, "tensor-core wgrad: H=%d does not fit", p.H);
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

// =============================================================================================
// Tensor-core backward, second generation: the conv-output gradient lives in HBM in OPERAND FORMAT (dc_op), so the
// data-gradient kernel has no transform phase: TMA lands the next utterance's operand tile in the second shared-memory
// buffer while the tensor pipe works on the current one, accumulators are double-buffered in TMEM, and the eight worker
// warps do nothing but drain TMEM (store g, reduce the BatchNorm-backward statistics).
// =============================================================================================
__host__ __device__ static inline int r8tc_dcop_rows_dev(int H) { return ((H + 2) * TC_PITCH + 15) & ~15; }
int r8tc_dcop_rows(int H) { return r8tc_dcop_rows_dev(H); }
size_t r8tc_dcop_bytes(int H) { return (size_t)12 * r8tc_dcop_rows(H) * 16; }

// BatchNorm backward + residual fan-in + ReLU mask (same arithmetic as bn_bwd_apply_kernel in res8.cu), one thread =
// 8 channels of one pixel; emits the (hi, lo) bf16 operand rows directly.  grid.y = channel chunk.
template <bool EVEN, bool GU_IN, bool BCAST>
__global__ void __launch_bounds__(256) bn_bwd_apply_op_kernel(const ApplyOpParams p) {
  const int HW = p.H * R8_W, Kp = r8tc_dcop_rows_dev(p.H), chunk = blockIdx.y;
  float mu[8], rs[8], m1[8], m2[8], live[8];
  int coff[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = chunk * 8 + j;
    const int cc = c < R8_C ? c : R8_C - 1;          // clamped address + zero weight: no guards around the loads
    live[j] = c < R8_C ? 1.f : 0.f;
    mu[j] = __ldg(p.mean_rstd + cc);
    rs[j] = __ldg(p.mean_rstd + R8_C + cc);
    m1[j] = (float)(p.stats[cc] / p.count);
    m2[j] = (float)(p.stats[R8_C + cc] / p.count);
    coff[j] = cc * HW;
  }
  const float inv_hw = 1.f / (float)HW;
  const int64_t items = p.B * HW;
  uint4* out = reinterpret_cast<uint4*>(p.dc_op);
  for (int64_t it = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; it < items; it += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = it / HW;
    const int pp = (int)(it - b * HW);
    const int y = pp / R8_W, x = pp - y * R8_W;
    const int64_t ub = b * (int64_t)R8_C * HW + pp;
    float g[8], u[8], gi[8], pv[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {       // every load of the item is issued before any is consumed
      const int64_t idx = ub + coff[j];
      u[j] = p.u[idx];
      g[j] = BCAST ? __ldg(p.g_bcast + b * R8_C + (coff[j] / HW)) * inv_hw : p.g[idx];
      gi[j] = GU_IN ? p.gu_in[idx] : 0.f;
      pv[j] = EVEN ? p.mask_prev[idx] : 0.f;
    }
    float d[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float G = rs[j] * (g[j] - m1[j] - (u[j] - mu[j]) * rs[j] * m2[j]) + gi[j];
      if (EVEN && live[j] != 0.f) p.gu_out[ub + coff[j]] = G;
      d[j] = (u[j] > pv[j]) ? G * live[j] : 0.f;
    }
    uint4 hi, lo;
    tc::split8(d, hi, lo);
    const int64_t base = b * 12 * (int64_t)Kp + (int64_t)chunk * Kp + (y + 1) * TC_PITCH + (x + 1);
    out[base] = hi;
    out[base + 6 * (int64_t)Kp] = lo;
  }
}

int r8tc_apply(howl_ctx_t* ctx, cudaStream_t st, const ApplyOpParams& p) {
  const int64_t items = p.B * p.H * R8_W;
  int64_t bx = howl_ceil_div(items, 256);
  const int64_t cap = (int64_t)ctx->sm_count * 4;
  if (bx > cap) bx = cap;
  const dim3 grid((unsigned)bx, 6, 1);
  const bool even = p.gu_out != nullptr;
  HOWL_REQUIRE(ctx, even == (p.mask_prev != nullptr) && (!p.gu_in || even) && ((p.g != nullptr) != (p.g_bcast != nullptr)),
               HOWL_E_INVALID, "apply: inconsistent arguments");
  if (p.g_bcast) {
    HOWL_REQUIRE(ctx, even && !p.gu_in, HOWL_E_INVALID, "apply: broadcast gradient is the last (even) layer");
    bn_bwd_apply_op_kernel<true, false, true><<<grid, 256, 0, st>>>(p);
  } else if (even && p.gu_in) {
    bn_bwd_apply_op_kernel<true, true, false><<<grid, 256, 0, st>>>(p);
  } else if (even) {
    bn_bwd_apply_op_kernel<true, false, false><<<grid, 256, 0, st>>>(p);
  } else {
    bn_bwd_apply_op_kernel<false, false, false><<<grid, 256, 0, st>>>(p);
  }
  HOWL_LAUNCHED(ctx, "bn_bwd_apply_op");
  return HOWL_OK;
}

#define TCD_THREADS 288          // warps 0-7: epilogue workers, warp 8: TMA + MMA issue
struct TcDgradArgs {
  ConvParams p;                  // p.in unused (operands come from dc_op); out / stats / aux as in the fp32 kernel
  const __nv_bfloat16* dc_op;
  const __nv_bfloat16* whi;
  const __nv_bfloat16* wlo;
  int Kp, tiles;
};

template <int STATS>
__global__ void __launch_bounds__(TCD_THREADS, 1) conv3x3_dgrad_tc_kernel(const TcDgradArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  const ConvParams& p = a.p;
  const int H = p.H, HW = H * R8_W, Kp = a.Kp, tiles = a.tiles;
  const uint32_t op_bytes = (uint32_t)(12 * Kp * 16);
  uint4* w_hi = reinterpret_cast<uint4*>(smem);
  uint4* w_lo = reinterpret_cast<uint4*>(smem + TC_WBYTES);
  unsigned char* a_buf = smem + 2 * TC_WBYTES;                          // 2 x [hi 6][lo 6][Kp] + tail pad
  float* s_f = reinterpret_cast<float*>(a_buf + 2 * (size_t)op_bytes + 128 * 16);
  float* s_amean = s_f;
  float* s_arstd = s_f + 48;
  float* s_red = s_f + 96;       // [8 warps][2][24]
  __shared__ __align__(8) uint64_t bar_w, bar_a[2], bar_tile[2][3], bar_free[2];
  __shared__ uint32_t s_tmem;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  if (warp == 8) {
    tc::tmem_alloc<512>(&s_tmem);
    if (lane == 0) {
      tc::mbar_init(&bar_w, 1);
      for (int i = 0; i < 2; ++i) {
        tc::mbar_init(&bar_a[i], 1);
        tc::mbar_init(&bar_free[i], 8);
        for (int t = 0; t < 3; ++t) tc::mbar_init(&bar_tile[i][t], 1);
      }
      tc::fence_barrier_init();
    }
  }
  // tail pad (rows a tile may read past the second buffer) must be finite
  for (int i = tid; i < 128; i += TCD_THREADS) reinterpret_cast<uint4*>(a_buf + 2 * (size_t)op_bytes)[i] = make_uint4(0, 0, 0, 0);
  if (tid < 48) {
    const bool c_ok = tid < R8_C;
    s_amean[tid] = (c_ok && STATS == 2) ? p.aux_mean[tid] : 0.f;
    s_arstd[tid] = (c_ok && STATS == 2) ? p.aux_rstd[tid] : 0.f;
  }
  tc::fence_proxy_async();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = s_tmem;
  const int64_t n_local = (p.B - blockIdx.x + gridDim.x - 1) / gridDim.x;   // utterances this CTA owns

  if (warp == 8) {
    // ================= producer / issuer warp (converged; one lane elected per instruction) =================
    const unsigned char* src0 = reinterpret_cast<const unsigned char*>(a.dc_op);
    auto load_op = [&](int64_t k) {     // utterance k of this CTA -> buffer k & 1
      if (lane == 0) {
        const int64_t b = blockIdx.x + k * (int64_t)gridDim.x;
        tc::mbar_expect_tx(&bar_a[k & 1], op_bytes);
        tc::tma_bulk_g2s(a_buf + (size_t)(k & 1) * op_bytes, src0 + (size_t)b * op_bytes, op_bytes, &bar_a[k & 1]);
      }
    };
    if (lane == 0) {
      tc::mbar_expect_tx(&bar_w, 2 * TC_WBYTES);
      tc::tma_bulk_g2s(w_hi, a.whi, TC_WBYTES, &bar_w);
      tc::tma_bulk_g2s(w_lo, a.wlo, TC_WBYTES, &bar_w);
    }
    if (n_local > 0) load_op(0);
    if (n_local > 1) load_op(1);
    __syncwarp();
    if (tc::elect_one()) {
      tc::mbar_wait(&bar_w, 0);
      const uint32_t idesc = tc::instr_desc_bf16(128, TC_N, 0, 0);
      const uint32_t w_hi_s = tc::smem_u32(w_hi), w_lo_s = tc::smem_u32(w_lo);
      const uint32_t bh_lo = tc::desc_lo(w_hi_s, TC_N * 16u), bl_lo = tc::desc_lo(w_lo_s, TC_N * 16u);
      const uint32_t d_hi128 = tc::desc_hi(128u);
      const uint32_t a_base = tc::smem_u32(a_buf);
      for (int64_t k = 0; k < n_local; ++k) {
        const int buf = (int)(k & 1);
        const uint32_t par = (uint32_t)((k >> 1) & 1);
        tc::mbar_wait(&bar_a[buf], par);                       // operand tile landed
        if (k >= 2) tc::mbar_wait(&bar_free[buf], par ^ 1u);   // epilogue of utterance k-2 has drained this TMEM half
        tc::fence_after_sync();
        const uint32_t a_hi_s = a_base + (uint32_t)buf * op_bytes, a_lo_s = a_hi_s + (uint32_t)(6 * Kp * 16);
        const uint32_t ah_lo = tc::desc_lo(a_hi_s, (uint32_t)Kp * 16u), al_lo = tc::desc_lo(a_lo_s, (uint32_t)Kp * 16u);
        for (int t = 0; t < tiles; ++t) {
          const uint32_t d = tmem + (uint32_t)(buf * 256 + t * TC_N);
          const uint32_t rowb = (uint32_t)(TC_Q0 + 128 * t);
#pragma unroll 1
          for (int tap = 0; tap < 9; ++tap) {
            const int shift = (tap / 3 - 1) * TC_PITCH + (tap % 3 - 1);
#pragma unroll
            for (int ks = 0; ks < 3; ++ks) {
              const uint32_t aoff = (uint32_t)(2 * ks) * (uint32_t)Kp + rowb + (uint32_t)shift;
              const uint32_t boff = (uint32_t)((tap * 6 + 2 * ks) * TC_N);
              const uint64_t ah = tc::desc_make(ah_lo + aoff, d_hi128), al = tc::desc_make(al_lo + aoff, d_hi128);
              const uint64_t bh = tc::desc_make(bh_lo + boff, d_hi128), bl = tc::desc_make(bl_lo + boff, d_hi128);
              tc::umma_bf16(d, al, bh, idesc, (tap | ks) ? 1u : 0u);
              tc::umma_bf16(d, ah, bl, idesc, 1u);
              tc::umma_bf16(d, ah, bh, idesc, 1u);
            }
          }
          tc::umma_commit(&bar_tile[buf][t]);
        }
        // refill the other buffer with utterance k + 1 once the MMAs of utterance k - 1 have released it
        if (k >= 1 && k + 1 < n_local) {
          tc::mbar_wait(&bar_tile[buf ^ 1][tiles - 1], (uint32_t)(((k - 1) >> 1) & 1));
          const int64_t b = blockIdx.x + (k + 1) * (int64_t)gridDim.x;
          tc::mbar_expect_tx(&bar_a[(k + 1) & 1], op_bytes);
          tc::tma_bulk_g2s(a_buf + (size_t)((k + 1) & 1) * op_bytes, src0 + (size_t)b * op_bytes, op_bytes, &bar_a[(k + 1) & 1]);
        }
      }
    }
    __syncwarp();
  } else {
    // ================= epilogue workers: thread = one raster row, warp quad = 24 channels =================
    const int half = warp >> 2;
    float st1[24], st2[24];
#pragma unroll
    for (int c = 0; c < 24; ++c) st1[c] = st2[c] = 0.f;
    for (int64_t k = 0; k < n_local; ++k) {
      const int buf = (int)(k & 1);
      const uint32_t par = (uint32_t)((k >> 1) & 1);
      const int64_t b = blockIdx.x + k * (int64_t)gridDim.x;
      for (int t = 0; t < tiles; ++t) {
        const int q = TC_Q0 + 128 * t + 32 * (warp & 3) + lane;
        const int y = q / TC_PITCH - 1, x = q % TC_PITCH - 1;
        const bool valid = (y >= 0) && (y < H) && (x >= 0) && (x < R8_W);
        const int64_t base = valid ? b * (int64_t)R8_C * HW + y * R8_W + x : 0;
        const uint32_t taddr = tmem + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)(buf * 256 + t * TC_N + 24 * half);
        float pre[24];
#pragma unroll
        for (int j = 0; j < 24; ++j) {
          const int c = 24 * half + j;
          pre[j] = (STATS == 2 && c < R8_C && valid) ? __ldg(p.aux + base + (int64_t)c * HW) : 0.f;
        }
        tc::mbar_wait(&bar_tile[buf][t], par);
        tc::fence_after_sync();
#pragma unroll
        for (int cb = 0; cb < 3; ++cb) {
          float v[8];
          tc::tmem_ld8(taddr + 8 * cb, v);
          if (valid) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const int jj = cb * 8 + j, c = 24 * half + jj;
              if (c < R8_C) {
                const float o = v[j];
                p.out[base + (int64_t)c * HW] = o;
                if (STATS == 2) {
                  st1[jj] += o;
                  st2[jj] = fmaf(o, (pre[jj] - s_amean[c]) * s_arstd[c], st2[jj]);
                }
              }
            }
          }
        }
      }
      tc::fence_before_sync();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&bar_free[buf]);
    }
    if (STATS) {
#pragma unroll
      for (int j = 0; j < 24; ++j) {
        const float a1 = warp_sum(st1[j]), a2 = warp_sum(st2[j]);
        if (lane == 0) {
          s_red[(warp * 2 + 0) * 24 + j] = a1;
          s_red[(warp * 2 + 1) * 24 + j] = a2;
        }
      }
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (STATS && tid < 2 * R8_C) {
    const int which = tid / R8_C, c = tid - which * R8_C, half = c / 24, j = c - 24 * half;
    double s = 0.0;
    for (int w = 4 * half; w < 4 * half + 4; ++w) s += (double)s_red[(w * 2 + which) * 24 + j];
    atomicAdd(&p.stats[tid], s);
  }
  if (warp == 8) tc::tmem_dealloc<512>(tmem);
}

int r8tc_dgrad(howl_ctx_t* ctx, cudaStream_t st, const ConvParams& p, const __nv_bfloat16* dc_op, const __nv_bfloat16* whi,
               const __nv_bfloat16* wlo, int stats) {
  TcDgradArgs a;
  a.p = p;
  a.dc_op = dc_op;
  a.whi = whi;
  a.wlo = wlo;
  a.Kp = r8tc_dcop_rows(p.H);
  a.tiles = (TC_PITCH * p.H - 1 + 127) / 128;
  HOWL_REQUIRE(ctx, a.tiles <= 3, HOWL_E_UNSUPPORTED, "tensor-core dgrad: H=%d needs more than 3 tiles", p.H);
  const size_t smem = 2 * (size_t)TC_WBYTES + 2 * r8tc_dcop_bytes(p.H) + 128 * 16 + (96 + 8 * 2 * 24) * 4;
  HOWL_REQUIRE(ctx, smem <= TC_SMEM_LIMIT, HOWL_E_UNSUPPORTED, "tensor-core dgrad: H=%d does not fit", p.H);
  const int grid = (int)(p.B < ctx->sm_count ? p.B : ctx->sm_count);
  if (stats == 2) {
    HOWL_CUDA(ctx, cudaFuncSetAttribute(conv3x3_dgrad_tc_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    conv3x3_dgrad_tc_kernel<2><<<grid, TCD_THREADS, smem, st>>>(a);
  } else {
    HOWL_CUDA(ctx, cudaFuncSetAttribute(conv3x3_dgrad_tc_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    conv3x3_dgrad_tc_kernel<0><<<grid, TCD_THREADS, smem, st>>>(a);
  }
  HOWL_LAUNCHED(ctx, "conv3x3_dgrad_tc");
  return HOWL_OK;
}

// =============================================================================================
// Forward, second generation: activations are ALSO kept in operand format ("u_op": per utterance
// [hi,lo][6 chunks][Rx = Kp + 24 rows][8 bf16], raster row q at row q + 12, channel 45 = 1 at valid pixels, halo rows 0),
// written by the producer's epilogue.  BatchNorm of the producer is folded into the consumer:
//     conv_W( (u - mean) * rstd, zero padded ) = conv_{W'}( u with halo := mean ) + bias,
//     W'[o][c] = W[o][c] * rstd[c],   bias[o] = - sum_{c,tap} W'[o][c][tap] * mean[c]
// so the forward needs no transform either: TMA lands the raw operand tile, the workers only overwrite the ~50 halo rows
// with split(mean), and the rest is the double-buffered pipeline of the data-gradient kernel.
// =============================================================================================
int r8tc_uop_rows(int H) { return r8tc_dcop_rows(H) + 24; }
size_t r8tc_uop_bytes(int H) { return (size_t)12 * r8tc_uop_rows(H) * 16; }

// per layer: folded (hi, lo) weights in the operand layout, bias[48], halo rows (split mean) [2][6] x 16 B
__global__ void tc_fold_kernel(const float* __restrict__ w, const float* __restrict__ mean_rstd,
                               __nv_bfloat16* __restrict__ whi, __nv_bfloat16* __restrict__ wlo, float* __restrict__ bias,
                               uint4* __restrict__ halo) {
  __shared__ float s_mu[48], s_rs[48];
  const int tid = threadIdx.x;
  if (tid < 48) {
    const bool ok = tid < R8_C && mean_rstd != nullptr;
    s_mu[tid] = ok ? mean_rstd[tid] : 0.f;
    s_rs[tid] = (tid < R8_C) ? (mean_rstd ? mean_rstd[R8_C + tid] : 1.f) : 0.f;
  }
  __syncthreads();
  for (int r = tid; r < R8TC_WBLOCK; r += blockDim.x) {
    const int j = r & 7, n = (r >> 3) % TC_N, chunk = ((r >> 3) / TC_N) % 6, tap = (r >> 3) / (TC_N * 6);
    const int c = chunk * 8 + j;
    float v = 0.f;
    if (n < R8_C && c < R8_C) v = w[(n * R8_C + c) * 9 + tap] * s_rs[c];
    const __nv_bfloat16 hi = __float2bfloat16_rn(v);
    whi[r] = hi;
    wlo[r] = __float2bfloat16_rn(v - __bfloat162float(hi));
  }
  if (tid < 48) {
    double b = 0.0;
    if (tid < R8_C)
      for (int c = 0; c < R8_C; ++c) {
        double s = 0.0;
        for (int tap = 0; tap < 9; ++tap) s += (double)w[(tid * R8_C + c) * 9 + tap];
        b -= s * (double)s_rs[c] * (double)s_mu[c];
      }
    bias[tid] = (float)b;
  }
  if (tid < 6) {
    float v[8];
    for (int j = 0; j < 8; ++j) v[j] = s_mu[tid * 8 + j];     // channels 45..47 (incl. the ones channel) have mean 0
    uint4 hi, lo;
    tc::split8(v, hi, lo);
    halo[tid] = hi;
    halo[6 + tid] = lo;
  }
}

struct TcFwdArgs {
  ConvParams p;                  // out (planar fp32), res, stats; p.in unused
  const __nv_bfloat16* in_op;    // u_{i-1} in operand format
  __nv_bfloat16* out_op;         // u_i in operand format, or null (eval)
  const __nv_bfloat16* whi;      // folded weights of this layer
  const __nv_bfloat16* wlo;
  const float* bias;             // [48]
  const uint4* halo;             // [2][6]
  int Kp, Rx, tiles;
};

template <int STATS>
__global__ void __launch_bounds__(TCD_THREADS, 1) conv3x3_fwd_op_tc_kernel(const TcFwdArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  const ConvParams& p = a.p;
  const int H = p.H, HW = H * R8_W, Kp = a.Kp, Rx = a.Rx, tiles = a.tiles;
  const uint32_t op_bytes = (uint32_t)(12 * Rx * 16);
  uint4* w_hi = reinterpret_cast<uint4*>(smem);
  uint4* w_lo = reinterpret_cast<uint4*>(smem + TC_WBYTES);
  unsigned char* a_buf = smem + 2 * TC_WBYTES;                          // 2 x [hi 6][lo 6][Rx] + tail pad
  float* s_f = reinterpret_cast<float*>(a_buf + 2 * (size_t)op_bytes + 128 * 16);
  float* s_bias = s_f;           // [48]
  float* s_red = s_f + 48;       // [8 warps][2][24]
  uint4* s_halo = reinterpret_cast<uint4*>(s_f + 48 + 8 * 2 * 24);      // [2][6]
  __shared__ __align__(8) uint64_t bar_w, bar_a[2], bar_halo[2], bar_tile[2][3], bar_free[2];
  __shared__ uint32_t s_tmem;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  if (warp == 8) {
    tc::tmem_alloc<512>(&s_tmem);
    if (lane == 0) {
      tc::mbar_init(&bar_w, 1);
      for (int i = 0; i < 2; ++i) {
        tc::mbar_init(&bar_a[i], 1);
        tc::mbar_init(&bar_halo[i], 8);
        tc::mbar_init(&bar_free[i], 8);
        for (int t = 0; t < 3; ++t) tc::mbar_init(&bar_tile[i][t], 1);
      }
      tc::fence_barrier_init();
    }
  }
  for (int i = tid; i < 128; i += TCD_THREADS) reinterpret_cast<uint4*>(a_buf + 2 * (size_t)op_bytes)[i] = make_uint4(0, 0, 0, 0);
  if (tid < 48) s_bias[tid] = a.bias[tid];
  if (tid < 12) s_halo[tid] = a.halo[tid];
  tc::fence_proxy_async();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = s_tmem;
  const int64_t n_local = (p.B - blockIdx.x + gridDim.x - 1) / gridDim.x;

  if (warp == 8) {
    // ================= producer / issuer warp =================
    const unsigned char* src0 = reinterpret_cast<const unsigned char*>(a.in_op);
    if (tc::elect_one()) {
      auto load_op = [&](int64_t k) {
        const int64_t b = blockIdx.x + k * (int64_t)gridDim.x;
        tc::mbar_expect_tx(&bar_a[k & 1], op_bytes);
        tc::tma_bulk_g2s(a_buf + (size_t)(k & 1) * op_bytes, src0 + (size_t)b * op_bytes, op_bytes, &bar_a[k & 1]);
      };
      tc::mbar_expect_tx(&bar_w, 2 * TC_WBYTES);
      tc::tma_bulk_g2s(w_hi, a.whi, TC_WBYTES, &bar_w);
      tc::tma_bulk_g2s(w_lo, a.wlo, TC_WBYTES, &bar_w);
      if (n_local > 0) load_op(0);
      if (n_local > 1) load_op(1);
      tc::mbar_wait(&bar_w, 0);
      const uint32_t idesc = tc::instr_desc_bf16(128, TC_N, 0, 0);
      const uint32_t w_hi_s = tc::smem_u32(w_hi), w_lo_s = tc::smem_u32(w_lo);
      const uint32_t bh_lo = tc::desc_lo(w_hi_s, TC_N * 16u), bl_lo = tc::desc_lo(w_lo_s, TC_N * 16u);
      const uint32_t d_hi128 = tc::desc_hi(128u);
      const uint32_t a_base = tc::smem_u32(a_buf);
      for (int64_t k = 0; k < n_local; ++k) {
        const int buf = (int)(k & 1);
        const uint32_t par = (uint32_t)((k >> 1) & 1);
        tc::mbar_wait(&bar_halo[buf], par);                    // operand tile landed AND halo rows rewritten by the workers
        if (k >= 2) tc::mbar_wait(&bar_free[buf], par ^ 1u);   // epilogue of utterance k-2 has drained this TMEM half
        tc::fence_after_sync();
        const uint32_t a_hi_s = a_base + (uint32_t)buf * op_bytes, a_lo_s = a_hi_s + (uint32_t)(6 * Rx * 16);
        const uint32_t ah_lo = tc::desc_lo(a_hi_s, (uint32_t)Rx * 16u), al_lo = tc::desc_lo(a_lo_s, (uint32_t)Rx * 16u);
        for (int t = 0; t < tiles; ++t) {
          const uint32_t d = tmem + (uint32_t)(buf * 256 + t * TC_N);
          const uint32_t rowb = (uint32_t)(12 + TC_Q0 + 128 * t);      // raster row q lives at operand row q + 12
#pragma unroll 1
          for (int tap = 0; tap < 9; ++tap) {
            const int shift = (tap / 3 - 1) * TC_PITCH + (tap % 3 - 1);
#pragma unroll
            for (int ks = 0; ks < 3; ++ks) {
              const uint32_t aoff = (uint32_t)(2 * ks) * (uint32_t)Rx + rowb + (uint32_t)shift;
              const uint32_t boff = (uint32_t)((tap * 6 + 2 * ks) * TC_N);
              const uint64_t ah = tc::desc_make(ah_lo + aoff, d_hi128), al = tc::desc_make(al_lo + aoff, d_hi128);
              const uint64_t bh = tc::desc_make(bh_lo + boff, d_hi128), bl = tc::desc_make(bl_lo + boff, d_hi128);
              tc::umma_bf16(d, al, bh, idesc, (tap | ks) ? 1u : 0u);
              tc::umma_bf16(d, ah, bl, idesc, 1u);
              tc::umma_bf16(d, ah, bh, idesc, 1u);
            }
          }
          tc::umma_commit(&bar_tile[buf][t]);
        }
        if (k >= 1 && k + 1 < n_local) {
          tc::mbar_wait(&bar_tile[buf ^ 1][tiles - 1], (uint32_t)(((k - 1) >> 1) & 1));
          load_op(k + 1);
        }
      }
    }
    __syncwarp();
  } else {
    // ================= workers: halo rows of the NEXT operand tile, then the epilogue of the current one =================
    const int half = warp >> 2;
    float st1[24], st2[24];
#pragma unroll
    for (int c = 0; c < 24; ++c) st1[c] = st2[c] = 0.f;
    auto fill_halo = [&](int64_t k) {     // utterance k of this CTA: wait for its TMA, overwrite the invalid raster rows with mean
      const int buf = (int)(k & 1);
      tc::mbar_wait(&bar_a[buf], (uint32_t)((k >> 1) & 1));
      uint4* hi = reinterpret_cast<uint4*>(a_buf + (size_t)buf * op_bytes);
      uint4* lo = hi + 6 * Rx;
      for (int q = tid; q < Kp; q += TC_WORKERS) {
        const int y = q / TC_PITCH - 1, x = q % TC_PITCH - 1;
        if (y >= 0 && y < H && x >= 0 && x < R8_W) continue;
#pragma unroll
        for (int ch = 0; ch < 6; ++ch) {
          hi[ch * Rx + 12 + q] = s_halo[ch];
          lo[ch * Rx + 12 + q] = s_halo[6 + ch];
        }
      }
      tc::fence_proxy_async();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&bar_halo[buf]);
    };
    if (n_local > 0) fill_halo(0);
    for (int64_t k = 0; k < n_local; ++k) {
      const int buf = (int)(k & 1);
      const uint32_t par = (uint32_t)((k >> 1) & 1);
      const int64_t b = blockIdx.x + k * (int64_t)gridDim.x;
      if (k + 1 < n_local) fill_halo(k + 1);
      uint4* o_hi = a.out_op ? reinterpret_cast<uint4*>(a.out_op) + (size_t)b * 12 * Rx : nullptr;
      if (o_hi) {   // guard rows the tile threads never reach
        for (int i = tid; i < 24 * 12; i += TC_WORKERS) o_hi[(i / 24) * Rx + (i % 24)] = make_uint4(0, 0, 0, 0);
      }
      for (int t = 0; t < tiles; ++t) {
        const int q = TC_Q0 + 128 * t + 32 * (warp & 3) + lane;
        const int y = q / TC_PITCH - 1, x = q % TC_PITCH - 1;
        const bool valid = (y >= 0) && (y < H) && (x >= 0) && (x < R8_W);
        const int64_t base = valid ? b * (int64_t)R8_C * HW + y * R8_W + x : 0;
        const uint32_t taddr = tmem + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)(buf * 256 + t * TC_N + 24 * half);
        float pre[24];
#pragma unroll
        for (int j = 0; j < 24; ++j) {
          const int c = 24 * half + j;
          pre[j] = (p.res && c < R8_C && valid) ? __ldg(p.res + base + (int64_t)c * HW) : 0.f;
        }
        tc::mbar_wait(&bar_tile[buf][t], par);
        tc::fence_after_sync();
#pragma unroll
        for (int cb = 0; cb < 3; ++cb) {
          float v[8], ov[8];
          tc::tmem_ld8(taddr + 8 * cb, v);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int jj = cb * 8 + j, c = 24 * half + jj;
            float o = 0.f;
            if (valid && c < R8_C) {
              o = fmaxf(v[j] + s_bias[c], 0.f) + pre[jj];
              p.out[base + (int64_t)c * HW] = o;
              if (STATS == 1) {
                st1[jj] += o;
                st2[jj] = fmaf(o, o, st2[jj]);
              }
            } else if (valid && c == R8_C) {
              o = 1.f;                      // the "ones" channel: lets the weight gradient fold BatchNorm (sum of dC per tap)
            }
            ov[j] = o;
          }
          if (o_hi && q + 12 < Rx) {
            uint4 hi, lo;
            tc::split8(ov, hi, lo);
            const int ch = 3 * half + cb;
            o_hi[ch * Rx + 12 + q] = hi;
            o_hi[(6 + ch) * Rx + 12 + q] = lo;
          }
        }
      }
      tc::fence_before_sync();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&bar_free[buf]);
    }
    if (STATS) {
#pragma unroll
      for (int j = 0; j < 24; ++j) {
        const float a1 = warp_sum(st1[j]), a2 = warp_sum(st2[j]);
        if (lane == 0) {
          s_red[(warp * 2 + 0) * 24 + j] = a1;
          s_red[(warp * 2 + 1) * 24 + j] = a2;
        }
      }
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (STATS && tid < 2 * R8_C) {
    const int which = tid / R8_C, c = tid - which * R8_C, half = c / 24, j = c - 24 * half;
    double s = 0.0;
    for (int w = 4 * half; w < 4 * half + 4; ++w) s += (double)s_red[(w * 2 + which) * 24 + j];
    atomicAdd(&p.stats[tid], s);
  }
  if (warp == 8) tc::tmem_dealloc<512>(tmem);
}

int r8tc_fold(howl_ctx_t* ctx, cudaStream_t st, const float* w_layer, const float* mean_rstd, __nv_bfloat16* whi,
              __nv_bfloat16* wlo, float* bias, void* halo) {
  tc_fold_kernel<<<1, 512, 0, st>>>(w_layer, mean_rstd, whi, wlo, bias, reinterpret_cast<uint4*>(halo));
  HOWL_LAUNCHED(ctx, "tc_fold");
  return HOWL_OK;
}

int r8tc_fwd_op(howl_ctx_t* ctx, cudaStream_t st, const ConvParams& p, const __nv_bfloat16* in_op, __nv_bfloat16* out_op,
                const __nv_bfloat16* whi, const __nv_bfloat16* wlo, const float* bias, const void* halo, int stats) {
  TcFwdArgs a;
  a.p = p;
  a.in_op = in_op; a.out_op = out_op; a.whi = whi; a.wlo = wlo; a.bias = bias; a.halo = reinterpret_cast<const uint4*>(halo);
  a.Kp = r8tc_dcop_rows(p.H);
  a.Rx = r8tc_uop_rows(p.H);
  a.tiles = (TC_PITCH * p.H - 1 + 127) / 128;
  HOWL_REQUIRE(ctx, a.tiles <= 3, HOWL_E_UNSUPPORTED, "tensor-core forward: H=%d needs more than 3 tiles", p.H);
  const size_t smem = 2 * (size_t)TC_WBYTES + 2 * r8tc_uop_bytes(p.H) + 128 * 16 + (48 + 8 * 2 * 24) * 4 + 12 * 16;
  HOWL_REQUIRE(ctx, smem <= TC_SMEM_LIMIT, HOWL_E_UNSUPPORTED, "tensor-core forward: H=%d does not fit", p.H);
  const int grid = (int)(p.B < ctx->sm_count ? p.B : ctx->sm_count);
  if (stats) {
    HOWL_CUDA(ctx, cudaFuncSetAttribute(conv3x3_fwd_op_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    conv3x3_fwd_op_tc_kernel<1><<<grid, TCD_THREADS, smem, st>>>(a);
  } else {
    HOWL_CUDA(ctx, cudaFuncSetAttribute(conv3x3_fwd_op_tc_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    conv3x3_fwd_op_tc_kernel<0><<<grid, TCD_THREADS, smem, st>>>(a);
  }
  HOWL_LAUNCHED(ctx, "conv3x3_fwd_tc");
  return HOWL_OK;
}

// =============================================================================================
// Weight gradient, second generation: BOTH operands arrive by TMA in operand format (dC from the BatchNorm-backward
// kernel, X = u_op from the forward), so there is no staging or transform and the eight epilogue warps sleep until the
// accumulators are final.  X is double buffered; dC is single buffered but split in two K halves that are issued
// half-outer, so the next utterance's first half lands while the second half of this one is being multiplied.
// BatchNorm of X is folded into the epilogue through the "ones" channel (column 45 of every tap):
//     dW[o][c] = rstd[c] * ( sum_q dC[q][o] X[q+s][c]  -  mean[c] * sum_q dC[q][o] 1[q+s] )
// =============================================================================================
struct TcWgrad2Args {
  const __nv_bfloat16* dc_op;
  const __nv_bfloat16* x_op;
  const float* x_mean;   // or null (layer 1: X = a0, no normalisation)
  const float* x_rstd;
  float* dw;
  int64_t B;
  int Kp, Rx, Kh;
};

__global__ void __launch_bounds__(TC_THREADS, 1) conv3x3_wgrad_op_tc_kernel(const TcWgrad2Args a) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int Kp = a.Kp, Rx = a.Rx, Kh = a.Kh;
  const uint32_t x_bytes = (uint32_t)(12 * Rx * 16);
  unsigned char* d_buf = smem;                                   // [hi 6 | lo 6][Kp] x 16 B; M groups 12..15 run into x_buf
  unsigned char* x_buf = smem + (size_t)12 * Kp * 16;            // 2 x [hi 6 | lo 6][Rx] x 16 B
  __shared__ __align__(8) uint64_t bar_x[2], bar_d[2], bar_h[2];
  __shared__ uint32_t s_tmem;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  if (warp == 8) {
    tc::tmem_alloc<512>(&s_tmem);
    if (lane == 0) {
      for (int i = 0; i < 2; ++i) {
        tc::mbar_init(&bar_x[i], 1);
        tc::mbar_init(&bar_d[i], 1);
        tc::mbar_init(&bar_h[i], 1);
      }
      tc::fence_barrier_init();
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = s_tmem;
  const int64_t n_local = (a.B - blockIdx.x + gridDim.x - 1) / gridDim.x;

  if (warp == 8) {
    if (tc::elect_one() && n_local > 0) {
      const unsigned char* xsrc = reinterpret_cast<const unsigned char*>(a.x_op);
      const unsigned char* dsrc = reinterpret_cast<const unsigned char*>(a.dc_op);
      const uint32_t d_bytes = (uint32_t)(12 * Kp * 16);
      auto load_x = [&](int64_t k) {
        const int64_t b = blockIdx.x + k * (int64_t)gridDim.x;
        tc::mbar_expect_tx(&bar_x[k & 1], x_bytes);
        tc::tma_bulk_g2s(x_buf + (size_t)(k & 1) * x_bytes, xsrc + (size_t)b * x_bytes, x_bytes, &bar_x[k & 1]);
      };
      auto load_d = [&](int64_t k, int hf) {     // K rows [hf * Kh, hf ? Kp : Kh) of all twelve 8-channel groups
        const int64_t b = blockIdx.x + k * (int64_t)gridDim.x;
        const uint32_t r0 = hf ? (uint32_t)Kh : 0u, nr = hf ? (uint32_t)(Kp - Kh) : (uint32_t)Kh;
        tc::mbar_expect_tx(&bar_d[hf], 12u * nr * 16u);
        for (uint32_t g = 0; g < 12; ++g)
          tc::tma_bulk_g2s(d_buf + ((size_t)g * Kp + r0) * 16, dsrc + (size_t)b * d_bytes + ((size_t)g * Kp + r0) * 16, nr * 16u,
                           &bar_d[hf]);
      };
      load_x(0);
      load_d(0, 0);
      load_d(0, 1);
      if (n_local > 1) load_x(1);
      const uint32_t idesc = tc::instr_desc_bf16(128, TC_N, 1, 1);   // both operands MN-major (K = raster positions)
      const uint32_t d_s = tc::smem_u32(d_buf), x_s = tc::smem_u32(x_buf);
      const uint32_t ad_lo = tc::desc_lo(d_s, 128u), ad_hi = tc::desc_hi((uint32_t)Kp * 16u);
      const uint32_t b_hi = tc::desc_hi((uint32_t)Rx * 16u);
      for (int64_t k = 0; k < n_local; ++k) {
        const uint32_t par = (uint32_t)(k & 1);
        const uint32_t xh_s = x_s + (uint32_t)(k & 1) * x_bytes;
        const uint32_t bh_lo = tc::desc_lo(xh_s, 128u), bl_lo = tc::desc_lo(xh_s + (uint32_t)(6 * Rx * 16), 128u);
        tc::mbar_wait(&bar_x[k & 1], (uint32_t)((k >> 1) & 1));
#pragma unroll 1
        for (int hf = 0; hf < 2; ++hf) {
          if (hf == 1 && k > 0) {
            // utterance k-1 is completely multiplied (its MMAs precede this one's first half in the pipe): its second dC half
            // and its X buffer are free.  Both loads land while the first half of utterance k is being multiplied.
            tc::mbar_wait(&bar_h[1], par ^ 1u);
            load_d(k, 1);
            if (k + 1 < n_local) load_x(k + 1);
          }
          tc::mbar_wait(&bar_d[hf], par);
          tc::fence_after_sync();
          const int kb = hf ? Kh : 0, ke = hf ? Kp : Kh;
#pragma unroll 1
          for (int tap = 0; tap < 9; ++tap) {
            const int shift = (tap / 3 - 1) * TC_PITCH + (tap % 3 - 1);
            const uint32_t d = tmem + (uint32_t)(tap * TC_N);
            uint32_t acc = (k == 0 && hf == 0) ? 0u : 1u;
#pragma unroll 2
            for (int k0 = kb; k0 < ke; k0 += 16) {
              const uint64_t ad = tc::desc_make(ad_lo + (uint32_t)k0, ad_hi);
              const uint32_t boff = (uint32_t)(12 + shift + k0);
              tc::umma_bf16(d, ad, tc::desc_make(bl_lo + boff, b_hi), idesc, acc);
              tc::umma_bf16(d, ad, tc::desc_make(bh_lo + boff, b_hi), idesc, 1u);
              acc = 1u;
            }
          }
          tc::umma_commit(&bar_h[hf]);
        }
        // first half done -> its dC rows take utterance k+1's first half while the second half still multiplies
        if (k + 1 < n_local) {
          tc::mbar_wait(&bar_h[0], par);
          load_d(k + 1, 0);
        }
      }
      tc::mbar_wait(&bar_h[1], (uint32_t)((n_local - 1) & 1));
    }
    __syncwarp();
  }
  tc::fence_before_sync();
  __syncthreads();          // the issuer arrives only after the last commit: every accumulator is final
  tc::fence_after_sync();
  if (n_local > 0 && warp < 8) {
    const int r = 32 * (warp & 3) + lane;
    const int o = r < 48 ? r : r - 48;
    const bool ok = r < 96 && o < R8_C;
    const int half = warp >> 2;                    // warps 0-3: taps 0..4, warps 4-7: taps 5..8
    for (int tap = half ? 5 : 0; tap < (half ? 9 : 5); ++tap) {
      const uint32_t taddr = tmem + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)(tap * TC_N);
      float v[48];
      tc::tmem_ld16(taddr, v);
      tc::tmem_ld16(taddr + 16, v + 16);
      tc::tmem_ld16(taddr + 32, v + 32);
      if (ok) {
        const float ones = v[R8_C];
#pragma unroll
        for (int c = 0; c < R8_C; ++c) {
          const float mu = a.x_mean ? __ldg(a.x_mean + c) : 0.f, rs = a.x_rstd ? __ldg(a.x_rstd + c) : 1.f;
          atomicAdd(&a.dw[(o * R8_C + c) * 9 + tap], rs * (v[c] - mu * ones));
        }
      }
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 8) tc::tmem_dealloc<512>(tmem);
}

int r8tc_wgrad_op(howl_ctx_t* ctx, cudaStream_t st, const __nv_bfloat16* dc_op, const __nv_bfloat16* x_op, const float* x_mean,
                  const float* x_rstd, float* dw, int64_t B, int H) {
  TcWgrad2Args a;
  a.dc_op = dc_op; a.x_op = x_op; a.x_mean = x_mean; a.x_rstd = x_rstd; a.dw = dw; a.B = B;
  a.Kp = r8tc_dcop_rows(H);
  a.Rx = r8tc_uop_rows(H);
  a.Kh = (a.Kp / 32) * 16;
  const size_t smem = (size_t)12 * a.Kp * 16 + 2 * r8tc_uop_bytes(H);
  HOWL_REQUIRE(ctx, smem <= TC_SMEM_LIMIT && a.Kh >= 16, HOWL_E_UNSUPPORTED, "tensor-core wgrad: H=%d does not fit", H);
  HOWL_CUDA(ctx, cudaFuncSetAttribute(conv3x3_wgrad_op_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int grid = (int)(B < ctx->sm_count ? B : ctx->sm_count);
  conv3x3_wgrad_op_tc_kernel<<<grid, TC_THREADS, smem, st>>>(a);
  HOWL_LAUNCHED(ctx, "conv3x3_wgrad_tc");
  return HOWL_OK;
}
