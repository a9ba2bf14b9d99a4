// Res8 45->45 3x3 convolutions on the 5th-generation tensor cores (tcgen05.mma + TMEM), sm_100a.
//
// Every tensor a convolution reads lives in HBM in OPERAND FORMAT (res8_common.cuh): per utterance
//     [hi, lo][6 chunks of 8 channels][R raster rows][8 x bf16]          R = round_up((H + 2) * 11, 64)
// in the padded "pitch 11" raster  q(y, x) = (y + 1) * 11 + (x + 1)  (one shared halo column between image rows), so
// that a 3x3 tap is a shift of the operand start address by (dy - 1) * 11 + (dx - 1) rows and no kernel stages or
// transforms anything: TMA bulk copies land the rows, tcgen05.mma reads them.  fp32 values are split x = hi + lo (bf16
// each) and a product is hi*hi + hi*lo + lo*hi with fp32 accumulation in TMEM (tc_common.cuh).
//
//  * forward / data gradient ("stream" kernel): the rasters of consecutive utterances form one row stream that is cut
//    into 128-row M tiles; two utterances (2R = 5 x 128 rows for 1 s clips) sit in a shared-memory ring, so only the halo
//    rows of the raster are wasted M rows.  Per tile, tap and 16-channel K step: one N = 96 MMA  A_hi x [W_hi | W_lo]  and
//    one N = 48 MMA  A_lo x W_hi  -- the kernels are bound by the shared-memory operand reads of these skinny MMAs, and
//    this ordering reads the 4 KB A tile twice instead of three times.  Five 96-column accumulators rotate through TMEM;
//    twelve epilogue warps drain them (ReLU, residual, BatchNorm statistics, planar and operand-format output), one warp
//    issues the MMAs, one warp issues the TMA copies the moment a ring slot is released.
//  * BatchNorm of the producer is folded into the consumer: into its weights for the forward (W' = W * rstd; the mean term
//    rides per tap on the "ones" channel 45, which reproduces zero padding of the NORMALISED tensor exactly, tc_fold_kernel)
//    and into the epilogue for the weight gradient.  Halo rows stay zero everywhere.
//  * weight gradient: 9 GEMMs  dW_tap[out, in] += sum_q dC[q, out] * X[q + shift, in]  with the 9 x 48 accumulator columns
//    resident in TMEM across all utterances of a CTA.  The dC tile of a K step is copied once from shared memory into tensor
//    memory (tcgen05.cp, from the channel-major copy dc_opT of the gradient) and feeds all 18 MMAs of the step as the A operand;
//    only the X tiles (MN-major, straight from the operand format) are read from shared memory per MMA.
#include "res8_common.cuh"
#include "tc_common.cuh"

#define TC_THREADS 320            // weight-gradient kernel: warps 0-7 epilogue, warp 8 MMA issue, warp 9 TMA loader
#define TS_THREADS 448            // stream kernel: warps 0-11 epilogue, warp 12 MMA issue, warp 13 TMA loader
#define TS_EPI_WARPS 12
#define TC_WORKERS 256
#define TC_PITCH 11
#define TC_PAD 12                 // largest |tap shift|
#define TC_N 48                   // channels padded to 48
#define TC_WBYTES (2 * R8TC_WBLOCK * 2)   // one layer's (hi | lo) weight operand
#define TC_SMEM_LIMIT (227 * 1024 - 512)

__host__ __device__ static inline int r8tc_dcop_rows_dev(int H) { return ((H + 2) * TC_PITCH + 63) & ~63; }
int r8tc_dcop_rows(int H) { return r8tc_dcop_rows_dev(H); }
size_t r8tc_dcop_bytes(int H) { return (size_t)12 * r8tc_dcop_rows(H) * 16; }

static size_t tc_stream_smem(int R) { return (size_t)TC_WBYTES + (size_t)12 * (2 * R + 2 * TC_PAD) * 16 + (48 * 3 + TS_EPI_WARPS * 2 * 16) * 4; }
// 3 transposed quarter slots (8-channel groups 144 B apart) + 3 raw quarter slots (planes padded by 16 B) + 2 X buffers + 1 KB overrun pad
static size_t tc_wgrad_smem(int R) { return (size_t)3 * ((R / 32) * 12 * 144) + (size_t)3 * 12 * ((R / 4) * 16 + 16) + 2 * (size_t)12 * (R + 2 * TC_PAD) * 16 + 1024; }
bool r8tc_supported(int H) {
  const int R = r8tc_dcop_rows(H);
  return H >= 1 && 2 * R / 128 * 96 <= 512 && tc_stream_smem(R) <= TC_SMEM_LIMIT && tc_wgrad_smem(R) <= TC_SMEM_LIMIT;
}

// =============================================================================================
// weight operands, one block per layer and direction:  bf16 [9 taps][6 chunks][96 n: 48 hi | 48 lo][8 k]
//   data gradient (tc_weight_prep): n = c, k = o, tap flipped (8 - tap), raw weights
//   forward (tc_fold)             : n = o, k = c, W' = W * rstd[c]  (BatchNorm of the producer folded in)
// =============================================================================================
__device__ __forceinline__ void tc_store_w(__nv_bfloat16* blk, int tap, int chunk, int n, int j, float v) {
  const __nv_bfloat16 hi = __float2bfloat16_rn(v);
  const size_t base = ((size_t)(tap * 6 + chunk) * 96) * 8 + j;
  blk[base + (size_t)n * 8] = hi;
  blk[base + (size_t)(48 + n) * 8] = __float2bfloat16_rn(v - __bfloat162float(hi));
}

__global__ void tc_weight_prep_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= R8_LAYERS * R8TC_WBLOCK) return;
  const int layer = i / R8TC_WBLOCK, r = i - layer * R8TC_WBLOCK;
  const int j = r & 7, n = (r >> 3) % TC_N, chunk = ((r >> 3) / TC_N) % 6, tap = (r >> 3) / (TC_N * 6);
  const int k = chunk * 8 + j;
  float v = 0.f;
  if (n < R8_C && k < R8_C) v = w[(size_t)layer * R8_KW + (k * R8_C + n) * 9 + (8 - tap)];
  tc_store_w(out + ((size_t)(layer * 2 + 1) * 2) * R8TC_WBLOCK, tap, chunk, n, j, v);
}

int r8tc_weight_prep(howl_ctx_t* ctx, cudaStream_t st, const float* w_layers, __nv_bfloat16* wprep) {
  const int n = R8_LAYERS * R8TC_WBLOCK;
  tc_weight_prep_kernel<<<(n + 255) / 256, 256, 0, st>>>(w_layers, wprep);
  HOWL_LAUNCHED(ctx, "tc_weight_prep");
  return HOWL_OK;
}

// forward fold: BatchNorm of the producer layer moves into the consumer's weights.  With xn = (u - mean) * rstd inside the
// image and 0 in the halo (zero padding of the NORMALISED tensor, as the reference does):
//   conv_W(xn)[q] = sum_tap sum_c W[o][c][tap] rstd[c] u[q+s] - sum_{tap: q+s inside} sum_c W[o][c][tap] rstd[c] mean[c]
// The second sum depends on which taps fall into the halo, i.e. on the pixel's border position -- exactly what the "ones"
// channel (input channel 45: 1 inside the image, 0 in the halo) reproduces when its weight is
//   W'[o][45][tap] = - sum_c W[o][c][tap] rstd[c] mean[c].
// So the operand halo stays zero and the kernel needs no bias, no border cases and no halo rewrite.
// grid = 9 taps; mean_rstd == null -> identity (layer 1 reads the un-normalised a0)
__global__ void tc_fold_kernel(const float* __restrict__ w, const float* __restrict__ mean_rstd, __nv_bfloat16* __restrict__ blk, const TcFoldBn bn) {
  __shared__ float s_mu[48], s_rs[48], s_bias[48];
  const int tid = threadIdx.x, tap = blockIdx.x;
  if (tid < 48) {
    float mu = 0.f, rs = tid < R8_C ? 1.f : 0.f;
    if (tid < R8_C && bn.stats) {
      // the producer's batch statistics are final (its kernel has completed): every CTA derives mean / rstd, CTA 0 publishes them
      const double mean = bn.stats[tid] / bn.count;
      double var = bn.stats[R8_C + tid] / bn.count - mean * mean;
      if (var < 0.0) var = 0.0;
      mu = (float)mean;
      rs = (float)(1.0 / sqrt(var + R8_BN_EPS));
      if (tap == 0) {
        bn.mean_rstd_out[tid] = mu;
        bn.mean_rstd_out[R8_C + tid] = rs;
        if (bn.running) {
          const double unbiased = bn.count > 1.0 ? var * bn.count / (bn.count - 1.0) : var;
          bn.running[tid] = (float)((1.0 - R8_BN_MOM) * bn.running[tid] + R8_BN_MOM * mean);
          bn.running[R8_C + tid] = (float)((1.0 - R8_BN_MOM) * bn.running[R8_C + tid] + R8_BN_MOM * unbiased);
        }
        if (tid == 0 && bn.nbt) *bn.nbt += 1;
      }
    } else if (tid < R8_C && mean_rstd) {
      mu = mean_rstd[tid];
      rs = mean_rstd[R8_C + tid];
    }
    s_mu[tid] = mu;
    s_rs[tid] = rs;
  }
  const bool folded = bn.stats || mean_rstd;
  // this tap's 45 x 45 weights: one round of independent (strided) global loads into shared memory, read twice from there
  __shared__ float s_w[R8_C * R8_C];
  for (int i = tid; i < R8_C * R8_C; i += blockDim.x) s_w[i] = w[(size_t)i * 9 + tap];
  __syncthreads();
  if (tid < 48) {
    double b = 0.0;
    if (tid < R8_C && folded)
      for (int c = 0; c < R8_C; ++c) b -= (double)s_w[tid * R8_C + c] * (double)s_rs[c] * (double)s_mu[c];
    s_bias[tid] = (float)b;
  }
  __syncthreads();
  for (int r = tid; r < 6 * 48 * 8; r += blockDim.x) {
    const int j = r & 7, n = (r >> 3) % TC_N, chunk = (r >> 3) / TC_N;
    const int c = chunk * 8 + j;
    float v = 0.f;
    if (n < R8_C && c < R8_C) v = s_w[n * R8_C + c] * s_rs[c];
    else if (n < R8_C && c == R8_C) v = s_bias[n];
    tc_store_w(blk, tap, chunk, n, j, v);
  }
}

int r8tc_fold(howl_ctx_t* ctx, cudaStream_t st, const float* w_layer, const float* mean_rstd, __nv_bfloat16* blk, const TcFoldBn* bn) {
  TcFoldBn none;
  memset(&none, 0, sizeof(none));
  tc_fold_kernel<<<9, 256, 0, st>>>(w_layer, mean_rstd, blk, bn ? *bn : none);
  HOWL_LAUNCHED(ctx, "tc_fold");
  return HOWL_OK;
}

// =============================================================================================
// stream kernel: forward (FWD) and data gradient of the 45->45 3x3 convolution
// =============================================================================================
struct TcStreamArgs {
  ConvParams p;                  // out (planar fp32 or null), stats, B, H
  const __nv_bfloat16* in_op;    // [B][2][6][R][8]
  __nv_bfloat16* out_op;         // forward: the output in operand format, or null
  const __nv_bfloat16* w;        // [9][6][96][8]
  // forward extras
  const __nv_bfloat16* res_op;   // residual (operand format) added after the ReLU, or null
  uint16_t* mask_out;            // [B][3 channel groups][R] ReLU decisions (bit jj = channel 16 * grp + jj), or null
  float* pooled_raw;             // [B][45] per-utterance spatial sums of the output (layer 6 -> head), or null
  // data gradient with the BatchNorm backward of the producer layer j fused into the epilogue (MODE 3):
  //   G = rstd * (g - m1 - xhat * m2) [+ gu_in];  gu_out = G;  dC = relu'(conv_j) ? G : 0  ->  dc_out (operand format, rows = pixels)
  const __nv_bfloat16* u_op;     // u_j in operand format (xhat and, for odd j, the ReLU decision u > 0)
  const float* bn_coef;          // [3][48]: rstd, A = -rstd m1 + mean rstd^2 m2, Bc = -rstd^2 m2   (G = rstd g + A + Bc u)
  const float* gu_in;            // planar [B,45,H,10] residual-path gradient flowing into u_j, or null
  float* gu_out;                 // planar G (even j), or null
  const uint16_t* mask_in;       // ReLU decisions of conv_j stored by the forward (even j), or null (odd j: u > 0)
  __nv_bfloat16* dc_out;
  int R;
  unsigned long long* prof;      // tuning aid: per-CTA cycle counters of the pipeline waits, or null
};

// 8 channels of one raster row: operand-format (hi, lo) pair -> fp32
__device__ __forceinline__ void tc_unpack8(const uint4& hi, const uint4& lo, float* v) {
  const uint32_t h[4] = {hi.x, hi.y, hi.z, hi.w}, l[4] = {lo.x, lo.y, lo.z, lo.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    v[2 * i] = __uint_as_float(h[i] << 16) + __uint_as_float(l[i] << 16);
    v[2 * i + 1] = __uint_as_float(h[i] & 0xffff0000u) + __uint_as_float(l[i] & 0xffff0000u);
  }
}

// 8 x 8 transpose among the 8 lanes of a raster-row group: in  d[j] = channel j of this lane's row,
// out d[j] = row j of the group for channel `sub` (= lane & 7)
__device__ __forceinline__ void tc_transpose8(float* d, int sub) {
#pragma unroll
  for (int s = 1; s < 8; s <<= 1) {
    const bool up = (sub & s) != 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if ((j & s) == 0) {
        const float send = up ? d[j] : d[j + s];
        const float got = __shfl_xor_sync(0xffffffffu, send, s);
        if (up) d[j] = got; else d[j + s] = got;
      }
    }
  }
}

// MODE 0: forward (eval), 1: forward + BatchNorm statistics (train), 2: plain data gradient (planar fp32 out; layer 1),
//      3: data gradient + fused BatchNorm backward / ReLU mask of the producer layer (operand-format dC out)
// SINGLE: fast mode, bf16 x bf16 only (the two low-order MMAs per product are skipped; conv_engine = 2)
template <int MODE, bool SINGLE>
__global__ void __launch_bounds__(TS_THREADS, 1) conv3x3_stream_tc_kernel(const TcStreamArgs a) {
  constexpr bool FWD = MODE <= 1;
  constexpr int STATS = MODE == 1 ? 1 : 0;
  extern __shared__ __align__(128) unsigned char smem[];
  const ConvParams& p = a.p;
  const int H = p.H, HW = H * R8_W, R = a.R, RS = 2 * R + 2 * TC_PAD, T = 2 * R / 128;
  const int j0_last = (R - 1) / 128;                 // last tile that reads slot 0; tiles > j0_last read slot 1 only
  const int j1_first = R / 128;                      // first tile that reads slot 1 (R % 128 != 0: it straddles both)
  uint4* w_s = reinterpret_cast<uint4*>(smem);                              // [9][6][96] x 16 B
  uint4* a_s = reinterpret_cast<uint4*>(smem + TC_WBYTES);                  // [hi 6 | lo 6][RS] x 16 B, ring row r at TC_PAD + r
  float* s_f = reinterpret_cast<float*>(smem + TC_WBYTES + (size_t)12 * RS * 16);
  float* s_coef = s_f;            // [3][48] BatchNorm-backward coefficients (MODE 3)
  float* s_red = s_f + 144;       // [12 warps][2][16]
  __shared__ __align__(8) uint64_t bar_w, bar_in[2], bar_tile[5], bar_tfree[5];
  __shared__ long long s_tload[2];
  __shared__ uint32_t s_tmem;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  if (warp == TS_EPI_WARPS) {
    tc::tmem_alloc<512>(&s_tmem);
    if (lane == 0) {
      tc::mbar_init(&bar_w, 1);
      for (int i = 0; i < 2; ++i) {
        tc::mbar_init(&bar_in[i], 1);
      }
      for (int i = 0; i < 5; ++i) {
        tc::mbar_init(&bar_tile[i], 1);
        tc::mbar_init(&bar_tfree[i], TS_EPI_WARPS);
      }
      tc::fence_barrier_init();
    }
  }
  if (tid < 144) s_coef[tid] = (MODE == 3) ? a.bn_coef[tid] : 0.f;
  // ring := 0 (a CTA with an odd utterance count multiplies a never-loaded slot into discarded rows); the zero pad rows in
  // front of and behind the ring stand for the neighbouring utterances' halo rows
  for (int i = tid; i < 12 * RS; i += TS_THREADS) a_s[i] = make_uint4(0, 0, 0, 0);
  tc::fence_proxy_async();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = s_tmem;
  const int64_t n_local = (p.B - blockIdx.x + gridDim.x - 1) / gridDim.x;
  const int64_t n_cycles = (n_local + 1) / 2;

  if (warp == TS_EPI_WARPS + 1) {
    // ================= TMA loader: refills a ring slot the moment its last reader tile has been multiplied =================
    // "Last reader" counts the tiles that own rows of the slot.  A neighbouring tile also reaches up to TC_PAD rows across the slot
    // boundary (tap shifts), i.e. into the first / last 12 raster rows of the other utterance -- halo rows, zero in every
    // utterance (and in the zero-initialised ring), so such a read may overlap the refill: it sees zeros before and after.
    if (tc::elect_one()) {
      const unsigned char* src0 = reinterpret_cast<const unsigned char*>(a.in_op);
      const size_t utt_bytes = (size_t)12 * R * 16;
      const __nv_bfloat16* side_op = FWD ? a.res_op : (MODE == 3 ? a.u_op : nullptr);
      const float* side = (MODE >= 2) ? a.gu_in : nullptr;
      auto load = [&](int64_t k) {            // utterance k of this CTA -> ring slot k & 1, one copy per 8-channel group
        const int s = (int)(k & 1);
        const int64_t b = blockIdx.x + k * (int64_t)gridDim.x;
        const unsigned char* src = src0 + (size_t)b * utt_bytes;
        if (a.prof) s_tload[s] = clock64();
        tc::mbar_expect_tx(&bar_in[s], (uint32_t)utt_bytes);
        for (int g = 0; g < 12; ++g)
          tc::tma_bulk_g2s(a_s + (size_t)g * RS + TC_PAD + s * R, src + (size_t)g * R * 16, (uint32_t)(R * 16), &bar_in[s]);
        // what the epilogue of this utterance will read (residual / BatchNorm-backward input): pull it into L2 now
        if (side_op) tc::l2_prefetch_bulk(reinterpret_cast<const unsigned char*>(side_op) + (size_t)b * utt_bytes, (uint32_t)utt_bytes);
        if (side) {
          const uintptr_t lo = reinterpret_cast<uintptr_t>(side + b * (int64_t)R8_C * HW);
          const uintptr_t lo16 = lo & ~(uintptr_t)15, hi16 = (lo + (uintptr_t)R8_C * HW * 4 + 15) & ~(uintptr_t)15;
          tc::l2_prefetch_bulk(reinterpret_cast<const void*>(lo16), (uint32_t)(hi16 - lo16));
        }
      };
      tc::mbar_expect_tx(&bar_w, TC_WBYTES);
      tc::tma_bulk_g2s(w_s, a.w, TC_WBYTES, &bar_w);
      if (n_local > 0) load(0);
      if (n_local > 1) load(1);
      for (int64_t c = 0; c < n_cycles; ++c) {
        const uint32_t par = (uint32_t)(c & 1);
        if (2 * c + 2 < n_local) {
          tc::mbar_wait(&bar_tile[j0_last], par);      // slot 0: tiles 0..j0_last of this cycle are done
          load(2 * c + 2);
        }
        if (2 * c + 3 < n_local) {
          tc::mbar_wait(&bar_tile[T - 1], par);        // slot 1: the cycle's last tile is done
          load(2 * c + 3);
        }
      }
    }
    __syncwarp();
  } else if (warp == TS_EPI_WARPS) {
    // ================= MMA issuer =================
    if (tc::elect_one()) {
      tc::mbar_wait(&bar_w, 0);
      const uint32_t idesc96 = tc::instr_desc_bf16(128, 96, 0, 0), idesc48 = tc::instr_desc_bf16(128, 48, 0, 0);
      const uint32_t b_lo = tc::desc_lo(tc::smem_u32(w_s), 96u * 16u);
      const uint32_t a_hi_s = tc::smem_u32(a_s), a_lo_s = a_hi_s + (uint32_t)(6 * RS * 16);
      const uint32_t ah_lo = tc::desc_lo(a_hi_s, (uint32_t)RS * 16u), al_lo = tc::desc_lo(a_lo_s, (uint32_t)RS * 16u);
      const uint32_t d_hi128 = tc::desc_hi(128u);
      unsigned long long pw[8] = {0, 0, 0, 0, 0, 0, 0, 0};
      const long long t_begin = clock64();
#define TS_WAIT(slot_, bar_, par_)                                  \
  do {                                                              \
    if (a.prof) {                                                   \
      const long long t0_ = clock64();                              \
      tc::mbar_wait(bar_, par_);                                    \
      const long long t1_ = clock64();                              \
      pw[slot_] += (unsigned long long)(t1_ - t0_);                 \
      if (slot_ < 2 && t1_ - t0_ > 200) {                           \
        pw[6] += (unsigned long long)(t1_ - s_tload[slot_]);        \
        pw[7] += 1;                                                 \
      }                                                             \
    } else {                                                        \
      tc::mbar_wait(bar_, par_);                                    \
    }                                                               \
  } while (0)
      for (int64_t c = 0; c < n_cycles; ++c) {
        const uint32_t par = (uint32_t)(c & 1);
        const bool has1 = 2 * c + 1 < n_local;
        const int jend = has1 ? T : j0_last + 1;
        for (int j = 0; j < jend; ++j) {
          if (j == 0) TS_WAIT(0, &bar_in[0], par);
          if (j == j1_first && has1) TS_WAIT(1, &bar_in[1], par);
          if (c > 0) TS_WAIT(3, &bar_tfree[j], par ^ 1u);       // epilogue of the previous cycle has drained this accumulator
          tc::fence_after_sync();
          const uint32_t d = tmem + (uint32_t)(96 * j);
          const uint32_t rowb = (uint32_t)(TC_PAD + 128 * j);
#pragma unroll 1
          for (int tap = 0; tap < 9; ++tap) {
            const int shift = (tap / 3 - 1) * TC_PITCH + (tap % 3 - 1);
#pragma unroll
            for (int ks = 0; ks < 3; ++ks) {
              const uint32_t aoff = (uint32_t)(2 * ks) * (uint32_t)RS + rowb + (uint32_t)shift;
              const uint64_t bd = tc::desc_make(b_lo + (uint32_t)((tap * 6 + 2 * ks) * 96), d_hi128);
              if (SINGLE) {
                tc::umma_bf16(d, tc::desc_make(ah_lo + aoff, d_hi128), bd, idesc48, (tap | ks) ? 1u : 0u);   // hi x hi only
              } else {
                tc::umma_bf16(d, tc::desc_make(ah_lo + aoff, d_hi128), bd, idesc96, (tap | ks) ? 1u : 0u);   // hi x (hi | lo)
                tc::umma_bf16(d, tc::desc_make(al_lo + aoff, d_hi128), bd, idesc48, 1u);                     // lo x hi
              }
            }
          }
          tc::umma_commit(&bar_tile[j]);
        }
      }
#undef TS_WAIT
      if (a.prof) {
        pw[5] = (unsigned long long)(clock64() - t_begin);
        for (int i = 0; i < 8; ++i) a.prof[blockIdx.x * 16 + i] = pw[i];
      }
    }
    __syncwarp();
  } else {
    // ================= epilogue: thread = one ring row, warp quad = 16 channels (two 8-channel groups) =================
    const int grp = warp >> 2;
    const bool full = grp < 2;            // the last quad owns channels 32..44 plus the ones channel and two pad channels
    float st1[16], st2[16];
#pragma unroll
    for (int c = 0; c < 16; ++c) st1[c] = st2[c] = 0.f;
    unsigned long long ep_wait = 0;
    const long long ep_begin = clock64();
    // layer 6: per-utterance spatial sums for the head.  A warp's 32-row block visits the rows of one utterance in consecutive
    // tiles, so one running set per thread is flushed (warp reduction + one atomic per channel) when the utterance changes.
    float pl[16];
    int64_t pl_b = -1;
#pragma unroll
    for (int c = 0; c < 16; ++c) pl[c] = 0.f;
    auto pool_flush = [&]() {
      if (pl_b >= 0) {
#pragma unroll
        for (int jj = 0; jj < 16; ++jj) {
          if (jj < 13 || full) {
            const float sum = warp_sum(pl[jj]);
            if (lane == 0) atomicAdd(a.pooled_raw + pl_b * R8_C + 16 * grp + jj, sum);
          }
          pl[jj] = 0.f;
        }
      }
    };
    for (int64_t cyc = 0; cyc < n_cycles; ++cyc) {
      const uint32_t par = (uint32_t)(cyc & 1);
      const bool has1 = 2 * cyc + 1 < n_local;
      const int jend = has1 ? T : j0_last + 1;
      for (int j = 0; j < jend; ++j) {
        const int r = 128 * j + 32 * (warp & 3) + lane;
        const int slot = r >= R ? 1 : 0, q = r - slot * R;
        const int64_t kk = 2 * cyc + slot;
        const bool live = kk < n_local;                       // warp uniform: R is a multiple of 64
        const int64_t b = blockIdx.x + kk * (int64_t)gridDim.x;
        const int y = q / TC_PITCH - 1, x = q % TC_PITCH - 1;
        const bool valid = live && (y >= 0) && (y < H) && (x >= 0) && (x < R8_W);
        // 32-bit element offset of (b, first channel of the quad, y, x); the host checks B * 45 * HW < 2^31
        const uint32_t off0 = valid ? (uint32_t)(b * R8_C + 16 * grp) * (uint32_t)HW + (uint32_t)(y * R8_W + x) : 0u;
        const uint32_t taddr = tmem + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)(96 * j + 16 * grp);
        const size_t op_row = (size_t)b * 12 * R + q;         // 16-byte units; chunk c of part p at (p * 6 + c) * R
        // ---- everything the tile's epilogue reads from HBM is requested before the accumulator wait
        uint4 sh[2], sl[2];                 // residual (forward) / u_j (fused BatchNorm backward): raw operand-format vectors
        float gin[16];
        uint32_t bits = 0;
#pragma unroll
        for (int cb = 0; cb < 2; ++cb) sh[cb] = sl[cb] = make_uint4(0, 0, 0, 0);
#pragma unroll
        for (int jj = 0; jj < 16; ++jj) gin[jj] = 0.f;
        const __nv_bfloat16* side_op = FWD ? a.res_op : (MODE == 3 ? a.u_op : nullptr);
        if (side_op != nullptr && valid) {
          const uint4* so = reinterpret_cast<const uint4*>(side_op) + op_row;
#pragma unroll
          for (int cb = 0; cb < 2; ++cb) {
            const int ch = 2 * grp + cb;
            sh[cb] = __ldg(so + (size_t)ch * R);
            sl[cb] = __ldg(so + (size_t)(6 + ch) * R);
          }
        }
        if (MODE >= 2 && valid) {
          if (a.gu_in) {
#pragma unroll
            for (int jj = 0; jj < 16; ++jj)
              if (jj < 13 || full) gin[jj] = __ldg(a.gu_in + off0 + (uint32_t)jj * (uint32_t)HW);
          }
          if (MODE == 3 && a.mask_in) bits = a.mask_in[((size_t)b * 3 + grp) * R + q];
        }
        const long long te0 = (a.prof && warp == 0) ? clock64() : 0;
        tc::mbar_wait(&bar_tile[j], par);
        if (a.prof && warp == 0) ep_wait += (unsigned long long)(clock64() - te0);
        tc::fence_after_sync();
        uint32_t mbits = 0;
        // two halves of 8 channels: keeps the live register set small (accumulator columns, side values, outputs of one half)
#pragma unroll
        for (int cb = 0; cb < 2; ++cb) {
          uint32_t v[8], v2[8];
          tc::tmem_ld8_nowait(taddr + 8 * cb, v);             // hi x hi + lo x hi
          if (!SINGLE) tc::tmem_ld8_nowait(taddr + 48 + 8 * cb, v2);   // hi x lo
          float pre[8], ov[8];
          tc_unpack8(sh[cb], sl[cb], pre);
          tc::tmem_ld_wait();
          const int ch = 2 * grp + cb;
#pragma unroll
          for (int j8 = 0; j8 < 8; ++j8) {
            const int jj = 8 * cb + j8;
            const bool chan = (jj < 13 || full);             // a real channel (0..44)
            float o = 0.f;
            if (valid && chan) {
              const float acc = __uint_as_float(v[j8]) + (SINGLE ? 0.f : __uint_as_float(v2[j8]));
              if (FWD) {
                if (acc > 0.f) mbits |= 1u << jj;
                o = fmaxf(acc, 0.f) + pre[j8];
                if (p.out) p.out[off0 + (uint32_t)jj * (uint32_t)HW] = o;
                if (STATS == 1) {
                  st1[jj] += o;
                  st2[jj] = fmaf(o, o, st2[jj]);
                }
              } else if (MODE == 2) {
                p.out[off0 + (uint32_t)jj * (uint32_t)HW] = acc + gin[jj];
              } else {
                // BatchNorm backward of the producer layer + residual fan-in + ReLU mask, straight from the accumulator
                const int c = 16 * grp + jj;
                const float u = pre[j8];
                const float G = fmaf(s_coef[c], acc, fmaf(s_coef[96 + c], u, s_coef[48 + c])) + gin[jj];
                if (a.gu_out) a.gu_out[off0 + (uint32_t)jj * (uint32_t)HW] = G;
                const bool on = a.mask_in ? ((bits >> jj) & 1u) != 0 : (u > 0.f);
                o = on ? G : 0.f;
              }
            } else if (FWD && valid && jj == 13) {
              o = 1.f;                      // channel 45: the "ones" channel (BatchNorm fold of the consumer and of its weight gradient)
            }
            ov[j8] = o;
          }
          if (live && MODE != 2) {            // halo rows are written too (zeros): no memset of the operand tensors
            uint4 hi, lo;
            if (FWD) {
              if (a.out_op) {
                uint4* o_op = reinterpret_cast<uint4*>(a.out_op) + op_row;
                tc::split8(ov, hi, lo);
                o_op[(size_t)ch * R] = hi;
                o_op[(size_t)(6 + ch) * R] = lo;
              }
              if (a.pooled_raw) {
                if (cb == 0 && b != pl_b) {      // warp uniform: a 32-row block never straddles two utterances
                  pool_flush();
                  pl_b = b;
                }
#pragma unroll
                for (int j8 = 0; j8 < 8; ++j8)
                  if (full || 8 * cb + j8 != 13) pl[8 * cb + j8] += ov[j8];      // (the ones channel of the last quad is not a channel)
              }
            } else {
              uint4* o_op = reinterpret_cast<uint4*>(a.dc_out) + op_row;
              tc::split8(ov, hi, lo);
              o_op[(size_t)ch * R] = hi;
              o_op[(size_t)(6 + ch) * R] = lo;
            }
          }
        }
        if (FWD && live && a.mask_out) a.mask_out[((size_t)b * 3 + grp) * R + q] = (uint16_t)mbits;
        tc::fence_before_sync();
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(&bar_tfree[j]);
      }
    }
    if (FWD && a.pooled_raw) pool_flush();
    if (a.prof && tid == 0) {
      a.prof[blockIdx.x * 16 + 8] = ep_wait;
      a.prof[blockIdx.x * 16 + 9] = (unsigned long long)(clock64() - ep_begin);
    }
    if (STATS) {
#pragma unroll
      for (int jj = 0; jj < 16; ++jj) {
        const float a1 = warp_sum(st1[jj]), a2 = warp_sum(st2[jj]);
        if (lane == 0) {
          s_red[(warp * 2 + 0) * 16 + jj] = a1;
          s_red[(warp * 2 + 1) * 16 + jj] = a2;
        }
      }
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (STATS && tid < 2 * R8_C) {
    const int which = tid / R8_C, c = tid - which * R8_C, grp = c / 16, jj = c - 16 * grp;
    double s = 0.0;
    for (int w = 4 * grp; w < 4 * grp + 4; ++w) s += (double)s_red[(w * 2 + which) * 16 + jj];
    atomicAdd(&p.stats[tid], s);
  }
  if (warp == TS_EPI_WARPS) tc::tmem_dealloc<512>(tmem);
}

int r8tc_conv(howl_ctx_t* ctx, cudaStream_t st, const TcConvCall& c) {
  TcStreamArgs a;
  memset(&a, 0, sizeof(a));
  a.p.B = c.B; a.p.H = c.H; a.p.out = c.out_planar; a.p.stats = c.stats;
  a.in_op = c.in_op; a.out_op = c.out_op; a.w = c.w;
  a.res_op = c.res_op; a.mask_out = c.mask_out; a.pooled_raw = c.pooled_raw;
  a.u_op = c.u_op; a.bn_coef = c.bn_coef; a.gu_in = c.gu_in; a.gu_out = c.gu_out; a.mask_in = c.mask_in;
  a.dc_out = c.dc_out;
  HOWL_REQUIRE(ctx, c.B * (int64_t)R8_C * c.H * R8_W < ((int64_t)1 << 31), HOWL_E_UNSUPPORTED,
               "tensor-core conv: batch of %lld utterances exceeds the 32-bit element offsets", (long long)c.B);
  HOWL_REQUIRE(ctx, c.in_op && c.w, HOWL_E_INVALID, "tensor-core conv: null operand");
  HOWL_REQUIRE(ctx, c.mode != 1 || c.stats, HOWL_E_INVALID, "tensor-core conv: statistics buffer missing");
  HOWL_REQUIRE(ctx, c.mode != 2 || c.out_planar, HOWL_E_INVALID, "tensor-core conv: output missing");
  HOWL_REQUIRE(ctx, c.mode != 3 || (c.u_op && c.bn_coef && c.dc_out && c.dc_out != c.in_op), HOWL_E_INVALID,
               "tensor-core conv: fused BatchNorm-backward arguments missing");
  a.R = r8tc_dcop_rows(c.H);
  const bool fwd = c.mode <= 1;
  a.prof = (ctx->tc_prof && ctx->tc_prof_kind == (fwd ? 1 : 2)) ? ctx->tc_prof : nullptr;
  HOWL_REQUIRE(ctx, r8tc_supported(c.H), HOWL_E_UNSUPPORTED, "tensor-core conv: H=%d does not fit", c.H);
  const size_t smem = tc_stream_smem(a.R);
  const int grid = (int)(c.B < ctx->sm_count ? c.B : ctx->sm_count);
  const bool single = ctx->conv_engine == 2;
#define TS_LAUNCH1(MODE_, SINGLE_)                                                                                       \
  do {                                                                                                                   \
    HOWL_CUDA(ctx, cudaFuncSetAttribute(conv3x3_stream_tc_kernel<MODE_, SINGLE_>,                                        \
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));                        \
    conv3x3_stream_tc_kernel<MODE_, SINGLE_><<<grid, TS_THREADS, smem, st>>>(a);                                         \
  } while (0)
#define TS_LAUNCH(MODE_)                  \
  do {                                    \
    if (single) TS_LAUNCH1(MODE_, true);  \
    else TS_LAUNCH1(MODE_, false);        \
  } while (0)
  if (c.mode == 0) TS_LAUNCH(0);
  else if (c.mode == 1) TS_LAUNCH(1);
  else if (c.mode == 2) TS_LAUNCH(2);
  else if (c.mode == 3) TS_LAUNCH(3);
  else HOWL_REQUIRE(ctx, false, HOWL_E_INVALID, "tensor-core conv: unsupported mode");
#undef TS_LAUNCH
#undef TS_LAUNCH1
  HOWL_LAUNCHED(ctx, fwd ? "conv3x3_fwd_tc" : "conv3x3_dgrad_tc");
  return HOWL_OK;
}

// =============================================================================================
// Weight gradient:  dW_tap[o][c] = sum_q dC[q][o] * X[q + shift_tap][c],  9 taps x 48 accumulator columns resident in TMEM
// across all utterances of a CTA.  The kernel is bound by operand reads, and the A tile (dC, 128 x 16) is the same for all
// nine taps and both halves of X -- so it is copied ONCE per 16-row K step from shared memory into tensor memory
// (tcgen05.cp) and the 18 MMAs of the step read A from there; only the 48 x 16 X tiles stream out of shared memory.
//   A rows (TMEM lanes): [dC_hi (48) | dC_lo (48) | 32 don't-care]  from dc_opT (rows = channels, K-major), one 128 x 256 bit copy
//   B: X_lo then X_hi, MN-major straight from the operand-format activations (rows = raster positions, 12 guard rows)
// so the four partial products land in rows o and 48 + o, which the epilogue adds.  Both operands arrive by TMA (a loader
// warp refills the buffers the moment their last MMA has retired): X is double buffered per utterance, dC streams through a ring
// of six quarter-utterance slots -- with the A tile out of the way an utterance is multiplied in ~9k cycles, about one TMA
// round trip under load, so the operands have to be in flight well over an utterance ahead.
// BatchNorm of X is folded into the epilogue through the "ones" channel (column 45 of every tap):
//     dW[o][c] = rstd[c] * ( sum_q dC[q][o] X[q+s][c]  -  mean[c] * sum_q dC[q][o] 1[q+s] )
// =============================================================================================
struct TcWgradArgs {
  const __nv_bfloat16* dc_op;   // conv-output gradient in operand format (rows = raster positions); transposed in shared memory here
  const __nv_bfloat16* x_op;
  const float* x_mean;   // or null (layer 1: X = a0, no normalisation)
  const float* x_rstd;
  float* dw;
  float* dones;          // [45][9]: sum_q dC[q][o] * 1[q + shift inside the image] -- the raw ones column (BatchNorm-backward statistics), or null
  int64_t B;
  int R;
};
#define TW_ASLOTS 8          // ring of A tiles in tensor memory behind the 9 x 48 accumulator columns
#define TW_DSLOTS 3          // shared-memory ring of TRANSPOSED dC quarter-utterances (A tiles for tcgen05.cp)
#define TW_RSLOTS 3          // raw (as in HBM) dC quarter-utterances, landed by TMA and transposed by the worker warps
#define TW_GSTRIDE 144       // bytes between the 8-channel groups of a transposed tile (128 + 16: the transposers' stores spread over all banks)
#define TW_TWARPS 4          // worker warps doing the transposes

template <bool SINGLE>   // SINGLE: fast mode, the X_lo MMA is skipped
__global__ void __launch_bounds__(TC_THREADS, 1) conv3x3_wgrad_tc_kernel(const TcWgradArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int R = a.R, Rx = R + 2 * TC_PAD, Rq = R / 4;            // a quarter utterance = Rq raster rows = Rq / 16 K steps
  const uint32_t x_bytes = (uint32_t)(12 * Rx * 16), u_bytes = (uint32_t)(12 * R * 16), q_bytes = u_bytes / 4;
  const uint32_t t_lbo = 12u * TW_GSTRIDE, t_bytes = (uint32_t)(Rq / 8) * t_lbo;      // transposed quarter: [Rq / 8 row groups][12 channel groups x 144 B]
  const uint32_t r_plane = (uint32_t)Rq * 16u + 16u, r_bytes = 12u * r_plane;          // raw quarter: 12 planes of Rq rows, 16 B of padding each
  unsigned char* x_buf = smem;                                   // 2 x [hi 6 | lo 6][Rx] x 16 B, raster row q at row q + 12
  unsigned char* r_ring = x_buf + 2 * (size_t)x_bytes;           // TW_RSLOTS raw quarters, as in HBM
  unsigned char* d_ring = r_ring + (size_t)TW_RSLOTS * r_bytes;  // TW_DSLOTS transposed quarters; the 128-lane copy of a tile reads 4 channel groups
                                                                 // (576 B) past the 12 real ones: don't-care lanes, inside the 1 KB pad at the end
  __shared__ __align__(8) uint64_t bar_x[2], bar_d[TW_DSLOTS], bar_free[TW_DSLOTS], bar_raw[TW_RSLOTS], bar_rawfree[TW_RSLOTS], bar_xfree[2];
  __shared__ uint32_t s_tmem;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  if (warp == 8) {
    tc::tmem_alloc<512>(&s_tmem);
    if (lane == 0) {
      for (int i = 0; i < 2; ++i) {
        tc::mbar_init(&bar_x[i], 1);
        tc::mbar_init(&bar_xfree[i], 1);
      }
      for (int i = 0; i < TW_DSLOTS; ++i) {
        tc::mbar_init(&bar_d[i], TW_TWARPS);
        tc::mbar_init(&bar_free[i], 1);
      }
      for (int i = 0; i < TW_RSLOTS; ++i) {
        tc::mbar_init(&bar_raw[i], 1);
        tc::mbar_init(&bar_rawfree[i], TW_TWARPS);
      }
      tc::fence_barrier_init();
    }
  }
  for (int i = tid; i < 2 * 12 * Rx; i += TC_THREADS) reinterpret_cast<uint4*>(x_buf)[i] = make_uint4(0, 0, 0, 0);   // guard rows
  tc::fence_proxy_async();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = s_tmem;
  const int64_t n_local = (a.B - blockIdx.x + gridDim.x - 1) / gridDim.x;
  const int64_t n_quarters = 4 * n_local;

  if (warp == 9) {
    // ================= TMA loader 1: raw dC quarters (12 row slices of the operand format) through TW_RSLOTS slots =================
    if (tc::elect_one() && n_local > 0) {
      const unsigned char* dsrc = reinterpret_cast<const unsigned char*>(a.dc_op);
      const uint32_t slice = (uint32_t)Rq * 16u;
      for (int64_t g = 0; g < n_quarters; ++g) {
        const int rs = (int)(g % TW_RSLOTS);
        if (g >= TW_RSLOTS) tc::mbar_wait(&bar_rawfree[rs], (uint32_t)(((g / TW_RSLOTS) - 1) & 1));
        const int64_t b = blockIdx.x + (g >> 2) * (int64_t)gridDim.x;
        tc::mbar_expect_tx(&bar_raw[rs], q_bytes);
        for (uint32_t pc = 0; pc < 12; ++pc)
          tc::tma_bulk_g2s(r_ring + (size_t)rs * r_bytes + (size_t)pc * r_plane, dsrc + (size_t)b * u_bytes + ((size_t)pc * R + (size_t)(g & 3) * Rq) * 16, slice,
                           &bar_raw[rs]);
      }
    }
    __syncwarp();
  } else if (warp == 7) {
    // ================= TMA loader 2: X, double buffered per utterance; a buffer is refilled once the MMAs of its utterance retired =================
    if (tc::elect_one() && n_local > 0) {
      const unsigned char* xsrc = reinterpret_cast<const unsigned char*>(a.x_op);
      for (int64_t k = 0; k < n_local; ++k) {
        if (k >= 2) tc::mbar_wait(&bar_xfree[k & 1], (uint32_t)(((k >> 1) - 1) & 1));
        const int64_t b = blockIdx.x + k * (int64_t)gridDim.x;
        tc::mbar_expect_tx(&bar_x[k & 1], u_bytes);
        for (uint32_t g = 0; g < 12; ++g)
          tc::tma_bulk_g2s(x_buf + (size_t)(k & 1) * x_bytes + ((size_t)g * Rx + TC_PAD) * 16,
                           xsrc + (size_t)b * u_bytes + (size_t)g * R * 16, (uint32_t)(R * 16), &bar_x[k & 1]);
      }
    }
    __syncwarp();
  } else if (warp < TW_TWARPS) {
    // ================= transposers: raw quarter [plane][row][8 channels] -> A-tile layout [row group][hi 48 | lo 48 channels][8 rows] =================
    // (an 8 x 8 bf16 transpose per (plane, row group) in registers: byte permutes pair the rows, the 32-bit words are then renamed)
    const int nblk = 12 * (Rq / 8);
    for (int64_t g = 0; g < n_quarters; ++g) {
      const int rs = (int)(g % TW_RSLOTS), slot = (int)(g % TW_DSLOTS);
      tc::mbar_wait(&bar_raw[rs], (uint32_t)((g / TW_RSLOTS) & 1));
      if (g >= TW_DSLOTS) tc::mbar_wait(&bar_free[slot], (uint32_t)(((g / TW_DSLOTS) - 1) & 1));      // MMAs that read the slot's previous tile retired
      tc::fence_after_sync();
      const unsigned char* raw = r_ring + (size_t)rs * r_bytes;
      unsigned char* dst = d_ring + (size_t)slot * t_bytes;
      for (int blk = tid; blk < nblk; blk += 32 * TW_TWARPS) {
        const int rg = blk / 12, pc = blk - rg * 12;     // plane fastest: neighbouring lanes are 1296 B (loads) / 144 B (stores) apart -> no bank conflicts
        uint4 r[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) r[i] = *reinterpret_cast<const uint4*>(raw + (size_t)pc * r_plane + (size_t)(rg * 8 + i) * 16);
        const uint32_t* w = reinterpret_cast<const uint32_t*>(r);       // w[4 * row + k] = channels (2k, 2k + 1) of the row
        uint4 o[8];
        uint32_t* ow = reinterpret_cast<uint32_t*>(o);                  // ow[4 * channel + i] = rows (2i, 2i + 1) of the channel
#pragma unroll
        for (int k = 0; k < 4; ++k)
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const uint32_t lo = w[4 * (2 * i) + k], hi = w[4 * (2 * i + 1) + k];
            ow[4 * (2 * k) + i] = __byte_perm(lo, hi, 0x5410);
            ow[4 * (2 * k + 1) + i] = __byte_perm(lo, hi, 0x7632);
          }
        uint4* d = reinterpret_cast<uint4*>(dst + (size_t)rg * t_lbo + (size_t)pc * TW_GSTRIDE);   // plane pc = part * 6 + chunk -> channel group pc
#pragma unroll
        for (int c = 0; c < 8; ++c) d[c] = o[c];
      }
      tc::fence_proxy_async();       // the tiles are read by tcgen05.cp (async proxy)
      __syncwarp();
      if (lane == 0) {
        tc::mbar_arrive(&bar_d[slot]);
        tc::mbar_arrive(&bar_rawfree[rs]);
      }
    }
  } else if (warp == 8) {
    // ================= MMA issuer =================
    if (tc::elect_one() && n_local > 0) {
      const uint32_t idesc = tc::instr_desc_bf16(128, TC_N, 0, 1);   // A K-major from tensor memory, B MN-major (K = raster rows)
      const uint32_t d_s = tc::smem_u32(d_ring), x_s = tc::smem_u32(x_buf);
      const uint32_t cp_hi = tc::desc_hi((uint32_t)TW_GSTRIDE);        // 8-channel groups 144 B apart; the two 8-row K groups of a tile t_lbo apart
      const uint32_t b_hi = tc::desc_hi((uint32_t)Rx * 16u);
      const uint32_t a_tmem0 = tmem + 9u * TC_N;
      uint32_t step = 0;
      for (int64_t g = 0; g < n_quarters; ++g) {
        const int64_t k = g >> 2;
        const int qi = (int)(g & 3), slot = (int)(g % TW_DSLOTS);
        const uint32_t xh_s = x_s + (uint32_t)(k & 1) * x_bytes;
        const uint32_t bh_lo = tc::desc_lo(xh_s, 128u), bl_lo = tc::desc_lo(xh_s + (uint32_t)(6 * Rx * 16), 128u);
        const uint32_t cp_lo = tc::desc_lo(d_s + (uint32_t)slot * t_bytes, t_lbo);
        if (qi == 0) tc::mbar_wait(&bar_x[k & 1], (uint32_t)((k >> 1) & 1));
        tc::mbar_wait(&bar_d[slot], (uint32_t)((g / TW_DSLOTS) & 1));
        tc::fence_after_sync();
        // the copy of K step i + 1 is issued ahead of the MMAs of step i (tcgen05 ops retire in issue order; the ring of A
        // tiles keeps a copy from overwriting a tile whose MMAs are still queued)
        tc::tmem_cp_128x256b(a_tmem0 + 8u * (step % TW_ASLOTS), tc::desc_make(cp_lo, cp_hi));
#pragma unroll 1
        for (int kq = 0; kq < Rq; kq += 16, ++step) {
          const uint32_t a_t = a_tmem0 + 8u * (step % TW_ASLOTS);
          if (kq + 16 < Rq)
            tc::tmem_cp_128x256b(a_tmem0 + 8u * ((step + 1) % TW_ASLOTS), tc::desc_make(cp_lo + (uint32_t)((kq + 16) >> 3) * (t_lbo >> 4), cp_hi));
          const uint32_t acc = step ? 1u : 0u;
          const int k0 = qi * Rq + kq;
#pragma unroll
          for (int tap = 0; tap < 9; ++tap) {
            const int shift = (tap / 3 - 1) * TC_PITCH + (tap % 3 - 1);
            const uint32_t d = tmem + (uint32_t)(tap * TC_N);
            const uint32_t boff = (uint32_t)(TC_PAD + shift + k0);
            if (!SINGLE) tc::umma_bf16_ts(d, a_t, tc::desc_make(bl_lo + boff, b_hi), idesc, acc);
            tc::umma_bf16_ts(d, a_t, tc::desc_make(bh_lo + boff, b_hi), idesc, SINGLE ? acc : 1u);
          }
        }
        tc::umma_commit(&bar_free[slot]);
        if (qi == 3) tc::umma_commit(&bar_xfree[k & 1]);      // the utterance's X buffer may be refilled
      }
      tc::mbar_wait(&bar_free[(n_quarters - 1) % TW_DSLOTS], (uint32_t)(((n_quarters - 1) / TW_DSLOTS) & 1));
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
        if (a.dones) atomicAdd(&a.dones[o * 9 + tap], ones);
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

int r8tc_wgrad(howl_ctx_t* ctx, cudaStream_t st, const __nv_bfloat16* dc_op, const __nv_bfloat16* x_op, const float* x_mean,
               const float* x_rstd, float* dw, float* dones, int64_t B, int H) {
  TcWgradArgs a;
  a.dc_op = dc_op; a.x_op = x_op; a.x_mean = x_mean; a.x_rstd = x_rstd; a.dw = dw; a.dones = dones; a.B = B;
  a.R = r8tc_dcop_rows(H);
  HOWL_REQUIRE(ctx, r8tc_supported(H), HOWL_E_UNSUPPORTED, "tensor-core wgrad: H=%d does not fit", H);
  const size_t smem = tc_wgrad_smem(a.R);
  const int grid = (int)(B < ctx->sm_count ? B : ctx->sm_count);
  if (ctx->conv_engine == 2) {
    HOWL_CUDA(ctx, cudaFuncSetAttribute(conv3x3_wgrad_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    conv3x3_wgrad_tc_kernel<true><<<grid, TC_THREADS, smem, st>>>(a);
  } else {
    HOWL_CUDA(ctx, cudaFuncSetAttribute(conv3x3_wgrad_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    conv3x3_wgrad_tc_kernel<false><<<grid, TC_THREADS, smem, st>>>(a);
  }
  HOWL_LAUNCHED(ctx, "conv3x3_wgrad_tc");
  return HOWL_OK;
}

// =============================================================================================
// conv0 (1 -> 45, 3x3, pad 1) + ReLU + AvgPool(3, 4) on the tensor cores.
// One output pixel of the pooled map averages 12 pre-pool pixels, each a 9-tap dot product: as a GEMM per sub-position t = (ty, tx) of the
// pooling window,  D_t[pp][oc] = sum_k A_t[pp][k] W[oc][k],  rows = 128 pooled pixels of the flattened (utterance, pixel) stream, K = 9 taps
// padded to 16, N = 48; fp32 via the same bf16 split as the 45 -> 45 layers (hi*hi + lo*hi + hi*lo).  The A tiles (an im2col of the 5 x 6
// feature patch of every pooled pixel) are BUILT in shared memory by the worker warps from the fp32 features -- 13 KB per utterance in HBM
// instead of a 583 KB pre-pool tensor -- the epilogue applies the ReLU, adds the 12 sub-positions, keeps their ReLU decisions as the 12 bits
// conv0_bwd needs, and writes a0 straight in operand format.  Two CTAs per SM (256 TMEM columns each = four sub-positions per pass) overlap
// one CTA's tile building / epilogue with the other's MMAs.
// =============================================================================================
#define C0T_THREADS 288          // warps 0-7 build + epilogue, warp 8 issues the MMAs
#define C0T_A_BYTES (12 * 2 * 2 * 128 * 16)      // [12 sub-positions][hi, lo][2 chunks of 8 taps][128 rows][8 x bf16]
#define C0T_W_BYTES (2 * 2 * 48 * 16)            // [hi, lo][2 chunks][48 rows][8 x bf16]

struct Conv0TcArgs {
  const float* feats;      // [B, F, 40]
  const float* w0;         // [45][9]
  uint4* a0_op;            // [B][12][R][8 bf16]
  uint16_t* bits0;         // [B][45][HW] or null
  int64_t B;
  int F, H, R;
};

__global__ void __launch_bounds__(C0T_THREADS, 2) conv0_tc_kernel(const Conv0TcArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  unsigned char* a_s = smem;
  unsigned char* w_s = smem + C0T_A_BYTES;
  __shared__ __align__(8) uint64_t bar_acc;
  __shared__ uint32_t s_tmem;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int HW = a.H * R8_W;
  const int64_t P = a.B * HW, tiles = (P + 127) / 128;
  if (warp == 8) {
    tc::tmem_alloc<256>(&s_tmem);
    if (lane == 0) {
      tc::mbar_init(&bar_acc, 1);
      tc::fence_barrier_init();
    }
  }
  // weight operand: W[oc][k] (k < 9), hi / lo
  for (int i = tid; i < 2 * 48; i += C0T_THREADS) {
    const int chunk = i / 48, n = i - chunk * 48;
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int k = chunk * 8 + j;
      v[j] = (n < R8_C && k < 9) ? a.w0[n * 9 + k] : 0.f;
    }
    uint4 hi, lo;
    tc::split8(v, hi, lo);
    reinterpret_cast<uint4*>(w_s)[chunk * 48 + n] = hi;
    reinterpret_cast<uint4*>(w_s)[2 * 48 + chunk * 48 + n] = lo;
  }
  tc::fence_proxy_async();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = s_tmem;
  const int row = tid & 127, half = (tid >> 7) & 1;      // builders: two threads per row, six sub-positions each; epilogue: 24 channels each
  uint32_t n_acc = 0;
  for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const int64_t gp = tile * 128 + row;
    const bool valid = gp < P;
    const int64_t b = valid ? gp / HW : 0;
    const int pp = valid ? (int)(gp - b * HW) : 0, h = pp / R8_W, w = pp - h * R8_W;
    if (warp < 8) {
      // ---- build the A tiles of this thread's six sub-positions from the 5 x 6 feature patch (rows 3h-1.., columns 4w-1..; zero outside)
      float patch[5][6];
      const float* f = a.feats + b * (int64_t)a.F * R8_MELS;
#pragma unroll
      for (int r = 0; r < 5; ++r) {
        const int y = 3 * h + r - 1;
#pragma unroll
        for (int c = 0; c < 6; ++c) {
          const int x = 4 * w + c - 1;
          patch[r][c] = (valid && y >= 0 && y < a.F && x >= 0 && x < R8_MELS) ? __ldg(f + (int64_t)y * R8_MELS + x) : 0.f;
        }
      }
      uint4* A = reinterpret_cast<uint4*>(a_s);
#pragma unroll
      for (int t = 0; t < 12; ++t) {
        if (t / 6 != half) continue;           // warp uniform; keeps every patch index a compile-time constant (registers, not local memory)
        const int ty = t >> 2, tx = t & 3;
        float v0[8], v1[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) v0[k] = patch[ty + k / 3][tx + k % 3];
#pragma unroll
        for (int k = 0; k < 8; ++k) v1[k] = 0.f;
        v1[0] = patch[ty + 2][tx + 2];
        uint4 hi, lo;
        tc::split8(v0, hi, lo);
        A[((t * 2 + 0) * 2 + 0) * 128 + row] = hi;
        A[((t * 2 + 1) * 2 + 0) * 128 + row] = lo;
        tc::split8(v1, hi, lo);
        A[((t * 2 + 0) * 2 + 1) * 128 + row] = hi;
        A[((t * 2 + 1) * 2 + 1) * 128 + row] = lo;
      }
      tc::fence_proxy_async();
    }
    float sum[24];
    uint32_t bits[12];           // ReLU decisions: channel pair (2i, 2i + 1) in the low / high 16 bits
#pragma unroll
    for (int j = 0; j < 24; ++j) sum[j] = 0.f;
#pragma unroll
    for (int j = 0; j < 12; ++j) bits[j] = 0u;
    for (int pass = 0; pass < 3; ++pass) {
      tc::fence_before_sync();
      __syncthreads();           // pass 0: the tiles are built; later passes: the previous pass's accumulators are drained
      tc::fence_after_sync();
      if (warp == 8) {
        if (tc::elect_one()) {
          const uint32_t idesc = tc::instr_desc_bf16(128, 48, 0, 0), hi128 = tc::desc_hi(128u);
          const uint32_t sa = tc::smem_u32(a_s), sw = tc::smem_u32(w_s);
          const uint64_t wh = tc::desc_make(tc::desc_lo(sw, 48u * 16u), hi128), wl = tc::desc_make(tc::desc_lo(sw + 2u * 48u * 16u, 48u * 16u), hi128);
#pragma unroll
          for (int tt = 0; tt < 4; ++tt) {
            const int t = pass * 4 + tt;
            const uint64_t ah = tc::desc_make(tc::desc_lo(sa + (uint32_t)(t * 2 + 0) * 4096u, 2048u), hi128);
            const uint64_t al = tc::desc_make(tc::desc_lo(sa + (uint32_t)(t * 2 + 1) * 4096u, 2048u), hi128);
            const uint32_t d = tmem + (uint32_t)(tt * 48);
            tc::umma_bf16(d, ah, wh, idesc, 0u);
            tc::umma_bf16(d, al, wh, idesc, 1u);
            tc::umma_bf16(d, ah, wl, idesc, 1u);
          }
          tc::umma_commit(&bar_acc);
        }
        __syncwarp();
      } else {
        tc::mbar_wait(&bar_acc, n_acc & 1);
        tc::fence_after_sync();
        const uint32_t taddr = tmem + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)(24 * half);
#pragma unroll
        for (int tp = 0; tp < 2; ++tp) {            // two sub-positions per wait: 48 accumulator values in flight
          uint32_t v[48];
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            tc::tmem_ld8_nowait(taddr + (2 * tp + u) * 48, v + 24 * u);
            tc::tmem_ld8_nowait(taddr + (2 * tp + u) * 48 + 8, v + 24 * u + 8);
            tc::tmem_ld8_nowait(taddr + (2 * tp + u) * 48 + 16, v + 24 * u + 16);
          }
          tc::tmem_ld_wait();
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            const int t = pass * 4 + 2 * tp + u;
#pragma unroll
            for (int j = 0; j < 24; ++j) {
              const float x = __uint_as_float(v[24 * u + j]);
              if (x > 0.f) {
                sum[j] += x;
                bits[j >> 1] |= 1u << (t + 16 * (j & 1));
              }
            }
          }
        }
      }
      ++n_acc;
    }
    if (warp < 8 && valid) {
      const int q = (h + 1) * TC_PITCH + (w + 1);
      uint4* op = a.a0_op + (size_t)b * 12 * a.R + q;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        float o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int oc = 24 * half + 8 * c + j;
          o[j] = oc < R8_C ? __fdiv_rn(sum[8 * c + j], 12.f) : (oc == R8_C ? 1.f : 0.f);      // channel 45: the ones channel
          if (a.bits0 && oc < R8_C) a.bits0[(b * R8_C + oc) * (int64_t)HW + pp] = (uint16_t)(bits[(8 * c + j) >> 1] >> (16 * (j & 1)));
        }
        uint4 hi, lo;
        tc::split8(o, hi, lo);
        const int chunk = 3 * half + c;
        op[(size_t)chunk * a.R] = hi;
        op[(size_t)(6 + chunk) * a.R] = lo;
        if (w == 0) {           // the shared halo column left of the image row (raster row q - 1)
          op[(size_t)chunk * a.R - 1] = make_uint4(0, 0, 0, 0);
          op[(size_t)(6 + chunk) * a.R - 1] = make_uint4(0, 0, 0, 0);
        }
      }
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 8) tc::tmem_dealloc<256>(tmem);
}

// halo rows of the operand-format a0 (raster rows outside the image) := 0
// zero the raster rows of a0_op the convolution kernel does not write (it writes the image rows and their shared halo column): the top
// halo row and everything from the bottom halo row to R -- two contiguous runs per plane; thread = one (utterance, plane, row)
__global__ void conv0_halo_kernel(uint4* __restrict__ a0_op, int64_t B, int H, int R) {
  const int bottom = (H + 1) * TC_PITCH, nh = TC_PITCH + (R - bottom);
  const int64_t n = B * 12 * nh;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int k = (int)(i % nh);
    const int64_t plane = i / nh;           // b * 12 + g
    const int q = k < TC_PITCH ? k : bottom + (k - TC_PITCH);
    a0_op[(size_t)plane * R + q] = make_uint4(0, 0, 0, 0);
  }
}

int r8tc_conv0(howl_ctx_t* ctx, cudaStream_t st, const float* feats, const float* w0, __nv_bfloat16* a0_op, uint16_t* bits0, int64_t B, int F,
               int H) {
  Conv0TcArgs a;
  a.feats = feats; a.w0 = w0; a.a0_op = reinterpret_cast<uint4*>(a0_op); a.bits0 = bits0; a.B = B; a.F = F; a.H = H; a.R = r8tc_dcop_rows(H);
  conv0_halo_kernel<<<ctx->sm_count * 4, 256, 0, st>>>(a.a0_op, B, H, a.R);
  HOWL_LAUNCHED(ctx, "conv0_halo");
  const size_t smem = C0T_A_BYTES + C0T_W_BYTES;
  HOWL_CUDA(ctx, cudaFuncSetAttribute(conv0_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int64_t tiles = (B * H * R8_W + 127) / 128;
  const int grid = (int)(tiles < 2 * ctx->sm_count ? tiles : 2 * ctx->sm_count);
  conv0_tc_kernel<<<grid, C0T_THREADS, smem, st>>>(a);
  HOWL_LAUNCHED(ctx, "conv0_tc");
  return HOWL_OK;
}

// =============================================================================================
// BatchNorm-backward statistics of layer j = i - 1 WITHOUT a pass over the gradient tensor: with g = dL/d(xn_j) the data gradient of
// conv_i,   sum_q g[q][c]        = sum_{o,tap} W_i[o][c][tap] * D1[o][tap]          D1 = the weight gradient's raw ones column
//           sum_q g[q][c] xhat_j = sum_{o,tap} W_i[o][c][tap] * dW_i[o][c][tap]     dW_i = the (BatchNorm-folded) weight gradient
// (substitute g = sum_{o,tap} W_i dC_i[q - s] and exchange the sums).  So the weight-gradient kernel of layer i, which runs before
// the data gradient, already holds both statistics, and the data-gradient epilogue can apply the BatchNorm backward on the fly.
// Output: the coefficients of  G = rstd g + A + Bc u   (= rstd (g - m1 - xhat m2)),  [3][48].   grid = 48 channels.
// =============================================================================================
__global__ void __launch_bounds__(128) bn_bwd_coef_kernel(const float* __restrict__ w, const float* __restrict__ dw,
                                                          const float* __restrict__ dones, const float* __restrict__ mean_rstd,
                                                          double count, float* __restrict__ coef) {
  const int c = blockIdx.x, tid = threadIdx.x;
  __shared__ double sh[2][4];
  double s1 = 0.0, s2 = 0.0;
  if (c < R8_C) {
    for (int i = tid; i < R8_C * 9; i += 128) {
      const int o = i / 9, tap = i - o * 9;
      const double wv = (double)w[(o * R8_C + c) * 9 + tap];
      s1 += wv * (double)dones[i];
      s2 += wv * (double)dw[(o * R8_C + c) * 9 + tap];
    }
  }
  s1 = warp_sum(s1);
  s2 = warp_sum(s2);
  if ((tid & 31) == 0) {
    sh[0][tid >> 5] = s1;
    sh[1][tid >> 5] = s2;
  }
  __syncthreads();
  if (tid == 0) {
    float rs = 0.f, A = 0.f, Bc = 0.f;
    if (c < R8_C) {
      const double m1 = (sh[0][0] + sh[0][1] + sh[0][2] + sh[0][3]) / count, m2 = (sh[1][0] + sh[1][1] + sh[1][2] + sh[1][3]) / count;
      const double mu = mean_rstd[c], r = mean_rstd[R8_C + c];
      rs = (float)r;
      A = (float)(-r * m1 + mu * r * r * m2);
      Bc = (float)(-r * r * m2);
    }
    coef[c] = rs;
    coef[48 + c] = A;
    coef[96 + c] = Bc;
  }
}

int r8tc_bn_bwd_coef(howl_ctx_t* ctx, cudaStream_t st, const float* w, const float* dw, const float* dones, const float* mean_rstd,
                     double count, float* coef) {
  bn_bwd_coef_kernel<<<48, 128, 0, st>>>(w, dw, dones, mean_rstd, count, coef);
  HOWL_LAUNCHED(ctx, "bn_bwd_coef");
  return HOWL_OK;
}

// =============================================================================================
// Head of the backward: BatchNorm backward + ReLU mask of layer 6, whose upstream gradient is the broadcast dh / HW (spatial mean).
// One thread = 8 channels of one RASTER position (halo positions compute zeros, so no memset).  Emits the conv-output gradient as
// both gradient operands, (hi, lo) bf16:
//   dc_op  [hi, lo][6 chunks][R rows][8 channels]               rows = pixels : A operand of the data gradient (K = channels)
//   dc_opT [R / 8 row groups][hi 48 | lo 48 channels][8 rows]   rows = channels: K-major A operand of the weight gradient
// and G itself (planar) as the residual-path gradient of layer 4.   grid.y = channel chunk.
// =============================================================================================
__global__ void __launch_bounds__(256) bn_bwd_head_op_kernel(const ApplyOpParams p) {
  const int HW = p.H * R8_W, R = r8tc_dcop_rows_dev(p.H), chunk = blockIdx.y;
  float mu[8], rs[8], m1[8], m2[8], live[8];
  int coff[8], cidx[8];
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
    cidx[j] = cc;
  }
  const float inv_hw = 1.f / (float)HW;
  const int64_t items = p.B * R;                     // multiple of 64: whole warps, aligned 8-lane row groups
  const uint4* uop = reinterpret_cast<const uint4*>(p.u_op);
  uint4* out = reinterpret_cast<uint4*>(p.dc_op);
  for (int64_t it = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; it < items; it += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = it / R;
    const int q = (int)(it - b * R);
    const int y = q / TC_PITCH - 1, x = q % TC_PITCH - 1;
    const bool valid = (y >= 0) && (y < p.H) && (x >= 0) && (x < R8_W);
    float d[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) d[j] = 0.f;
    const int64_t base = b * 12 * (int64_t)R + (int64_t)chunk * R + q;
    if (valid) {
      const int64_t ub = b * (int64_t)R8_C * HW + y * R8_W + x;
      float u[8];
      tc_unpack8(__ldg(uop + base), __ldg(uop + base + 6 * (int64_t)R), u);
      const uint32_t bits = (uint32_t)p.mask_bits[(b * 3 + (chunk >> 1)) * R + q] >> (8 * (chunk & 1));
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float g = __ldg(p.g_bcast + b * R8_C + cidx[j]) * inv_hw;
        const float G = rs[j] * (g - m1[j] - (u[j] - mu[j]) * rs[j] * m2[j]);
        if (live[j] != 0.f) p.gu_out[ub + coff[j]] = G;
        d[j] = ((bits >> j) & 1u) ? G * live[j] : 0.f;
      }
    }
    uint4 hi, lo;
    tc::split8(d, hi, lo);
    out[base] = hi;                                   // halo rows are written too (zeros): no memset of dc_op needed
    out[base + 6 * (int64_t)R] = lo;
  }
}

int r8tc_apply_head(howl_ctx_t* ctx, cudaStream_t st, const ApplyOpParams& p) {
  const int64_t items = p.B * r8tc_dcop_rows(p.H);
  int64_t bx = howl_ceil_div(items, 256);
  const int64_t cap = (int64_t)ctx->sm_count * 4;
  if (bx > cap) bx = cap;
  const dim3 grid((unsigned)bx, 6, 1);
  HOWL_REQUIRE(ctx, p.g_bcast && p.u_op && p.mask_bits && p.gu_out && p.dc_op, HOWL_E_INVALID, "apply_head: null argument");
  bn_bwd_head_op_kernel<<<grid, 256, 0, st>>>(p);
  HOWL_LAUNCHED(ctx, "bn_bwd_head_op");
  return HOWL_OK;
}

// test hook: ReLU decisions of the tensor-core engine as bytes [B,45,H,10]: the stored bits (even layers), or u > 0 (odd layers)
__global__ void tc_debug_mask_kernel(const __nv_bfloat16* __restrict__ u_op, const uint16_t* __restrict__ bits, uint8_t* __restrict__ mask,
                                     int64_t B, int H) {
  const int HW = H * R8_W, R = r8tc_dcop_rows_dev(H);
  const int64_t n = B * R8_C * HW;
  const uint4* uop = reinterpret_cast<const uint4*>(u_op);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int pix = (int)(i % HW), c = (int)((i / HW) % R8_C);
    const int64_t b = i / ((int64_t)HW * R8_C);
    const int q = (pix / R8_W + 1) * TC_PITCH + (pix % R8_W + 1);
    if (bits) {
      mask[i] = (bits[(b * 3 + c / 16) * R + q] >> (c % 16)) & 1;
    } else {
      float u[8];
      const int64_t base = b * 12 * (int64_t)R + (int64_t)(c / 8) * R + q;
      tc_unpack8(uop[base], uop[base + 6 * (int64_t)R], u);
      mask[i] = u[c % 8] > 0.f ? 1 : 0;
    }
  }
}

int r8tc_debug_mask(howl_ctx_t* ctx, cudaStream_t st, const __nv_bfloat16* u_op, const uint16_t* bits, uint8_t* mask, int64_t B, int H) {
  tc_debug_mask_kernel<<<ctx->sm_count * 8, 256, 0, st>>>(u_op, bits, mask, B, H);
  HOWL_LAUNCHED(ctx, "debug_mask");
  return HOWL_OK;
}
