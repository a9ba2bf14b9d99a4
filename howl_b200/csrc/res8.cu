// Res8 forward / backward for sm_100a -- exact-fp32 (FFMA) path.
//
// Reference semantics: howl/model/cnn.py:107-145 (Res8), nn.BatchNorm2d(affine=False) in train / eval mode,
// nn.CrossEntropyLoss(mean) and autograd's backward of all of it (training/run/train.py:292-301).
//
// Data layout in HBM (all fp32, caller-provided workspace):
//   feats  [B, F, 40]         time-major log-mel (K1 output; == the [B,1,F,40] tensor of cnn.py:128-129)
//   a0     [B, 45, H, 10]     relu(conv0) average-pooled (3,4);  H = F / 3
//   u_i    [B, 45, H, 10]     i = 1..6, the tensor fed to bn_i (relu(conv_i) [+ residual]); BatchNorm is never
//                             materialised: the consumer applies (x - mean) * rstd while staging its tile
//   g, dc, gu                 gradient ping-pong buffers of the same shape
// Kernels are persistent (one CTA per SM looping over utterances) so that weights stay in shared memory and
// per-channel statistics / weight gradients are reduced on chip before touching global atomics.
#include <math.h>

#include "common.cuh"

#include "res8_common.cuh"
#include "tc_common.cuh"
#include "../../include/howl_b200_debug.h"

// =============================================================================================
// workspace carve-up
// =============================================================================================
struct R8Ws {
  double* stats_fwd;   // [6][2][45]  sum(u), sum(u^2)
  double* stats_bwd;   // [6][2][45]  sum(g), sum(g * xhat)
  double* loss_acc;    // [1]
  float* mean_rstd;    // [6][2][45]
  float* wT;           // [6][18225]  transposed + flipped weights for dgrad
  __nv_bfloat16* wprep; // tensor-core weight operands (R8TC_WELEMS)
  float* pooled;       // [B,45]
  float* dh;           // [B,45]
  float* dlogits;      // [B,L]
  float* logits;       // [B,L]  copy of the forward's logits for the backward
  uint16_t* bits0;     // [B,45,H,10] the 12 ReLU decisions of every conv0 pooling window
  float* a0;
  float* u[R8_LAYERS];
  float* g;
  float* dc;
  float* gu[2];
  // ---- tensor-core engine (null when H is unsupported): everything the convolutions read is in operand format
  __nv_bfloat16* dc2;              // second conv-output-gradient buffer (the fused data gradient reads one and writes the other)
  __nv_bfloat16* uop[R8_LAYERS + 1];   // a0, u1..u6
  uint16_t* mask_bits[3];          // ReLU decisions of the residual layers 2, 4, 6: [B][3 channel groups][R]
  float* pooled_raw;               // [B,45] spatial sums of u6 (accumulated by the forward epilogue of layer 6)
  float* dones;                    // [6][45*9] raw ones columns of the weight gradients
  float* bn_coef;                  // [3][48] BatchNorm-backward coefficients of the layer being differentiated
  size_t bytes;
};

static R8Ws r8_carve(void* base, int64_t B, int H, int L) {
  R8Ws w;
  size_t off = 0;
  char* p = (char*)base;
  auto take = [&](size_t bytes) {
    void* r = p ? (void*)(p + off) : nullptr;
    off += howl_align_up(bytes, 256);
    return r;
  };
  w.stats_fwd = (double*)take(sizeof(double) * R8_LAYERS * 2 * R8_C);
  w.stats_bwd = (double*)take(sizeof(double) * R8_LAYERS * 2 * R8_C);
  w.loss_acc = (double*)take(sizeof(double) * 2);
  w.mean_rstd = (float*)take(sizeof(float) * R8_LAYERS * 2 * R8_C);
  w.wT = (float*)take(sizeof(float) * R8_LAYERS * R8_KW);
  w.wprep = (__nv_bfloat16*)take(sizeof(__nv_bfloat16) * R8TC_WELEMS);
  w.pooled = (float*)take(sizeof(float) * B * R8_C);
  w.dh = (float*)take(sizeof(float) * B * R8_C);
  w.dlogits = (float*)take(sizeof(float) * B * L);
  w.logits = (float*)take(sizeof(float) * B * L);
  const size_t n = sizeof(float) * (size_t)B * R8_C * H * R8_W;
  w.bits0 = (uint16_t*)take(n / 2);
  w.a0 = (float*)take(n);
  for (int i = 0; i < R8_LAYERS; ++i) w.u[i] = (float*)take(n);
  w.g = (float*)take(n);
  {   // dc doubles as the operand-format gradient of the tensor-core engine (r8tc_dcop_bytes per utterance)
    const size_t nop = (size_t)B * r8tc_dcop_bytes(H);
    w.dc = (float*)take(n > nop ? n : nop);
  }
  w.gu[0] = (float*)take(n);
  w.gu[1] = (float*)take(n);
  const bool tc = r8tc_supported(H);
  w.dc2 = tc ? (__nv_bfloat16*)take((size_t)B * r8tc_dcop_bytes(H)) : nullptr;
  for (int i = 0; i <= R8_LAYERS; ++i) w.uop[i] = tc ? (__nv_bfloat16*)take((size_t)B * r8tc_dcop_bytes(H)) : nullptr;
  for (int i = 0; i < 3; ++i) w.mask_bits[i] = tc ? (uint16_t*)take((size_t)B * 3 * r8tc_dcop_rows(H) * sizeof(uint16_t)) : nullptr;
  w.pooled_raw = (float*)take(sizeof(float) * B * R8_C);
  w.dones = (float*)take(sizeof(float) * R8_LAYERS * R8_C * 9);
  w.bn_coef = (float*)take(sizeof(float) * 3 * 48);
  w.bytes = off;
  return w;
}

// =============================================================================================
// conv0 (1 -> 45, 3x3, pad 1) + ReLU + AvgPool(3,4), one CTA per utterance
// =============================================================================================
#define C0_THREADS 288
#define C0_STRIDE (R8_MELS + 4)   // padded row of the staged feature tile (col 0 = left halo, 16-byte aligned groups)

__device__ __forceinline__ void c0_stage_tile(float* s_x, const float* __restrict__ feats, int F, int rows, int tid,
                                              int nthreads) {
  // s_x[(y + 1) * C0_STRIDE + (x + 1)], y in [-1, rows-2]; zero outside the clip
  for (int i = tid; i < rows * C0_STRIDE; i += nthreads) s_x[i] = 0.f;
  __syncthreads();
  const int nvalid = min(F, rows - 1);
  for (int i = tid; i < nvalid * (R8_MELS / 4); i += nthreads) {
    const int y = i / (R8_MELS / 4), x4 = i - y * (R8_MELS / 4);
    const float4 v = __ldg(reinterpret_cast<const float4*>(feats + (size_t)y * R8_MELS) + x4);
    float* d = s_x + (y + 1) * C0_STRIDE + x4 * 4 + 1;
    d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
  }
}

__global__ void __launch_bounds__(C0_THREADS) conv0_pool_kernel(const float* __restrict__ feats,
                                                                 const float* __restrict__ w0, float* __restrict__ a0,
                                                                 uint4* __restrict__ a0_op, uint16_t* __restrict__ bits0, int Rx, int F,
                                                                 int H) {
  extern __shared__ __align__(16) float smem[];
  const int rows = 3 * H + 2;
  float* s_x = smem;
  float* s_w = smem + rows * C0_STRIDE;
  const int tid = threadIdx.x;
  const int64_t b = blockIdx.x;
  uint4* op = a0_op ? a0_op + (size_t)b * 12 * Rx : nullptr;   // operand-format copy (res8_common.cuh), Rx rows
  if (op) {
    for (int r = tid; r < Rx; r += C0_THREADS) {                // rows outside the image are zero
      const int y = r / 11 - 1, x = r % 11 - 1;
      if (y >= 0 && y < H && x >= 0 && x < R8_W) continue;
#pragma unroll
      for (int g = 0; g < 12; ++g) op[g * Rx + r] = make_uint4(0, 0, 0, 0);
    }
  }
  for (int i = tid; i < R8_C * 9; i += C0_THREADS) s_w[i] = __ldg(w0 + i);
  c0_stage_tile(s_x, feats + b * (int64_t)F * R8_MELS, F, rows, tid, C0_THREADS);
  __syncthreads();
  const int HW = H * R8_W;
  for (int pp = tid; pp < HW; pp += C0_THREADS) {
    const int h = pp / R8_W, w = pp - h * R8_W;
    float patch[5][6];
#pragma unroll
    for (int r = 0; r < 5; ++r)
#pragma unroll
      for (int c = 0; c < 6; ++c) patch[r][c] = s_x[(3 * h + r) * C0_STRIDE + 4 * w + c];
    float* dst = a0 ? a0 + (b * R8_C) * (int64_t)HW + pp : nullptr;
    float ov[8];
#pragma unroll 1
    for (int oc = 0; oc < 48; ++oc) {
      if (oc >= R8_C) {
        ov[oc & 7] = (oc == R8_C) ? 1.f : 0.f;   // channel 45 = ones (weight-gradient BatchNorm fold), 46 / 47 padding
        if (oc == 47 && op) {
          uint4 hi, lo;
          tc::split8(ov, hi, lo);
          const int r = (h + 1) * 11 + (w + 1);
          op[5 * Rx + r] = hi;
          op[11 * Rx + r] = lo;
        }
        continue;
      }
      float wk[9];
#pragma unroll
      for (int k = 0; k < 9; ++k) wk[k] = s_w[oc * 9 + k];
      float sum = 0.f;
      uint32_t mb = 0;
#pragma unroll
      for (int py = 0; py < 3; ++py)
#pragma unroll
        for (int px = 0; px < 4; ++px) {
          float pre = 0.f;
#pragma unroll
          for (int ky = 0; ky < 3; ++ky)
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) pre = fmaf(wk[ky * 3 + kx], patch[py + ky][px + kx], pre);
          if (pre > 0.f) mb |= 1u << (py * 4 + px);
          sum += fmaxf(pre, 0.f);
        }
      if (bits0) bits0[(b * R8_C + oc) * (int64_t)HW + pp] = (uint16_t)mb;   // ReLU decisions of the pooling window, for conv0_bwd
      const float o = __fdiv_rn(sum, 12.f);
      if (dst) dst[(int64_t)oc * HW] = o;
      ov[oc & 7] = o;
      if ((oc & 7) == 7 && op) {
        uint4 hi, lo;
        tc::split8(ov, hi, lo);
        const int r = (h + 1) * 11 + (w + 1);
        op[(oc >> 3) * Rx + r] = hi;
        op[(6 + (oc >> 3)) * Rx + r] = lo;
      }
    }
  }
}

// backward of conv0: dW0[oc][k] = sum_{b,pp} G[oc,pp] / 12 * sum_{t in 3x4 window} relu'(pre[oc,pp,t]) * x[pos(pp,t) + k].  The ReLU
// decisions are the 12 bits per (oc, pooled pixel) that conv0_pool stored, so nothing is recomputed: per (oc, pp) the inner sum
// S[k] costs 108 multiply-adds with the 0/1 mask and the fold into the accumulators 9 more.  Persistent CTAs, one per SM.  The three
// per-utterance inputs (G planes, bits, features) are contiguous in HBM and arrive by TMA bulk copies into a two-deep ring, one
// utterance ahead of the arithmetic (the kernel used to spend most of its time in the staging loop's load latency); thread =
// (pixel group pg, output channel oc): the nine tap accumulators of its channel stay in registers across all pixels and utterances,
// the 5x6 input patch of a pooled pixel is a shared-memory broadcast.
#define C0B_GROUPS 12
#define C0B_THREADS (C0B_GROUPS * 48)
struct C0bLayout {
  uint32_t g_bytes, m_bytes, f_bytes;     // ring slot sizes (16-byte multiples, with slack for the alignment offset)
  uint32_t off_g, off_m, off_f, off_x, off_acc, total;
};
__host__ __device__ static inline C0bLayout c0b_layout(int F, int H, int slots) {
  C0bLayout L;
  const uint32_t HW = (uint32_t)H * R8_W;
  L.g_bytes = (R8_C * HW * 4 + 16 + 15) & ~15u;
  L.m_bytes = (R8_C * HW * 2 + 16 + 15) & ~15u;
  L.f_bytes = ((uint32_t)F * R8_MELS * 4 + 15) & ~15u;
  L.off_g = 0;
  L.off_m = L.off_g + slots * L.g_bytes;
  L.off_f = L.off_m + slots * L.m_bytes;
  L.off_x = L.off_f + slots * L.f_bytes;
  L.off_acc = L.off_x + (uint32_t)(3 * H + 2) * C0_STRIDE * 4;
  L.total = L.off_acc + R8_C * 9 * 4;
  return L;
}

__global__ void __launch_bounds__(C0B_THREADS, 1) conv0_bwd_kernel(const float* __restrict__ feats,
                                                                 const uint16_t* __restrict__ bits0,
                                                                 const float* __restrict__ g0, float* __restrict__ dw0,
                                                                 int64_t B, int F, int H, int slots) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ __align__(8) uint64_t bar_full[2];
  const C0bLayout L = c0b_layout(F, H, slots);
  const int rows = 3 * H + 2, HW = H * R8_W;
  float* s_x = reinterpret_cast<float*>(smem_raw + L.off_x);       // [(3H+2)][44]  feature tile with a zero halo
  float* s_acc = reinterpret_cast<float*>(smem_raw + L.off_acc);   // [45][9]
  const int tid = threadIdx.x;
  const int pg = tid / 48, oc = tid - pg * 48;
  const bool active = oc < R8_C;
  float acc[9];
#pragma unroll
  for (int k = 0; k < 9; ++k) acc[k] = 0.f;
  for (int i = tid; i < R8_C * 9; i += C0B_THREADS) s_acc[i] = 0.f;
  for (int i = tid; i < rows * C0_STRIDE; i += C0B_THREADS) s_x[i] = 0.f;      // halo stays zero for good
  if (tid == 0) {
    tc::mbar_init(&bar_full[0], 1);
    tc::mbar_init(&bar_full[1], 1);
    tc::fence_barrier_init();
  }
  __syncthreads();
  const int64_t n_local = (B - blockIdx.x + gridDim.x - 1) / gridDim.x;
  // 16-byte aligned supersets of the utterance's G planes / bits (their per-utterance sizes are not multiples of 16 bytes)
  auto issue = [&](int64_t k) {
    const int s = (int)(k % slots);
    const int64_t b = blockIdx.x + k * (int64_t)gridDim.x;
    const uintptr_t ga = reinterpret_cast<uintptr_t>(g0 + b * (int64_t)R8_C * HW), ga0 = ga & ~(uintptr_t)15;
    const uintptr_t ma = reinterpret_cast<uintptr_t>(bits0 + b * (int64_t)R8_C * HW), ma0 = ma & ~(uintptr_t)15;
    const uint32_t gb = (uint32_t)(((ga + (uintptr_t)R8_C * HW * 4 + 15) & ~(uintptr_t)15) - ga0);
    const uint32_t mb = (uint32_t)(((ma + (uintptr_t)R8_C * HW * 2 + 15) & ~(uintptr_t)15) - ma0);
    tc::mbar_expect_tx(&bar_full[s], gb + mb + L.f_bytes);
    tc::tma_bulk_g2s(smem_raw + L.off_g + s * L.g_bytes, reinterpret_cast<const void*>(ga0), gb, &bar_full[s]);
    tc::tma_bulk_g2s(smem_raw + L.off_m + s * L.m_bytes, reinterpret_cast<const void*>(ma0), mb, &bar_full[s]);
    tc::tma_bulk_g2s(smem_raw + L.off_f + s * L.f_bytes, feats + b * (int64_t)F * R8_MELS, L.f_bytes, &bar_full[s]);
  };
  if (tid == 0 && n_local > 0) issue(0);
  for (int64_t k = 0; k < n_local; ++k) {
    const int s = (int)(k % slots);
    const int64_t b = blockIdx.x + k * (int64_t)gridDim.x;
    // two slots: the other one was released by the barrier that ended iteration k - 1, so the next utterance loads during this one
    if (slots == 2 && tid == 0 && k + 1 < n_local) issue(k + 1);
    tc::mbar_wait(&bar_full[s], (uint32_t)((k / slots) & 1));
    const float* s_f = reinterpret_cast<const float*>(smem_raw + L.off_f + s * L.f_bytes);
    const float* s_g = reinterpret_cast<const float*>(smem_raw + L.off_g + s * L.g_bytes +
                                                      (reinterpret_cast<uintptr_t>(g0 + b * (int64_t)R8_C * HW) & 15));
    const uint16_t* s_m = reinterpret_cast<const uint16_t*>(smem_raw + L.off_m + s * L.m_bytes +
                                                            (reinterpret_cast<uintptr_t>(bits0 + b * (int64_t)R8_C * HW) & 15));
    // padded tile s_x[(y + 1) * 44 + (x + 1)] from the raw [F][40] features (rows beyond 3H + 1 are not needed)
    const int nvalid = min(F, rows - 1);
    for (int i = tid; i < nvalid * (R8_MELS / 4); i += C0B_THREADS) {
      const int y = i / (R8_MELS / 4), x4 = i - y * (R8_MELS / 4);
      const float4 v = reinterpret_cast<const float4*>(s_f + (size_t)y * R8_MELS)[x4];
      float* d = s_x + (y + 1) * C0_STRIDE + x4 * 4 + 1;
      d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
    }
    __syncthreads();
    if (active) {
      for (int pp = pg; pp < HW; pp += C0B_GROUPS) {
        const int h = pp / R8_W, w = pp - h * R8_W;
        const float gv = s_g[oc * HW + pp];
        const uint32_t mb = s_m[oc * HW + pp];
        float patch[5][6];
#pragma unroll
        for (int r = 0; r < 5; ++r) {
          const float* rowp = s_x + (3 * h + r) * C0_STRIDE + 4 * w;       // 16-byte aligned
          const float4 a = *reinterpret_cast<const float4*>(rowp);
          const float2 c2 = *reinterpret_cast<const float2*>(rowp + 4);
          patch[r][0] = a.x; patch[r][1] = a.y; patch[r][2] = a.z; patch[r][3] = a.w;
          patch[r][4] = c2.x; patch[r][5] = c2.y;
        }
        float S[9];
#pragma unroll
        for (int kk = 0; kk < 9; ++kk) S[kk] = 0.f;
#pragma unroll
        for (int py = 0; py < 3; ++py)
#pragma unroll
          for (int px = 0; px < 4; ++px) {
            const float m = ((mb >> (py * 4 + px)) & 1u) ? 1.f : 0.f;
#pragma unroll
            for (int ky = 0; ky < 3; ++ky)
#pragma unroll
              for (int kx = 0; kx < 3; ++kx) S[ky * 3 + kx] = fmaf(m, patch[py + ky][px + kx], S[ky * 3 + kx]);
          }
#pragma unroll
        for (int kk = 0; kk < 9; ++kk) acc[kk] = fmaf(gv, S[kk], acc[kk]);
      }
    }
    __syncthreads();       // slot s and the tile are free again
    if (slots == 1 && tid == 0 && k + 1 < n_local) issue(k + 1);      // long clips: one slot, no overlap
  }
  if (active) {
#pragma unroll
    for (int k = 0; k < 9; ++k) atomicAdd(&s_acc[oc * 9 + k], acc[k] * (1.f / 12.f));    // the average pooling's 1 / 12
  }
  __syncthreads();
  for (int i = tid; i < R8_C * 9; i += C0B_THREADS) atomicAdd(&dw0[i], s_acc[i]);
}

// =============================================================================================
// generic 45 -> 45 3x3 convolution (forward and data-gradient), persistent over utterances
// =============================================================================================
#define CV_THREADS 256
#define CV_OCG 9                  // output-channel groups of 5
#define CV_WSTRIDE 16             // floats per (c, dy, ocg) weight packet: [dx 3][o 5] + 1 pad


static size_t conv_smem_bytes(int H) {
  size_t f = (size_t)R8_C * 3 * CV_OCG * CV_WSTRIDE + (size_t)R8_C * (H + 2) * R8_WPAD + (size_t)H * CV_OCG * 10 +
             4 * 48;
  return f * sizeof(float);
}

template <bool RELU, int STATS>
__global__ void __launch_bounds__(CV_THREADS, 1) conv3x3_kernel(const ConvParams p) {
  extern __shared__ __align__(16) float smem[];
  const int H = p.H, HP = H + 2, HW = H * R8_W;
  float* s_w = smem;                                         // [45][3][9][16]
  float* s_in = s_w + R8_C * 3 * CV_OCG * CV_WSTRIDE;        // [45][H+2][12]
  float* s_part = s_in + R8_C * HP * R8_WPAD;                // [H*9][10]
  float* s_mean = s_part + H * CV_OCG * 10;                  // [48] x 4
  float* s_rstd = s_mean + 48;
  float* s_amean = s_rstd + 48;
  float* s_arstd = s_amean + 48;
  const int tid = threadIdx.x;

  // weights -> [c][dy][ocg][dx*5 + o]
  for (int i = tid; i < R8_KW; i += CV_THREADS) {
    const int oc = i / (R8_C * 9), rem = i - oc * (R8_C * 9);
    const int c = rem / 9, k = rem - c * 9, dy = k / 3, dx = k - dy * 3;
    s_w[((c * 3 + dy) * CV_OCG + oc / 5) * CV_WSTRIDE + dx * 5 + (oc % 5)] = __ldg(p.w + i);
  }
  for (int i = tid; i < R8_C * 3 * CV_OCG; i += CV_THREADS) s_w[i * CV_WSTRIDE + 15] = 0.f;
  for (int i = tid; i < R8_C * HP * R8_WPAD; i += CV_THREADS) s_in[i] = 0.f;   // halo stays zero for good
  if (tid < R8_C) {
    s_mean[tid] = p.in_mean ? p.in_mean[tid] : 0.f;
    s_rstd[tid] = p.in_rstd ? p.in_rstd[tid] : 1.f;
    if (STATS == 2) {
      s_amean[tid] = p.aux_mean[tid];
      s_arstd[tid] = p.aux_rstd[tid];
    }
  }
  const int items = H * CV_OCG;
  double stat_acc = 0.0;   // threads < 90: running per-channel statistic over this CTA's utterances

  for (int64_t b = blockIdx.x; b < p.B; b += gridDim.x) {
    __syncthreads();
    // ---- stage the (normalised) input tile
    const float* src = p.in + b * (int64_t)R8_C * HW;
    for (int e = tid; e < R8_C * H * 5; e += CV_THREADS) {
      const int c = e / (H * 5), rem = e - c * (H * 5), r = rem / 5, x2 = rem - r * 5;
      const float2 v = __ldg(reinterpret_cast<const float2*>(src) + e);
      const float mu = s_mean[c], rs = s_rstd[c];
      float* d = s_in + (c * HP + r + 1) * R8_WPAD + 1 + 2 * x2;
      d[0] = (v.x - mu) * rs;
      d[1] = (v.y - mu) * rs;
    }
    __syncthreads();
    for (int it0 = 0; it0 < items; it0 += CV_THREADS) {
      const int id = it0 + tid;
      const bool active = id < items;
      const int ocg = active ? id / H : 0, row = active ? id - (id / H) * H : 0;
      float acc[5][10];
#pragma unroll
      for (int o = 0; o < 5; ++o)
#pragma unroll
        for (int x = 0; x < 10; ++x) acc[o][x] = 0.f;
      const float* wp = s_w + ocg * CV_WSTRIDE;
      const float* ip = s_in + row * R8_WPAD;
      for (int c = 0; c < R8_C; ++c) {
#pragma unroll
        for (int dy = 0; dy < 3; ++dy) {
          float xin[12], wv[16];
          const float4* i4 = reinterpret_cast<const float4*>(ip + (c * HP + dy) * R8_WPAD);
          const float4* w4 = reinterpret_cast<const float4*>(wp + (c * 3 + dy) * CV_OCG * CV_WSTRIDE);
#pragma unroll
          for (int q = 0; q < 3; ++q) {
            const float4 t = i4[q];
            xin[4 * q] = t.x; xin[4 * q + 1] = t.y; xin[4 * q + 2] = t.z; xin[4 * q + 3] = t.w;
          }
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float4 t = w4[q];
            wv[4 * q] = t.x; wv[4 * q + 1] = t.y; wv[4 * q + 2] = t.z; wv[4 * q + 3] = t.w;
          }
#pragma unroll
          for (int dx = 0; dx < 3; ++dx)
#pragma unroll
            for (int o = 0; o < 5; ++o)
#pragma unroll
              for (int x = 0; x < 10; ++x) acc[o][x] = fmaf(wv[dx * 5 + o], xin[x + dx], acc[o][x]);
        }
      }
      // ---- epilogue
      if (active) {
#pragma unroll
        for (int o = 0; o < 5; ++o) {
          const int oc = ocg * 5 + o;
          const int64_t base = ((b * R8_C + oc) * (int64_t)H + row) * R8_W;
          float s1 = 0.f, s2 = 0.f;
          float am = 0.f, ar = 0.f;
          if (STATS == 2) {
            am = s_amean[oc];
            ar = s_arstd[oc];
          }
#pragma unroll
          for (int x2 = 0; x2 < 5; ++x2) {
            float v0 = acc[o][2 * x2], v1 = acc[o][2 * x2 + 1];
            if (RELU) {
              v0 = fmaxf(v0, 0.f);
              v1 = fmaxf(v1, 0.f);
            }
            if (p.res) {
              const float2 r = __ldg(reinterpret_cast<const float2*>(p.res + base) + x2);
              v0 += r.x;
              v1 += r.y;
            }
            reinterpret_cast<float2*>(p.out + base)[x2] = make_float2(v0, v1);
            if (STATS == 1) {
              s1 += v0 + v1;
              s2 = fmaf(v0, v0, fmaf(v1, v1, s2));
            } else if (STATS == 2) {
              const float2 a = __ldg(reinterpret_cast<const float2*>(p.aux + base) + x2);
              s1 += v0 + v1;
              s2 = fmaf(v0, (a.x - am) * ar, fmaf(v1, (a.y - am) * ar, s2));
            }
          }
          if (STATS) {
            s_part[id * 10 + o * 2] = s1;
            s_part[id * 10 + o * 2 + 1] = s2;
          }
        }
      }
    }
    if (STATS) {
      __syncthreads();
      if (tid < 2 * R8_C) {
        const int which = tid / R8_C, ch = tid - which * R8_C, ocg = ch / 5, o = ch - ocg * 5;
        float s = 0.f;
        for (int r = 0; r < H; ++r) s += s_part[(ocg * H + r) * 10 + o * 2 + which];
        stat_acc += (double)s;
      }
    }
  }
  if (STATS) {
    if (tid < 2 * R8_C) atomicAdd(&p.stats[tid], stat_acc);
  }
}

// =============================================================================================
// weight gradient of a 45 -> 45 3x3 convolution, persistent; thread tile 3 (out) x 3 (in) x 9 taps
// =============================================================================================
#define WG_THREADS 256


static size_t wgrad_smem_bytes(int H) {
  return sizeof(float) * ((size_t)R8_C * H * R8_WPAD + (size_t)R8_C * (H + 2) * R8_WPAD + 2 * 48);
}

__global__ void __launch_bounds__(WG_THREADS, 1) conv3x3_wgrad_kernel(const WgradParams p) {
  extern __shared__ __align__(16) float smem[];
  const int H = p.H, HP = H + 2, HW = H * R8_W;
  float* s_dc = smem;                              // [45][H][12]   (cols 0..9 used)
  float* s_x = s_dc + R8_C * H * R8_WPAD;          // [45][H+2][12] (halo zero)
  float* s_mean = s_x + R8_C * HP * R8_WPAD;
  float* s_rstd = s_mean + 48;
  const int tid = threadIdx.x;
  for (int i = tid; i < R8_C * HP * R8_WPAD; i += WG_THREADS) s_x[i] = 0.f;
  for (int i = tid; i < R8_C * H * R8_WPAD; i += WG_THREADS) s_dc[i] = 0.f;
  if (tid < R8_C) {
    s_mean[tid] = p.x_mean ? p.x_mean[tid] : 0.f;
    s_rstd[tid] = p.x_rstd ? p.x_rstd[tid] : 1.f;
  }
  const bool active = tid < 225;
  const int og = active ? tid / 15 : 0, cg = active ? tid - (tid / 15) * 15 : 0;
  float acc[3][3][9];
#pragma unroll
  for (int a = 0; a < 3; ++a)
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
      for (int k = 0; k < 9; ++k) acc[a][c][k] = 0.f;

  for (int64_t b = blockIdx.x; b < p.B; b += gridDim.x) {
    __syncthreads();
    const float* sdc = p.dc + b * (int64_t)R8_C * HW;
    const float* sx = p.x + b * (int64_t)R8_C * HW;
    for (int e = tid; e < R8_C * H * 5; e += WG_THREADS) {
      const int c = e / (H * 5), rem = e - c * (H * 5), r = rem / 5, x2 = rem - r * 5;
      const float2 d = __ldg(reinterpret_cast<const float2*>(sdc) + e);
      const float2 v = __ldg(reinterpret_cast<const float2*>(sx) + e);
      reinterpret_cast<float2*>(s_dc + (c * H + r) * R8_WPAD)[x2] = d;
      const float mu = s_mean[c], rs = s_rstd[c];
      float* q = s_x + (c * HP + r + 1) * R8_WPAD + 1 + 2 * x2;
      q[0] = (v.x - mu) * rs;
      q[1] = (v.y - mu) * rs;
    }
    __syncthreads();
    if (active) {
      for (int r = 0; r < H; ++r) {
        float dv[3][12];
#pragma unroll
        for (int a = 0; a < 3; ++a) {
          const float4* d4 = reinterpret_cast<const float4*>(s_dc + ((og * 3 + a) * H + r) * R8_WPAD);
#pragma unroll
          for (int q = 0; q < 3; ++q) {
            const float4 t = d4[q];
            dv[a][4 * q] = t.x; dv[a][4 * q + 1] = t.y; dv[a][4 * q + 2] = t.z; dv[a][4 * q + 3] = t.w;
          }
        }
#pragma unroll
        for (int dy = 0; dy < 3; ++dy) {
          float xv[3][12];
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            const float4* x4 = reinterpret_cast<const float4*>(s_x + ((cg * 3 + c) * HP + r + dy) * R8_WPAD);
#pragma unroll
            for (int q = 0; q < 3; ++q) {
              const float4 t = x4[q];
              xv[c][4 * q] = t.x; xv[c][4 * q + 1] = t.y; xv[c][4 * q + 2] = t.z; xv[c][4 * q + 3] = t.w;
            }
          }
#pragma unroll
          for (int a = 0; a < 3; ++a)
#pragma unroll
            for (int c = 0; c < 3; ++c)
#pragma unroll
              for (int dx = 0; dx < 3; ++dx)
#pragma unroll
                for (int x = 0; x < 10; ++x)
                  acc[a][c][dy * 3 + dx] = fmaf(dv[a][x], xv[c][x + dx], acc[a][c][dy * 3 + dx]);
        }
      }
    }
  }
  if (active) {
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int k = 0; k < 9; ++k) atomicAdd(&p.dw[((og * 3 + a) * R8_C + cg * 3 + c) * 9 + k], acc[a][c][k]);
  }
}

// =============================================================================================
// small kernels
// =============================================================================================
// BatchNorm statistics -> mean / rstd (+ running-stat update with momentum 0.1 and unbiased variance)
__global__ void bn_finalize_kernel(const double* __restrict__ stats, double count, float* __restrict__ mean_rstd,
                                   float* __restrict__ running, int64_t* __restrict__ nbt) {
  const int c = threadIdx.x;
  if (c < R8_C) {
    const double mean = stats[c] / count;
    double var = stats[R8_C + c] / count - mean * mean;
    if (var < 0.0) var = 0.0;
    mean_rstd[c] = (float)mean;
    mean_rstd[R8_C + c] = (float)(1.0 / sqrt(var + R8_BN_EPS));
    if (running) {
      const double unbiased = count > 1.0 ? var * count / (count - 1.0) : var;
      running[c] = (float)((1.0 - R8_BN_MOM) * running[c] + R8_BN_MOM * mean);
      running[R8_C + c] = (float)((1.0 - R8_BN_MOM) * running[R8_C + c] + R8_BN_MOM * unbiased);
    }
  }
  if (c == 0 && nbt) *nbt += 1;
}

// eval mode: running statistics -> mean / rstd for all six layers
__global__ void bn_eval_prepare_kernel(const float* __restrict__ running, float* __restrict__ mean_rstd) {
  const int i = blockIdx.x, c = threadIdx.x;
  if (c < R8_C) {
    mean_rstd[i * 2 * R8_C + c] = running[i * 2 * R8_C + c];
    mean_rstd[i * 2 * R8_C + R8_C + c] = 1.f / sqrtf(running[i * 2 * R8_C + R8_C + c] + (float)R8_BN_EPS);
  }
}

// bn6 -> spatial mean -> Linear(45 -> L); one CTA (128 threads) per utterance
__global__ void __launch_bounds__(128) head_fwd_kernel(const float* __restrict__ u6, const float* __restrict__ pooled_raw,
                                                       const float* __restrict__ mean_rstd,
                                                       const float* __restrict__ wout, const float* __restrict__ bout,
                                                       float* __restrict__ pooled, float* __restrict__ logits,
                                                       float* __restrict__ logits_ws, int HW, int L) {
  __shared__ float s_pool[48];
  const int64_t b = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int c = warp; c < R8_C; c += 4) {
    float s = 0.f;
    if (pooled_raw) {
      s = pooled_raw[b * R8_C + c];
    } else {
      const float* src = u6 + (b * R8_C + c) * (int64_t)HW;
      for (int i = lane; i < HW; i += 32) s += src[i];
      s = warp_sum(s);
    }
    if (lane == 0) {
      const float v = (s / (float)HW - mean_rstd[c]) * mean_rstd[R8_C + c];
      s_pool[c] = v;
      pooled[b * R8_C + c] = v;
    }
  }
  __syncthreads();
  for (int l = threadIdx.x; l < L; l += blockDim.x) {
    float acc = bout[l];
    for (int c = 0; c < R8_C; ++c) acc = fmaf(wout[l * R8_C + c], s_pool[c], acc);
    logits[b * L + l] = acc;
    logits_ws[b * L + l] = acc;
  }
}

// softmax cross-entropy backward at the head: one thread per utterance
__global__ void head_bwd_kernel(const float* __restrict__ logits, const int64_t* __restrict__ labels,
                                const float* __restrict__ dlogits_in, const float* __restrict__ wout,
                                float* __restrict__ dlogits, float* __restrict__ dh, double* __restrict__ loss_acc,
                                int64_t B, int L, float inv_batch) {
  const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  double nll = 0.0;
  if (b < B && !labels) {
    // upstream gradient supplied by the caller (autograd path): only dh = dlogits . W_out is needed
    float acc[R8_C];
#pragma unroll
    for (int c = 0; c < R8_C; ++c) acc[c] = 0.f;
    for (int l = 0; l < L; ++l) {
      const float d = dlogits_in[b * L + l];
      dlogits[b * L + l] = d;
#pragma unroll
      for (int c = 0; c < R8_C; ++c) acc[c] = fmaf(d, __ldg(wout + l * R8_C + c), acc[c]);
    }
#pragma unroll
    for (int c = 0; c < R8_C; ++c) dh[b * R8_C + c] = acc[c];
  } else if (b < B) {
    const float* z = logits + b * L;
    float mx = z[0];
    for (int l = 1; l < L; ++l) mx = fmaxf(mx, z[l]);
    float se = 0.f;
    for (int l = 0; l < L; ++l) se += expf(z[l] - mx);
    const float lse = mx + logf(se);
    const int64_t y = labels[b];
    float acc[R8_C];
#pragma unroll
    for (int c = 0; c < R8_C; ++c) acc[c] = 0.f;
    for (int l = 0; l < L; ++l) {
      const float pl = expf(z[l] - lse);
      const float d = (pl - (l == y ? 1.f : 0.f)) * inv_batch;
      dlogits[b * L + l] = d;
#pragma unroll
      for (int c = 0; c < R8_C; ++c) acc[c] = fmaf(d, __ldg(wout + l * R8_C + c), acc[c]);
    }
#pragma unroll
    for (int c = 0; c < R8_C; ++c) dh[b * R8_C + c] = acc[c];
    if (y >= 0 && y < L) nll = (double)(lse - z[y]);
  }
  nll = warp_sum(nll);
  if ((threadIdx.x & 31) == 0 && nll != 0.0) atomicAdd(loss_acc, nll * (double)inv_batch);
}

// head parameter gradients + the two BatchNorm-backward statistics of layer 6 (g6 = dh / HW broadcast over pixels)
//   blocks 0..L-1 : d output.weight[l][:] ; block L : d output.bias ; block L+1 : stats_bwd[5]
__global__ void __launch_bounds__(256) head_wgrad_kernel(const float* __restrict__ dlogits,
                                                          const float* __restrict__ pooled,
                                                          const float* __restrict__ dh, float* __restrict__ dwout,
                                                          float* __restrict__ dbout, double* __restrict__ stats6,
                                                          const double* __restrict__ loss_acc, float* __restrict__ loss,
                                                          int64_t B, int L) {
  // grid = (L + 2 roles, batch slices): every CTA reduces its slice of the batch and accumulates into the (zeroed) outputs
  __shared__ double sh[8][128];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int blk = blockIdx.x;
  const int64_t per = (B + gridDim.y - 1) / gridDim.y;
  const int64_t lo = (int64_t)blockIdx.y * per, hi = lo + per < B ? lo + per : B;
  double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
  if (blk < L) {                       // d output.weight[blk][lane], [lane + 32]
    for (int64_t b = lo + warp; b < hi; b += 8) {
      const double d = dlogits[b * L + blk];
      a0 += d * pooled[b * R8_C + lane];
      if (lane + 32 < R8_C) a1 += d * pooled[b * R8_C + lane + 32];
    }
  } else if (blk == L) {               // d output.bias
    for (int64_t b = lo + warp; b < hi; b += 8) {
      if (lane < L) a0 += dlogits[b * L + lane];
      if (lane + 32 < L) a1 += dlogits[b * L + lane + 32];
      if (lane + 64 < L) a2 += dlogits[b * L + lane + 64];
    }
  } else {                             // BatchNorm-6 backward statistics: sum dh, sum dh * pooled per channel
    for (int64_t b = lo + warp; b < hi; b += 8) {
      const float d0 = dh[b * R8_C + lane];
      a0 += d0;
      a1 += (double)d0 * pooled[b * R8_C + lane];
      if (lane + 32 < R8_C) {
        const float d1 = dh[b * R8_C + lane + 32];
        a2 += d1;
        a3 += (double)d1 * pooled[b * R8_C + lane + 32];
      }
    }
  }
  sh[warp][lane] = a0;
  sh[warp][32 + lane] = a1;
  sh[warp][64 + lane] = a2;
  sh[warp][96 + lane] = a3;
  __syncthreads();
  if (warp != 0) return;
  double t0 = 0.0, t1 = 0.0, t2 = 0.0, t3 = 0.0;
  for (int w = 0; w < 8; ++w) {
    t0 += sh[w][lane];
    t1 += sh[w][32 + lane];
    t2 += sh[w][64 + lane];
    t3 += sh[w][96 + lane];
  }
  if (blk < L) {
    atomicAdd(dwout + blk * R8_C + lane, (float)t0);
    if (lane + 32 < R8_C) atomicAdd(dwout + blk * R8_C + lane + 32, (float)t1);
  } else if (blk == L) {
    if (lane < L) atomicAdd(dbout + lane, (float)t0);
    if (lane + 32 < L) atomicAdd(dbout + lane + 32, (float)t1);
    if (lane + 64 < L) atomicAdd(dbout + lane + 64, (float)t2);
  } else {
    atomicAdd(stats6 + lane, t0);
    atomicAdd(stats6 + R8_C + lane, t1);
    if (lane + 32 < R8_C) {
      atomicAdd(stats6 + 32 + lane, t2);
      atomicAdd(stats6 + R8_C + 32 + lane, t3);
    }
    if (lane == 0 && blockIdx.y == 0) *loss = (float)(*loss_acc);
  }
}

// BatchNorm backward + residual fan-in + ReLU mask:
//   G  = rstd * (g - m1 - xhat * m2) [+ gu_in];   gu_out = G (even layers);   dc = mask ? G : 0
//   mask = (u > mask_prev) for residual layers (relu output y = u - prev), (u > 0) otherwise
struct ApplyParams {
  const float* g;          // full tensor, or null when g_bcast is used
  const float* g_bcast;    // [B,45]: g = g_bcast / HW (layer 6)
  const float* u;
  const float* mean_rstd;  // [2][45] of this layer
  const double* stats;     // [2][45] sum(g), sum(g xhat)
  const float* gu_in;
  const float* mask_prev;
  float* gu_out;
  float* dc;
  int64_t n2;              // number of float2 elements
  int HW;
  double count;
};

__global__ void __launch_bounds__(256) bn_bwd_apply_kernel(const ApplyParams p) {
  const int hw2 = p.HW / 2;
  const float inv_hw = 1.f / (float)p.HW;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < p.n2; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t plane = i / hw2;            // b * 45 + c
    const int c = (int)(plane % R8_C);
    const float mu = __ldg(p.mean_rstd + c), rs = __ldg(p.mean_rstd + R8_C + c);
    const float m1 = (float)(p.stats[c] / p.count), m2 = (float)(p.stats[R8_C + c] / p.count);
    float2 g;
    if (p.g) {
      g = reinterpret_cast<const float2*>(p.g)[i];
    } else {
      const float v = __ldg(p.g_bcast + plane) * inv_hw;
      g = make_float2(v, v);
    }
    const float2 u = reinterpret_cast<const float2*>(p.u)[i];
    float2 G;
    G.x = rs * (g.x - m1 - (u.x - mu) * rs * m2);
    G.y = rs * (g.y - m1 - (u.y - mu) * rs * m2);
    if (p.gu_in) {
      const float2 t = reinterpret_cast<const float2*>(p.gu_in)[i];
      G.x += t.x;
      G.y += t.y;
    }
    if (p.gu_out) reinterpret_cast<float2*>(p.gu_out)[i] = G;
    float2 prev = make_float2(0.f, 0.f);
    if (p.mask_prev) prev = reinterpret_cast<const float2*>(p.mask_prev)[i];
    float2 d;
    d.x = (u.x > prev.x) ? G.x : 0.f;
    d.y = (u.y > prev.y) ? G.y : 0.f;
    reinterpret_cast<float2*>(p.dc)[i] = d;
  }
}

// Wt[l][c][o][2-dy][2-dx] = W[l][o][c][dy][dx]
__global__ void transpose_weights_kernel(const float* __restrict__ w, float* __restrict__ wt) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= R8_LAYERS * R8_KW) return;
  const int l = i / R8_KW, rem = i - l * R8_KW;
  const int o = rem / (R8_C * 9), r2 = rem - o * (R8_C * 9), c = r2 / 9, k = r2 - c * 9;
  wt[l * R8_KW + (c * R8_C + o) * 9 + (8 - k)] = w[i];
}

// =============================================================================================
// host side
// =============================================================================================
extern "C" int64_t howl_b200_res8_param_count(int32_t num_labels) {
  if (num_labels < 1) return -1;
  return (int64_t)R8_C * 9 + (int64_t)R8_LAYERS * R8_KW + (int64_t)num_labels * R8_C + num_labels;
}

extern "C" int64_t howl_b200_res8_workspace_bytes(int64_t B, int32_t frames, int32_t n_mels, int32_t num_labels,
                                                  int train) {
  (void)train;
  if (B < 0 || frames < 3 || n_mels != R8_MELS || num_labels < 1) return -1;
  return (int64_t)r8_carve(nullptr, B, frames / 3, num_labels).bytes;
}

static int r8_check(howl_ctx_t* ctx, int64_t B, int frames, int n_mels, int L, const void* ws, size_t ws_bytes,
                    R8Ws* out) {
  HOWL_REQUIRE(ctx, n_mels == R8_MELS, HOWL_E_UNSUPPORTED, "res8: n_mels=%d (kernels are built for 40)", n_mels);
  HOWL_REQUIRE(ctx, frames >= 3, HOWL_E_INVALID, "res8: %d frames is fewer than one pooling window", frames);
  HOWL_REQUIRE(ctx, L >= 1 && L <= 96, HOWL_E_UNSUPPORTED, "res8: num_labels=%d outside 1..96", L);
  HOWL_REQUIRE(ctx, B >= 1, HOWL_E_INVALID, "res8: empty batch");
  const int H = frames / 3;
  HOWL_REQUIRE(ctx, conv_smem_bytes(H) <= 227 * 1024 && wgrad_smem_bytes(H) <= 227 * 1024, HOWL_E_UNSUPPORTED,
               "res8: %d frames exceeds the shared-memory tile of the conv kernels", frames);
  HOWL_REQUIRE(ctx, ws != nullptr, HOWL_E_WORKSPACE, "res8: null workspace");
  *out = r8_carve(const_cast<void*>(ws), B, H, L);
  HOWL_REQUIRE(ctx, out->bytes <= ws_bytes, HOWL_E_WORKSPACE, "res8: workspace %zu < required %zu", ws_bytes,
               out->bytes);
  return HOWL_OK;
}

static int r8_grid(howl_ctx_t* ctx, int64_t B) { return (int)(B < ctx->sm_count ? B : ctx->sm_count); }

extern "C" int howl_b200_res8_fwd(howl_ctx_t* ctx, void* stream, const float* feats, int64_t B, int32_t frames,
                                  int32_t n_mels, int32_t num_labels, const float* params, float* bn_running,
                                  int64_t* num_batches_tracked, int train, float* logits, void* workspace,
                                  size_t workspace_bytes) {
  if (!ctx) return HOWL_E_INVALID;
  HOWL_REQUIRE(ctx, feats && params && bn_running && logits, HOWL_E_INVALID, "res8_fwd: null pointer");
  R8Ws ws;
  int rc = r8_check(ctx, B, frames, n_mels, num_labels, workspace, workspace_bytes, &ws);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const int H = frames / 3, HW = H * R8_W, L = num_labels;
  const float* w0 = params;
  const float* wl = params + R8_C * 9;
  const float* wout = wl + (size_t)R8_LAYERS * R8_KW;
  const float* bout = wout + (size_t)L * R8_C;

  const bool use_tc = ctx->conv_engine >= 1 && r8tc_supported(H);
  if (use_tc) {
    // tensor-core engine: conv0 + ReLU + pool as split-bf16 GEMMs over im2col tiles built in shared memory; a0 in operand format only
    rc = r8tc_conv0(ctx, st, feats, w0, ws.uop[0], train ? ws.bits0 : nullptr, B, frames, H);
    if (rc) return rc;
  } else {
    const size_t sm = sizeof(float) * ((3 * H + 2) * C0_STRIDE + R8_C * 9);
    HOWL_CUDA(ctx, cudaFuncSetAttribute(conv0_pool_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    conv0_pool_kernel<<<(unsigned)B, C0_THREADS, sm, st>>>(feats, w0, ws.a0, nullptr, train ? ws.bits0 : nullptr, r8tc_dcop_rows(H), frames, H);
    HOWL_LAUNCHED(ctx, "conv0_pool");
  }
  if (train) {
    HOWL_CUDA(ctx, cudaMemsetAsync(ws.stats_fwd, 0, sizeof(double) * R8_LAYERS * 2 * R8_C, st));
  } else {
    bn_eval_prepare_kernel<<<R8_LAYERS, 64, 0, st>>>(bn_running, ws.mean_rstd);
    HOWL_LAUNCHED(ctx, "bn_eval_prepare");
  }
  if (use_tc) HOWL_CUDA(ctx, cudaMemsetAsync(ws.pooled_raw, 0, sizeof(float) * B * R8_C, st));
  const size_t csm = conv_smem_bytes(H);
  HOWL_CUDA(ctx, cudaFuncSetAttribute(conv3x3_kernel<true, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)csm));
  HOWL_CUDA(ctx, cudaFuncSetAttribute(conv3x3_kernel<true, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)csm));
  const int grid = r8_grid(ctx, B);
  const double count = (double)B * HW;
  for (int i = 1; i <= R8_LAYERS; ++i) {
    const float* in_mean = (i > 1) ? ws.mean_rstd + (i - 2) * 2 * R8_C : nullptr;
    const float* w_i = wl + (size_t)(i - 1) * R8_KW;
    double* stats = train ? ws.stats_fwd + (i - 1) * 2 * R8_C : nullptr;
    if (use_tc) {
      // operand-format activations only; BatchNorm of layer i-1 folded into this layer's weights (res8_tc.cu)
      __nv_bfloat16* wblk = ws.wprep + ((size_t)((i - 1) * 2 + 0) * 2) * R8TC_WBLOCK;
      // train: the BatchNorm finalisation of layer i - 1 (mean / rstd, running statistics) happens inside this fold
      TcFoldBn fb;
      memset(&fb, 0, sizeof(fb));
      if (train && i > 1) {
        fb.stats = ws.stats_fwd + (i - 2) * 2 * R8_C;
        fb.count = count;
        fb.mean_rstd_out = ws.mean_rstd + (i - 2) * 2 * R8_C;
        fb.running = bn_running + (i - 2) * 2 * R8_C;
        fb.nbt = num_batches_tracked ? num_batches_tracked + (i - 2) : nullptr;
      }
      rc = r8tc_fold(ctx, st, w_i, in_mean, wblk, fb.stats ? &fb : nullptr);
      if (rc) return rc;
      TcConvCall c;
      memset(&c, 0, sizeof(c));
      c.mode = train ? 1 : 0;
      c.B = B; c.H = H;
      c.in_op = ws.uop[i - 1];
      c.w = wblk;
      c.out_op = ws.uop[i];
      c.res_op = (i % 2 == 0) ? ws.uop[i - 2] : nullptr;
      c.mask_out = (train && i % 2 == 0) ? ws.mask_bits[i / 2 - 1] : nullptr;
      c.pooled_raw = (i == R8_LAYERS) ? ws.pooled_raw : nullptr;
      c.stats = stats;
      rc = r8tc_conv(ctx, st, c);
      if (rc) return rc;
    } else {
      ConvParams p;
      memset(&p, 0, sizeof(p));
      p.in = (i == 1) ? ws.a0 : ws.u[i - 2];
      p.in_mean = in_mean;
      p.in_rstd = in_mean ? in_mean + R8_C : nullptr;
      p.w = w_i;
      p.res = (i % 2 == 0) ? ((i == 2) ? ws.a0 : ws.u[i - 3]) : nullptr;
      p.out = ws.u[i - 1];
      p.B = B;
      p.H = H;
      p.stats = stats;
      if (train) conv3x3_kernel<true, 1><<<grid, CV_THREADS, csm, st>>>(p);
      else conv3x3_kernel<true, 0><<<grid, CV_THREADS, csm, st>>>(p);
      HOWL_LAUNCHED(ctx, "conv3x3_fwd");
    }
    if (train && !(use_tc && i < R8_LAYERS)) {      // (tensor-core engine: layers 1..5 are finalised by the next layer's fold)
      bn_finalize_kernel<<<1, 64, 0, st>>>(stats, count, ws.mean_rstd + (i - 1) * 2 * R8_C, bn_running + (i - 1) * 2 * R8_C,
                                           num_batches_tracked ? num_batches_tracked + (i - 1) : nullptr);
      HOWL_LAUNCHED(ctx, "bn_finalize");
    }
  }
  head_fwd_kernel<<<(unsigned)B, 128, 0, st>>>(ws.u[5], use_tc ? ws.pooled_raw : nullptr, ws.mean_rstd + 5 * 2 * R8_C, wout, bout,
                                               ws.pooled, logits, ws.logits, HW, L);
  HOWL_LAUNCHED(ctx, "head_fwd");
  return HOWL_OK;
}

static int r8_bwd_impl(howl_ctx_t* ctx, void* stream, const float* feats, const int64_t* labels,
                       const float* dlogits_in, int64_t B, int32_t frames, int32_t n_mels, int32_t num_labels,
                       int64_t loss_scale_batch, const float* params, float* grads, float* loss, void* workspace,
                       size_t workspace_bytes) {
  R8Ws ws;
  int rc = r8_check(ctx, B, frames, n_mels, num_labels, workspace, workspace_bytes, &ws);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const int H = frames / 3, HW = H * R8_W, L = num_labels;
  const float* wl = params + R8_C * 9;
  const float* wout = wl + (size_t)R8_LAYERS * R8_KW;
  float* g_w0 = grads;
  float* g_wl = grads + R8_C * 9;
  float* g_wout = g_wl + (size_t)R8_LAYERS * R8_KW;
  float* g_bout = g_wout + (size_t)L * R8_C;
  const int64_t nparam = howl_b200_res8_param_count(L);
  HOWL_CUDA(ctx, cudaMemsetAsync(grads, 0, sizeof(float) * nparam, st));
  HOWL_CUDA(ctx, cudaMemsetAsync(ws.stats_bwd, 0, sizeof(double) * R8_LAYERS * 2 * R8_C, st));
  HOWL_CUDA(ctx, cudaMemsetAsync(ws.loss_acc, 0, sizeof(double) * 2, st));

  const bool use_tc = ctx->conv_engine >= 1 && r8tc_supported(frames / 3);
  if (use_tc) {
    rc = r8tc_weight_prep(ctx, st, wl, ws.wprep);
    if (rc) return rc;
  } else {
    transpose_weights_kernel<<<(R8_LAYERS * R8_KW + 255) / 256, 256, 0, st>>>(wl, ws.wT);
    HOWL_LAUNCHED(ctx, "transpose_weights");
  }
  head_bwd_kernel<<<(unsigned)howl_ceil_div(B, 128), 128, 0, st>>>(ws.logits, labels, dlogits_in, wout, ws.dlogits,
                                                                   ws.dh, ws.loss_acc, B, L,
                                                                   1.f / (float)loss_scale_batch);
  HOWL_LAUNCHED(ctx, "head_bwd");
  head_wgrad_kernel<<<dim3(L + 2, (unsigned)std::max<int64_t>(1, std::min<int64_t>(16, B / 128))), 256, 0, st>>>(ws.dlogits, ws.pooled, ws.dh, g_wout, g_bout,
                                           ws.stats_bwd + 5 * 2 * R8_C, ws.loss_acc, loss ? loss : (float*)ws.loss_acc + 2,
                                           B, L);
  HOWL_LAUNCHED(ctx, "head_wgrad");

  const double count = (double)B * HW;
  if (use_tc) {
    // ---- tensor-core engine.  Per layer i = 6..1:  weight gradient (also yields the BatchNorm-backward statistics of layer i-1,
    // bn_bwd_coef_kernel)  ->  data gradient whose epilogue applies BatchNorm backward + residual fan-in + ReLU mask of layer i-1 and
    // writes the next conv-output gradient directly in both operand formats.  No planar gradient tensor, no separate apply pass.
    HOWL_CUDA(ctx, cudaMemsetAsync(ws.dones, 0, sizeof(float) * R8_LAYERS * R8_C * 9, st));
    __nv_bfloat16* dc[2] = {reinterpret_cast<__nv_bfloat16*>(ws.dc), ws.dc2};
    int cur = 0;
    {
      ApplyOpParams ao;
      memset(&ao, 0, sizeof(ao));
      ao.g_bcast = ws.dh;
      ao.u_op = ws.uop[R8_LAYERS];
      ao.mask_bits = ws.mask_bits[2];
      ao.mean_rstd = ws.mean_rstd + 5 * 2 * R8_C;
      ao.stats = ws.stats_bwd + 5 * 2 * R8_C;
      ao.gu_out = ws.gu[(R8_LAYERS / 2) & 1];
      ao.dc_op = dc[cur];
      ao.B = B; ao.H = H; ao.count = count;
      rc = r8tc_apply_head(ctx, st, ao);
      if (rc) return rc;
    }
    for (int i = R8_LAYERS; i >= 1; --i) {
      const int j = i - 1;                                  // the layer whose output feeds conv_i
      const float* x_mean = (i > 1) ? ws.mean_rstd + (j - 1) * 2 * R8_C : nullptr;
      float* dw = g_wl + (size_t)(i - 1) * R8_KW;
      float* dones = ws.dones + (size_t)(i - 1) * R8_C * 9;
      rc = r8tc_wgrad(ctx, st, dc[cur], ws.uop[j], x_mean, x_mean ? x_mean + R8_C : nullptr, dw, dones, B, H);
      if (rc) return rc;
      TcConvCall c;
      memset(&c, 0, sizeof(c));
      c.B = B; c.H = H;
      c.in_op = dc[cur];
      c.w = ws.wprep + ((size_t)((i - 1) * 2 + 1) * 2) * R8TC_WBLOCK;
      if (i > 1) {
        rc = r8tc_bn_bwd_coef(ctx, st, wl + (size_t)(i - 1) * R8_KW, dw, dones, x_mean, count, ws.bn_coef);
        if (rc) return rc;
        c.mode = 3;
        c.u_op = ws.uop[j];
        c.bn_coef = ws.bn_coef;
        if (j % 2 == 0) {
          c.gu_in = ws.gu[((j + 2) / 2) & 1];
          c.gu_out = ws.gu[(j / 2) & 1];
          c.mask_in = ws.mask_bits[j / 2 - 1];
        }
        c.dc_out = dc[cur ^ 1];
      } else {
        c.mode = 2;
        c.out_planar = ws.g;                                // dL/d(a0) = conv1 path + residual path (G_2, added in the epilogue)
        c.gu_in = ws.gu[1];
      }
      rc = r8tc_conv(ctx, st, c);
      if (rc) return rc;
      cur ^= 1;
    }
  } else {
  const size_t csm = conv_smem_bytes(H), wsm = wgrad_smem_bytes(H);
  HOWL_CUDA(ctx, cudaFuncSetAttribute(conv3x3_kernel<false, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)csm));
  HOWL_CUDA(ctx, cudaFuncSetAttribute(conv3x3_kernel<false, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)csm));
  HOWL_CUDA(ctx, cudaFuncSetAttribute(conv3x3_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wsm));
  const int grid = r8_grid(ctx, B);
  const int64_t n2 = (int64_t)B * R8_C * HW / 2;
  int64_t ablocks = howl_ceil_div(n2, 256);
  if (ablocks > (int64_t)ctx->sm_count * 16) ablocks = (int64_t)ctx->sm_count * 16;

  for (int i = R8_LAYERS; i >= 1; --i) {
    const bool even = (i % 2 == 0);
    ApplyParams a;
    memset(&a, 0, sizeof(a));
    if (i == R8_LAYERS) a.g_bcast = ws.dh; else a.g = ws.g;
    a.u = ws.u[i - 1];
    a.mean_rstd = ws.mean_rstd + (i - 1) * 2 * R8_C;
    a.stats = ws.stats_bwd + (i - 1) * 2 * R8_C;
    if (even) {
      a.gu_in = (i < R8_LAYERS) ? ws.gu[((i + 2) / 2) & 1] : nullptr;
      a.gu_out = ws.gu[(i / 2) & 1];
      a.mask_prev = (i == 2) ? ws.a0 : ws.u[i - 3];
    }
    a.dc = ws.dc;
    a.n2 = n2;
    a.HW = HW;
    a.count = count;
    bn_bwd_apply_kernel<<<(unsigned)ablocks, 256, 0, st>>>(a);
    HOWL_LAUNCHED(ctx, "bn_bwd_apply");

    WgradParams wg;
    memset(&wg, 0, sizeof(wg));
    wg.dc = ws.dc;
    wg.x = (i == 1) ? ws.a0 : ws.u[i - 2];
    if (i > 1) {
      wg.x_mean = ws.mean_rstd + (i - 2) * 2 * R8_C;
      wg.x_rstd = wg.x_mean + R8_C;
    }
    wg.dw = g_wl + (size_t)(i - 1) * R8_KW;
    wg.B = B;
    wg.H = H;
    conv3x3_wgrad_kernel<<<grid, WG_THREADS, wsm, st>>>(wg);
    HOWL_LAUNCHED(ctx, "conv3x3_wgrad");

    ConvParams p;
    memset(&p, 0, sizeof(p));
    p.in = ws.dc;
    p.w = ws.wT + (size_t)(i - 1) * R8_KW;
    p.out = ws.g;
    p.B = B;
    p.H = H;
    if (i > 1) {
      p.stats = ws.stats_bwd + (i - 2) * 2 * R8_C;
      p.aux = ws.u[i - 2];
      p.aux_mean = ws.mean_rstd + (i - 2) * 2 * R8_C;
      p.aux_rstd = p.aux_mean + R8_C;
    }
    if (i == 1) p.res = ws.gu[1];      // dL/d(a0) = conv1 path + residual path (G_2), summed here for conv0_bwd
    if (i > 1) conv3x3_kernel<false, 2><<<grid, CV_THREADS, csm, st>>>(p);
    else conv3x3_kernel<false, 0><<<grid, CV_THREADS, csm, st>>>(p);
    HOWL_LAUNCHED(ctx, "conv3x3_dgrad");
  }
  }
  {
    // G_0 = dL/d(a0): ws.g already holds the sum of the conv1 path and the residual path (the layer-1 data gradient adds gu)
    const int slots = c0b_layout(frames, H, 2).total <= 227 * 1024 ? 2 : 1;
    const size_t sm = c0b_layout(frames, H, slots).total;
    HOWL_REQUIRE(ctx, sm <= 227 * 1024, HOWL_E_UNSUPPORTED, "res8: %d frames exceeds the shared-memory ring of conv0_bwd", frames);
    HOWL_REQUIRE(ctx, (reinterpret_cast<uintptr_t>(feats) & 15) == 0, HOWL_E_INVALID, "res8_bwd: feats must be 16-byte aligned");
    HOWL_CUDA(ctx, cudaFuncSetAttribute(conv0_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    const int g0 = (int)(B < ctx->sm_count ? B : ctx->sm_count);
    conv0_bwd_kernel<<<g0, C0B_THREADS, sm, st>>>(feats, ws.bits0, ws.g, g_w0, B, frames, H, slots);
    HOWL_LAUNCHED(ctx, "conv0_bwd");
  }
  return HOWL_OK;
}


extern "C" int howl_b200_res8_bwd(howl_ctx_t* ctx, void* stream, const float* feats, const int64_t* labels, int64_t B,
                                  int32_t frames, int32_t n_mels, int32_t num_labels, int64_t loss_scale_batch,
                                  const float* params, float* grads, float* loss, void* workspace,
                                  size_t workspace_bytes) {
  if (!ctx) return HOWL_E_INVALID;
  HOWL_REQUIRE(ctx, feats && labels && params && grads && loss, HOWL_E_INVALID, "res8_bwd: null pointer");
  HOWL_REQUIRE(ctx, loss_scale_batch >= 1, HOWL_E_INVALID, "res8_bwd: loss_scale_batch must be >= 1");
  return r8_bwd_impl(ctx, stream, feats, labels, nullptr, B, frames, n_mels, num_labels, loss_scale_batch, params, grads,
                     loss, workspace, workspace_bytes);
}

extern "C" int howl_b200_res8_bwd_dlogits(howl_ctx_t* ctx, void* stream, const float* feats, const float* dlogits,
                                          int64_t B, int32_t frames, int32_t n_mels, int32_t num_labels,
                                          const float* params, float* grads, void* workspace, size_t workspace_bytes) {
  if (!ctx) return HOWL_E_INVALID;
  HOWL_REQUIRE(ctx, feats && dlogits && params && grads, HOWL_E_INVALID, "res8_bwd_dlogits: null pointer");
  return r8_bwd_impl(ctx, stream, feats, nullptr, dlogits, B, frames, n_mels, num_labels, 1, params, grads, nullptr,
                     workspace, workspace_bytes);
}

// =============================================================================================
// test hook (include/howl_b200_debug.h): the ReLU decisions of the backward, for the mask-forced gradient oracle
// =============================================================================================
// conv0: the bits conv0_pool stored (bit py * 4 + px of pooled pixel (h, w) <-> pre-pool pixel (3h + py, 4w + px))
__global__ void debug_mask0_kernel(const uint16_t* __restrict__ bits0, uint8_t* __restrict__ mask0, int64_t B, int H) {
  const int rows = 3 * H, HW = H * R8_W;
  const int64_t n = B * R8_C * rows * R8_MELS;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int x = (int)(i % R8_MELS), y = (int)((i / R8_MELS) % rows);
    const int64_t plane = i / ((int64_t)R8_MELS * rows);      // b * 45 + oc
    const uint32_t mb = bits0[plane * HW + (y / 3) * R8_W + x / 4];
    mask0[i] = (mb >> ((y % 3) * 4 + (x % 4))) & 1u;
  }
}

__global__ void debug_mask_kernel(const float* __restrict__ u, const float* __restrict__ prev, uint8_t* __restrict__ mask, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    mask[i] = u[i] > (prev ? prev[i] : 0.f) ? 1 : 0;
}

extern "C" int howl_b200_res8_debug_masks(howl_ctx_t* ctx, void* stream, const float* feats, const float* params, int64_t B,
                                          int32_t frames, int32_t n_mels, int32_t num_labels, const void* workspace,
                                          size_t workspace_bytes, uint8_t* mask0, uint8_t* masks16) {
  if (!ctx) return HOWL_E_INVALID;
  HOWL_REQUIRE(ctx, feats && params, HOWL_E_INVALID, "res8_debug_masks: null pointer");
  R8Ws ws;
  int rc = r8_check(ctx, B, frames, n_mels, num_labels, workspace, workspace_bytes, &ws);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const int H = frames / 3;
  const int64_t n = B * R8_C * H * R8_W;
  if (mask0) {
    debug_mask0_kernel<<<ctx->sm_count * 8, 256, 0, st>>>(ws.bits0, mask0, B, H);
    HOWL_LAUNCHED(ctx, "debug_mask0");
  }
  if (masks16) {
    const bool use_tc = ctx->conv_engine >= 1 && r8tc_supported(H);
    for (int i = 1; i <= R8_LAYERS; ++i) {
      if (use_tc) {   // odd layers: u > 0 from the operand-format activations; residual layers: the bits the forward stored
        rc = r8tc_debug_mask(ctx, st, ws.uop[i], (i % 2 == 0) ? ws.mask_bits[i / 2 - 1] : nullptr, masks16 + (size_t)(i - 1) * n, B, H);
        if (rc) return rc;
        continue;
      }
      const float* prev = (i % 2 == 0) ? ((i == 2) ? ws.a0 : ws.u[i - 3]) : nullptr;
      debug_mask_kernel<<<ctx->sm_count * 8, 256, 0, st>>>(ws.u[i - 1], prev, masks16 + (size_t)(i - 1) * n, n);
      HOWL_LAUNCHED(ctx, "debug_mask");
    }
  }
  return HOWL_OK;
}
