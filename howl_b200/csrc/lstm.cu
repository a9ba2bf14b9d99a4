// K5 -- lstm / seq-lstm (howl/model/rnn.py:41-91): nn.LSTM(n_mels -> 128) over the first `length` frames of the
// time-major features, then Linear(128 -> 256) + ReLU + Linear(256 -> L); frame objective (CrossEntropy on dnn(h_n)) and
// its full backward (BPTT), exact fp32.
//
// Recurrence: batch-parallel.  One CTA owns 16 sequences for all time steps (512 threads):
//   gates[b][j] = sum_k Wt[k][j] * xh[b][k],  xh = [x_t (n_mels) | h_{t-1} (128)]
// with xh in shared memory ([k][sequence] tile, broadcast 128-bit reads) and the transposed weights streamed from L2 (344 KB per
// step and CTA, coalesced along j).  Three engines with the same summation order per output (option "lstm_engine"): 0 = the plain
// kernels (thread = gate row x 16 sequences); 1 = software pipelined (weights of the next 8 k in a register double buffer, x_{t+1} and
// the backward's saved tensors fetched a step ahead); 2 (default) = pipelined with a 2 rows x 8 sequences register tile.  The cell update is done by the same 512 threads on (sequence, unit) pairs.  For
// training the activated gates, cell states and the xh rows are kept in the workspace; the backward walks the steps in
// reverse (dh through W_hh with a 4-way split of the 512-long reduction) and leaves the pre-activation gradients in
// place of the gates, so that the weight gradients are tall-skinny A^T B products over all (t, b) rows: on the tensor cores
// (operands split into (hi, lo) bf16, three products through one TMEM accumulator, mbn_gemm.cu) for training-size batches, the
// exact-fp32 lstm_atb_kernel for a few hundred rows.  The per-frame MLP head of seq-lstm (X W^T over all T * B rows) runs on
// the tensor cores the same way (bf16 x 3 folded into K, fp32 epilogue); the recurrence itself is FP32 FFMA.
#include <math.h>

#include "common.cuh"
#include "mbn_common.cuh"

#define LS_H 128
#define LS_G 512            // 4 * hidden, rows ordered (i, f, g, o) as in torch.nn.LSTM
#define LS_MLP 256
#define LS_NB 16            // sequences per CTA
#define LS_THREADS 512
#define CTC_UMAX 32         // longest target sequence supported by the CTC kernel
#define CTC_SMAX (2 * CTC_UMAX + 1 + 3)

struct LstmWs {
  float* wt;        // [M + 128][512]  transposed [W_ih | W_hh]
  float* bsum;      // [512]  b_ih + b_hh
  float* w1t;       // [128][256]  dnn.0.weight transposed
  float* h0;        // [B][128]  initial state (zeros or streaming state)
  float* c0;
  float* gates;     // [T][B][512]  activated gates, overwritten by pre-activation gradients in the backward
  float* cs;        // [T][B][128]  cell states
  float* xh;        // [T][B][M + 128]  inputs of every step
  float* hfin;      // [B][128]
  float* cfin;
  float* hseq;      // [T][B][128]  (sequential mode)
  float* z1;        // [R][256]
  float* logits;    // [B][L]
  float* dlogits;
  float* dz1;       // [R][256]
  float* dh;        // [R][128]  gradient at h_n (R = B) or at every h_t (R = T*B, sequential)
  float* alpha;     // [B][T][CTC_SMAX]  CTC forward variables (sequential + train)
  double* loss_acc;
  // (hi, lo) bf16 operands of the weight-gradient products on the tensor cores (mbn_atb3_packed)
  __nv_bfloat16 *px_hi, *px_lo;   // X side: up to [T * B][512]
  __nv_bfloat16 *py_hi, *py_lo;   // Y side: up to [T * B][256]
  // head products X W^T on the tensor cores (mbn_gemm_nt3_f32): X3 = [hi | hi | lo] of h or dz1, W3 operands of dnn.0.weight
  __nv_bfloat16* p3;              // [rows][3 * 256]
  __nv_bfloat16* wop_w1;          // dnn.0.weight [256][128]   -> z1 = h W1^T
  __nv_bfloat16* wop_w1t;         // its transpose [128][256]  -> dh = dz1 W1
  size_t bytes;
};

static LstmWs lstm_carve(void* base, int64_t B, int T, int M, int L, int train, int sequential) {
  LstmWs w;
  memset(&w, 0, sizeof(w));
  size_t off = 0;
  char* p = (char*)base;
  auto take = [&](size_t bytes) {
    void* r = p ? (void*)(p + off) : nullptr;
    off += howl_align_up(bytes, 256);
    return r;
  };
  const int K = M + LS_H;
  w.wt = (float*)take(sizeof(float) * K * LS_G);
  w.bsum = (float*)take(sizeof(float) * LS_G);
  w.w1t = (float*)take(sizeof(float) * LS_H * LS_MLP);
  w.h0 = (float*)take(sizeof(float) * B * LS_H);
  w.c0 = (float*)take(sizeof(float) * B * LS_H);
  w.hfin = (float*)take(sizeof(float) * B * LS_H);
  w.cfin = (float*)take(sizeof(float) * B * LS_H);
  w.loss_acc = (double*)take(sizeof(double) * 2);
  const int64_t rows = sequential ? (int64_t)T * B : B;
  w.z1 = (float*)take(sizeof(float) * rows * LS_MLP);
  w.logits = (float*)take(sizeof(float) * rows * L);
  w.dlogits = (float*)take(sizeof(float) * rows * L);
  w.dz1 = (float*)take(sizeof(float) * rows * LS_MLP);
  w.dh = (float*)take(sizeof(float) * rows * LS_H);
  if (sequential) w.hseq = (float*)take(sizeof(float) * (size_t)T * B * LS_H);
  if (sequential && train) w.alpha = (float*)take(sizeof(float) * (size_t)B * T * CTC_SMAX);
  if (train) {
    w.gates = (float*)take(sizeof(float) * (size_t)T * B * LS_G);
    w.cs = (float*)take(sizeof(float) * (size_t)T * B * LS_H);
    w.xh = (float*)take(sizeof(float) * (size_t)T * B * K);
    const int64_t R = (int64_t)T * B;
    w.px_hi = (__nv_bfloat16*)take(mbn_tmo_bytes(R, LS_G));
    w.px_lo = (__nv_bfloat16*)take(mbn_tmo_bytes(R, LS_G));
    w.py_hi = (__nv_bfloat16*)take(mbn_tmo_bytes(R, LS_MLP));
    w.py_lo = (__nv_bfloat16*)take(mbn_tmo_bytes(R, LS_MLP));
    w.p3 = (__nv_bfloat16*)take(mbn_tmo_bytes(rows, 3 * LS_MLP));
    w.wop_w1 = (__nv_bfloat16*)take(mbn_weight_operand3_bytes(LS_MLP, LS_H));
    w.wop_w1t = (__nv_bfloat16*)take(mbn_weight_operand3_bytes(LS_H, LS_MLP));
  }
  w.bytes = off;
  return w;
}

// ---- parameter views in the flat layout (state_dict order) ------------------------
struct LstmParams {
  const float *w_ih, *w_hh, *b_ih, *b_hh, *w1, *b1, *w2, *b2;
};
static LstmParams lstm_views(const float* p, int M, int L) {
  LstmParams v;
  v.w_ih = p;
  v.w_hh = v.w_ih + (size_t)LS_G * M;
  v.b_ih = v.w_hh + (size_t)LS_G * LS_H;
  v.b_hh = v.b_ih + LS_G;
  v.w1 = v.b_hh + LS_G;
  v.b1 = v.w1 + (size_t)LS_MLP * LS_H;
  v.w2 = v.b1 + LS_MLP;
  v.b2 = v.w2 + (size_t)L * LS_MLP;
  return v;
}

extern "C" int64_t howl_b200_lstm_param_count(int32_t num_labels, int32_t n_mels) {
  if (num_labels < 1 || n_mels < 1) return -1;
  return (int64_t)LS_G * n_mels + (int64_t)LS_G * LS_H + 2 * LS_G + (int64_t)LS_MLP * LS_H + LS_MLP +
         (int64_t)num_labels * LS_MLP + num_labels;
}

extern "C" int64_t howl_b200_lstm_workspace_bytes(int64_t B, int32_t max_steps, int32_t n_mels, int32_t num_labels,
                                                  int train, int sequential) {
  if (B < 1 || max_steps < 1 || n_mels < 1 || n_mels > HOWL_MAX_MELS || num_labels < 1) return -1;
  return (int64_t)lstm_carve(nullptr, B, max_steps, n_mels, num_labels, train, sequential).bytes;
}

// =============================================================================================
// small helpers
// =============================================================================================
__global__ void lstm_prep_kernel(LstmParams v, int M, float* __restrict__ wt, float* __restrict__ bsum,
                                 float* __restrict__ w1t) {
  const int K = M + LS_H;
  const int n = K * LS_G;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int k = i / LS_G, j = i - k * LS_G;
    wt[i] = k < M ? v.w_ih[(size_t)j * M + k] : v.w_hh[(size_t)j * LS_H + (k - M)];
  }
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < LS_G; i += gridDim.x * blockDim.x) bsum[i] = v.b_ih[i] + v.b_hh[i];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < LS_H * LS_MLP; i += gridDim.x * blockDim.x) {
    const int k = i / LS_MLP, nn = i - k * LS_MLP;
    w1t[i] = v.w1[(size_t)nn * LS_H + k];
  }
}

__device__ __forceinline__ float sigmoidf_acc(float x) { return 1.f / (1.f + expf(-x)); }

// =============================================================================================
// forward recurrence
// =============================================================================================
struct LstmFwdArgs {
  const float* feats;     // [B][F][M]
  const int64_t* lengths; // [B]
  const float* wt;        // [K][512]
  const float* bsum;
  const float* h0;
  const float* c0;
  float* gates;           // or null
  float* cs;
  float* xh;
  float* hfin;
  float* cfin;
  float* hseq;            // or null
  int64_t B;
  int F, M, T;
};

__global__ void __launch_bounds__(LS_THREADS, 1) lstm_fwd_kernel(const LstmFwdArgs a) {
  extern __shared__ __align__(16) float smem[];
  const int M = a.M, K = M + LS_H;
  float* s_xh = smem;                        // [K][16]
  float* s_g = s_xh + K * LS_NB;             // [16][512]
  float* s_c = s_g + LS_NB * LS_G;           // [16][128]
  __shared__ int s_len[LS_NB];
  const int tid = threadIdx.x;
  const int64_t b0 = (int64_t)blockIdx.x * LS_NB;
  if (tid < LS_NB) s_len[tid] = (b0 + tid < a.B) ? (int)a.lengths[b0 + tid] : 0;
  for (int p = tid; p < LS_NB * LS_H; p += LS_THREADS) {
    const int b = p / LS_H, u = p - b * LS_H;
    const bool ok = b0 + b < a.B;
    s_xh[(M + u) * LS_NB + b] = ok ? a.h0[(b0 + b) * LS_H + u] : 0.f;
    s_c[b * LS_H + u] = ok ? a.c0[(b0 + b) * LS_H + u] : 0.f;
  }
  const float bias = a.bsum[tid];
  for (int t = 0; t < a.T; ++t) {
    // x_t of the 16 sequences
    for (int p = tid; p < LS_NB * M; p += LS_THREADS) {
      const int b = p / M, k = p - b * M;
      s_xh[k * LS_NB + b] = (b0 + b < a.B && t < a.F) ? __ldg(a.feats + ((b0 + b) * (int64_t)a.F + t) * M + k) : 0.f;
    }
    __syncthreads();
    // gate row `tid` for the 16 sequences
    float acc[LS_NB];
#pragma unroll
    for (int b = 0; b < LS_NB; ++b) acc[b] = bias;
    const float* wp = a.wt + tid;
#pragma unroll 4
    for (int k = 0; k < K; ++k) {
      const float w = __ldg(wp + (size_t)k * LS_G);
      const float4* x4 = reinterpret_cast<const float4*>(s_xh + k * LS_NB);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float4 x = x4[q];
        acc[4 * q] = fmaf(w, x.x, acc[4 * q]);
        acc[4 * q + 1] = fmaf(w, x.y, acc[4 * q + 1]);
        acc[4 * q + 2] = fmaf(w, x.z, acc[4 * q + 2]);
        acc[4 * q + 3] = fmaf(w, x.w, acc[4 * q + 3]);
      }
    }
#pragma unroll
    for (int b = 0; b < LS_NB; ++b) s_g[b * LS_G + tid] = acc[b];
    // keep the inputs of this step for the weight gradients
    if (a.xh) {
      for (int p = tid; p < LS_NB * K; p += LS_THREADS) {
        const int b = p / K, k = p - b * K;
        if (b0 + b < a.B) a.xh[((size_t)t * a.B + b0 + b) * K + k] = s_xh[k * LS_NB + b];
      }
    }
    __syncthreads();
    // cell update on (sequence, unit) pairs
    for (int p = tid; p < LS_NB * LS_H; p += LS_THREADS) {
      const int b = p / LS_H, u = p - b * LS_H;
      const float gi = sigmoidf_acc(s_g[b * LS_G + u]);
      const float gf = sigmoidf_acc(s_g[b * LS_G + LS_H + u]);
      const float gg = tanhf(s_g[b * LS_G + 2 * LS_H + u]);
      const float go = sigmoidf_acc(s_g[b * LS_G + 3 * LS_H + u]);
      const bool live = t < s_len[b];
      const float c_old = s_c[b * LS_H + u];
      const float c_new = gf * c_old + gi * gg;
      const float h_new = go * tanhf(c_new);
      if (live) {
        s_c[b * LS_H + u] = c_new;
        s_xh[(M + u) * LS_NB + b] = h_new;
      }
      if (b0 + b < a.B) {
        const size_t row = (size_t)t * a.B + b0 + b;
        if (a.gates) {
          float* g = a.gates + row * LS_G;
          g[u] = gi; g[LS_H + u] = gf; g[2 * LS_H + u] = gg; g[3 * LS_H + u] = go;
          a.cs[row * LS_H + u] = live ? c_new : c_old;
        }
        if (a.hseq) a.hseq[row * LS_H + u] = live ? h_new : 0.f;
      }
    }
    __syncthreads();
  }
  for (int p = tid; p < LS_NB * LS_H; p += LS_THREADS) {
    const int b = p / LS_H, u = p - b * LS_H;
    if (b0 + b < a.B) {
      a.hfin[(b0 + b) * LS_H + u] = s_xh[(M + u) * LS_NB + b];
      a.cfin[(b0 + b) * LS_H + u] = s_c[b * LS_H + u];
    }
  }
}

// ---------------------------------------------------------------------------------------------
// forward recurrence, software pipelined (option "lstm_engine" = 1, and = 2 with the register tile below).  Same mapping and the SAME summation order as
// lstm_fwd_kernel above (bit-identical results); what changes is when the data moves:
//   * the weights of the next 8 k are loaded from L2 while the current 8 are multiplied (register double buffer; the first
//     batch never changes and stays resident), so the ~300-600 cycle L2 round trip no longer sits in front of every 4 k;
//   * x_{t+1} is fetched from HBM during the product of step t and dropped into shared memory during the cell update
//     (one __syncthreads less per step, no DRAM latency on the critical path);
//   * the [k][sequence] tiles are padded to 20 floats per k (16-way -> 4-way bank conflicts of the transposing stores).
// ---------------------------------------------------------------------------------------------
#define LS_NBP 20           // padded sequence stride of the [k][16] shared-memory tiles (multiple of 4: 128-bit broadcast reads)
#define LS_WB 8             // weights in flight per thread

__device__ __forceinline__ void lstm_fma16(float (&acc)[LS_NB], const float w, const float* __restrict__ xrow) {
  const float4* x4 = reinterpret_cast<const float4*>(xrow);
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const float4 x = x4[q];
    acc[4 * q] = fmaf(w, x.x, acc[4 * q]);
    acc[4 * q + 1] = fmaf(w, x.y, acc[4 * q + 1]);
    acc[4 * q + 2] = fmaf(w, x.z, acc[4 * q + 2]);
    acc[4 * q + 3] = fmaf(w, x.w, acc[4 * q + 3]);
  }
}

// acc[b] += sum_{k < n} w[k * ldw] * xs[k * LS_NBP + b], k ascending.  wc[] holds the first LS_WB weights on entry AND on exit (the
// prefetch of the last batch wraps around to batch 0, i.e. it is the next step's first batch): the double buffer lives across steps.
__device__ __forceinline__ void lstm_dot_pipelined(float (&acc)[LS_NB], const float* __restrict__ w, const size_t ldw, const float* __restrict__ xs,
                                                   const int n, float (&wc)[LS_WB]) {
  const int nb = n / LS_WB;
#pragma unroll 2
  for (int kb = 0; kb < nb; ++kb) {
    float wn[LS_WB];
    const int kn = (kb + 1 < nb) ? (kb + 1) * LS_WB : 0;
#pragma unroll
    for (int i = 0; i < LS_WB; ++i) wn[i] = __ldg(w + (size_t)(kn + i) * ldw);
#pragma unroll
    for (int i = 0; i < LS_WB; ++i) lstm_fma16(acc, wc[i], xs + (kb * LS_WB + i) * LS_NBP);
#pragma unroll
    for (int i = 0; i < LS_WB; ++i) wc[i] = wn[i];
  }
  for (int k = nb * LS_WB; k < n; ++k) lstm_fma16(acc, __ldg(w + (size_t)k * ldw), xs + k * LS_NBP);
}

// Register tile of 2 rows x 8 sequences (option "lstm_engine" = 2): the same 16 FFMA per k and thread, but they consume two 128-bit
// shared-memory reads instead of four -- the product of the 1 x 16 tile is bound by the shared-memory pipe (every lane of a warp
// receives the 16 sequences of a k), not by FFMA issue.  Lanes 0-15 / 16-31 of a warp hold the two sequence halves of the same 16
// row pairs, so the weight loads of the two half-warps coalesce into one request.  Per output the k order is unchanged (bit-identical).
__device__ __forceinline__ void lstm_fma2x8(float (&acc)[2][8], const float wa, const float wb, const float* __restrict__ xrow) {
  const float4* x4 = reinterpret_cast<const float4*>(xrow);
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    const float4 x = x4[q];
    acc[0][4 * q] = fmaf(wa, x.x, acc[0][4 * q]);
    acc[0][4 * q + 1] = fmaf(wa, x.y, acc[0][4 * q + 1]);
    acc[0][4 * q + 2] = fmaf(wa, x.z, acc[0][4 * q + 2]);
    acc[0][4 * q + 3] = fmaf(wa, x.w, acc[0][4 * q + 3]);
    acc[1][4 * q] = fmaf(wb, x.x, acc[1][4 * q]);
    acc[1][4 * q + 1] = fmaf(wb, x.y, acc[1][4 * q + 1]);
    acc[1][4 * q + 2] = fmaf(wb, x.z, acc[1][4 * q + 2]);
    acc[1][4 * q + 3] = fmaf(wb, x.w, acc[1][4 * q + 3]);
  }
}

// acc[r][j] += sum_{k < n} w[k * ldw + r * row_b] * xs[k * LS_NBP + j]; wc as in lstm_dot_pipelined
__device__ __forceinline__ void lstm_dot2x8_pipelined(float (&acc)[2][8], const float* __restrict__ w, const size_t ldw, const int row_b,
                                                      const float* __restrict__ xs, const int n, float (&wc)[2][LS_WB]) {
  const int nb = n / LS_WB;
#pragma unroll 2
  for (int kb = 0; kb < nb; ++kb) {
    float wn[2][LS_WB];
    const int kn = (kb + 1 < nb) ? (kb + 1) * LS_WB : 0;
#pragma unroll
    for (int i = 0; i < LS_WB; ++i) {
      wn[0][i] = __ldg(w + (size_t)(kn + i) * ldw);
      wn[1][i] = __ldg(w + (size_t)(kn + i) * ldw + row_b);
    }
#pragma unroll
    for (int i = 0; i < LS_WB; ++i) lstm_fma2x8(acc, wc[0][i], wc[1][i], xs + (kb * LS_WB + i) * LS_NBP);
#pragma unroll
    for (int i = 0; i < LS_WB; ++i) {
      wc[0][i] = wn[0][i];
      wc[1][i] = wn[1][i];
    }
  }
  for (int k = nb * LS_WB; k < n; ++k) lstm_fma2x8(acc, __ldg(w + (size_t)k * ldw), __ldg(w + (size_t)k * ldw + row_b), xs + k * LS_NBP);
}

template <int RT>   // RT = 1: thread = gate row x 16 sequences; RT = 2: thread = 2 gate rows x 8 sequences
__global__ void __launch_bounds__(LS_THREADS, 1) lstm_fwd_pipe_kernel(const LstmFwdArgs a) {
  extern __shared__ __align__(16) float smem[];
  const int M = a.M, K = M + LS_H;
  float* s_xh = smem;                        // [K][20]
  float* s_g = s_xh + K * LS_NBP;            // [16][512]
  float* s_c = s_g + LS_NB * LS_G;           // [16][128]
  __shared__ int s_len[LS_NB];
  const int tid = threadIdx.x;
  const int64_t b0 = (int64_t)blockIdx.x * LS_NB;
  if (tid < LS_NB) s_len[tid] = (b0 + tid < a.B) ? (int)a.lengths[b0 + tid] : 0;
  for (int p = tid; p < LS_NB * LS_H; p += LS_THREADS) {
    const int b = p / LS_H, u = p - b * LS_H;
    const bool ok = b0 + b < a.B;
    s_xh[(M + u) * LS_NBP + b] = ok ? a.h0[(b0 + b) * LS_H + u] : 0.f;
    s_c[b * LS_H + u] = ok ? a.c0[(b0 + b) * LS_H + u] : 0.f;
  }
  // x of one step: element p = (sequence p / M, mel p % M), up to four per thread (M <= 128)
  float xr[4];
  auto load_x = [&](int t) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int p = tid + i * LS_THREADS;
      const int b = p / M, k = p - b * M;
      xr[i] = (p < LS_NB * M && b0 + b < a.B && t < a.F && t < a.T) ? __ldg(a.feats + ((b0 + b) * (int64_t)a.F + t) * M + k) : 0.f;
    }
  };
  auto store_x = [&]() {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int p = tid + i * LS_THREADS;
      const int b = p / M, k = p - b * M;
      if (p < LS_NB * M) s_xh[k * LS_NBP + b] = xr[i];
    }
  };
  load_x(0);
  store_x();
  load_x(1);
  // RT = 1: gate row tid, sequences 0..15.  RT = 2: gate rows rA and rA + 256 (warp w: rows 16 w .. 16 w + 15), sequences sq0 .. sq0 + 7
  const int rA = RT == 1 ? tid : (tid >> 5) * 16 + (tid & 15);
  const int sq0 = RT == 1 ? 0 : 8 * ((tid >> 4) & 1);
  const float bias = a.bsum[rA], bias_b = a.bsum[RT == 1 ? rA : rA + LS_G / 2];
  const float* wp = a.wt + rA;
  float w0[RT][LS_WB];
#pragma unroll
  for (int i = 0; i < LS_WB; ++i) {                                          // K >= 132 > LS_WB
    w0[0][i] = __ldg(wp + (size_t)i * LS_G);
    if (RT == 2) w0[RT - 1][i] = __ldg(wp + (size_t)i * LS_G + LS_G / 2);
  }
  __syncthreads();
  for (int t = 0; t < a.T; ++t) {
    if constexpr (RT == 1) {
      float acc[LS_NB];
#pragma unroll
      for (int b = 0; b < LS_NB; ++b) acc[b] = bias;
      lstm_dot_pipelined(acc, wp, LS_G, s_xh, K, w0[0]);
#pragma unroll
      for (int b = 0; b < LS_NB; ++b) s_g[b * LS_G + tid] = acc[b];
    } else {
      float acc[2][8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        acc[0][j] = bias;
        acc[1][j] = bias_b;
      }
      lstm_dot2x8_pipelined(acc, wp, LS_G, LS_G / 2, s_xh + sq0, K, w0);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        s_g[(sq0 + j) * LS_G + rA] = acc[0][j];
        s_g[(sq0 + j) * LS_G + rA + LS_G / 2] = acc[1][j];
      }
    }
    // keep the inputs of this step for the weight gradients
    if (a.xh) {
      for (int p = tid; p < LS_NB * K; p += LS_THREADS) {
        const int b = p / K, k = p - b * K;
        if (b0 + b < a.B) a.xh[((size_t)t * a.B + b0 + b) * K + k] = s_xh[k * LS_NBP + b];
      }
    }
    __syncthreads();
    // cell update on (sequence, unit) pairs
    for (int p = tid; p < LS_NB * LS_H; p += LS_THREADS) {
      const int b = p / LS_H, u = p - b * LS_H;
      const float gi = sigmoidf_acc(s_g[b * LS_G + u]);
      const float gf = sigmoidf_acc(s_g[b * LS_G + LS_H + u]);
      const float gg = tanhf(s_g[b * LS_G + 2 * LS_H + u]);
      const float go = sigmoidf_acc(s_g[b * LS_G + 3 * LS_H + u]);
      const bool live = t < s_len[b];
      const float c_old = s_c[b * LS_H + u];
      const float c_new = gf * c_old + gi * gg;
      const float h_new = go * tanhf(c_new);
      if (live) {
        s_c[b * LS_H + u] = c_new;
        s_xh[(M + u) * LS_NBP + b] = h_new;
      }
      if (b0 + b < a.B) {
        const size_t row = (size_t)t * a.B + b0 + b;
        if (a.gates) {
          float* g = a.gates + row * LS_G;
          g[u] = gi; g[LS_H + u] = gf; g[2 * LS_H + u] = gg; g[3 * LS_H + u] = go;
          a.cs[row * LS_H + u] = live ? c_new : c_old;
        }
        if (a.hseq) a.hseq[row * LS_H + u] = live ? h_new : 0.f;
      }
    }
    // x_{t+1} (in registers since the previous step) replaces x_t, whose last readers finished before the barrier above
    store_x();
    load_x(t + 2);
    __syncthreads();
  }
  for (int p = tid; p < LS_NB * LS_H; p += LS_THREADS) {
    const int b = p / LS_H, u = p - b * LS_H;
    if (b0 + b < a.B) {
      a.hfin[(b0 + b) * LS_H + u] = s_xh[(M + u) * LS_NBP + b];
      a.cfin[(b0 + b) * LS_H + u] = s_c[b * LS_H + u];
    }
  }
}

// =============================================================================================
// head: Linear(128 -> 256) + ReLU + Linear(256 -> L) on `rows` rows of h; 16 rows per CTA, 256 threads
// =============================================================================================
__global__ void __launch_bounds__(256) lstm_head_fwd_kernel(const float* __restrict__ h, int64_t rows,
                                                            const float* __restrict__ w1t, const float* __restrict__ b1,
                                                            const float* __restrict__ w2, const float* __restrict__ b2,
                                                            float* __restrict__ z1_save, float* __restrict__ out,
                                                            float* __restrict__ out2, int L) {
  __shared__ __align__(16) float s_h[LS_H * 16];      // [k][16]
  __shared__ float s_z[16 * LS_MLP];                  // [r][256]
  const int tid = threadIdx.x;
  const int64_t r0 = (int64_t)blockIdx.x * 16;
  for (int p = tid; p < 16 * LS_H; p += 256) {
    const int r = p / LS_H, k = p - r * LS_H;
    s_h[k * 16 + r] = (r0 + r < rows) ? h[(r0 + r) * LS_H + k] : 0.f;
  }
  __syncthreads();
  float acc[16];
  const float bb = b1[tid];
#pragma unroll
  for (int r = 0; r < 16; ++r) acc[r] = bb;
#pragma unroll 4
  for (int k = 0; k < LS_H; ++k) {
    const float w = __ldg(w1t + k * LS_MLP + tid);
    const float4* x4 = reinterpret_cast<const float4*>(s_h + k * 16);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float4 x = x4[q];
      acc[4 * q] = fmaf(w, x.x, acc[4 * q]);
      acc[4 * q + 1] = fmaf(w, x.y, acc[4 * q + 1]);
      acc[4 * q + 2] = fmaf(w, x.z, acc[4 * q + 2]);
      acc[4 * q + 3] = fmaf(w, x.w, acc[4 * q + 3]);
    }
  }
#pragma unroll
  for (int r = 0; r < 16; ++r) {
    const float z = fmaxf(acc[r], 0.f);
    s_z[r * LS_MLP + tid] = z;
    if (z1_save && r0 + r < rows) z1_save[(r0 + r) * LS_MLP + tid] = z;
  }
  __syncthreads();
  for (int p = tid; p < 16 * L; p += 256) {
    const int r = p / L, l = p - r * L;
    if (r0 + r >= rows) continue;
    float s = b2[l];
    for (int n = 0; n < LS_MLP; ++n) s = fmaf(__ldg(w2 + l * LS_MLP + n), s_z[r * LS_MLP + n], s);
    out[(r0 + r) * L + l] = s;
    if (out2) out2[(r0 + r) * L + l] = s;
  }
}

// ---- large row counts (training on every frame): the two 128 <-> 256 products run on the tensor cores (mbn_gemm_nt3_f32) and these
// two kernels do the L-wide rest.  One warp per row.
#define LH_TC_ROWS 4096          // below this the 16-row FFMA kernels are used
__global__ void __launch_bounds__(256) lstm_scores_kernel(const float* __restrict__ z1, int64_t rows, const float* __restrict__ w2,
                                                          const float* __restrict__ b2, float* __restrict__ out, float* __restrict__ out2, int L) {
  const int lane = threadIdx.x & 31;
  const int64_t r = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (r >= rows) return;
  float z[LS_MLP / 32];
#pragma unroll
  for (int i = 0; i < LS_MLP / 32; ++i) z[i] = z1[r * LS_MLP + i * 32 + lane];
  for (int l = 0; l < L; ++l) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < LS_MLP / 32; ++i) s = fmaf(__ldg(w2 + l * LS_MLP + i * 32 + lane), z[i], s);
    s = warp_sum(s);
    if (lane == 0) {
      s += b2[l];
      out[r * L + l] = s;
      if (out2) out2[r * L + l] = s;
    }
  }
}

// CE (or given dlogits) -> dlogits, loss;  dz1[r][n] = relu'(z1) * sum_l dlogits[r][l] W2[l][n]
__global__ void __launch_bounds__(256) lstm_dz1_kernel(const float* __restrict__ logits, const int64_t* __restrict__ labels,
                                                       const float* __restrict__ dlogits_in, const float* __restrict__ z1,
                                                       const float* __restrict__ w2, float* __restrict__ dlogits, float* __restrict__ dz1,
                                                       double* __restrict__ loss_acc, int64_t rows, int L, float inv_batch) {
  __shared__ float s_dl[8][96];
  const int lane = threadIdx.x & 31, wv = threadIdx.x >> 5;
  const int64_t r = (int64_t)blockIdx.x * 8 + wv;
  if (r < rows && lane == 0) {
    if (labels) {
      const float* z = logits + r * L;
      float mx = z[0];
      for (int l = 1; l < L; ++l) mx = fmaxf(mx, z[l]);
      float se = 0.f;
      for (int l = 0; l < L; ++l) se += expf(z[l] - mx);
      const float lse = mx + logf(se);
      const int64_t y = labels[r];
      for (int l = 0; l < L; ++l) s_dl[wv][l] = (expf(z[l] - lse) - (l == y ? 1.f : 0.f)) * inv_batch;
      if (y >= 0 && y < L) atomicAdd(loss_acc, (double)(lse - z[y]) * (double)inv_batch);
    } else {
      for (int l = 0; l < L; ++l) s_dl[wv][l] = dlogits_in[r * L + l];
    }
    for (int l = 0; l < L; ++l) dlogits[r * L + l] = s_dl[wv][l];
  }
  __syncwarp();
  if (r >= rows) return;
#pragma unroll
  for (int i = 0; i < LS_MLP / 32; ++i) {
    const int n = i * 32 + lane;
    float s = 0.f;
    for (int l = 0; l < L; ++l) s = fmaf(s_dl[wv][l], __ldg(w2 + l * LS_MLP + n), s);
    dz1[r * LS_MLP + n] = z1[r * LS_MLP + n] > 0.f ? s : 0.f;
  }
}

// head backward for the frame objective: CE (or given dlogits) -> dz1 (masked by ReLU) -> dh; 16 rows per CTA
__global__ void __launch_bounds__(256) lstm_head_bwd_kernel(const float* __restrict__ logits, const int64_t* __restrict__ labels,
                                                            const float* __restrict__ dlogits_in, const float* __restrict__ z1,
                                                            const float* __restrict__ w1, const float* __restrict__ w2,
                                                            float* __restrict__ dlogits, float* __restrict__ dz1,
                                                            float* __restrict__ dh, double* __restrict__ loss_acc,
                                                            int64_t B, int L, float inv_batch) {
  __shared__ float s_dl[16 * 96];
  __shared__ float s_dz[16 * LS_MLP];
  const int tid = threadIdx.x;
  const int64_t r0 = (int64_t)blockIdx.x * 16;
  if (tid < 16) {
    const int64_t b = r0 + tid;
    if (b < B) {
      if (labels) {
        const float* z = logits + b * L;
        float mx = z[0];
        for (int l = 1; l < L; ++l) mx = fmaxf(mx, z[l]);
        float se = 0.f;
        for (int l = 0; l < L; ++l) se += expf(z[l] - mx);
        const float lse = mx + logf(se);
        const int64_t y = labels[b];
        for (int l = 0; l < L; ++l) s_dl[tid * 96 + l] = (expf(z[l] - lse) - (l == y ? 1.f : 0.f)) * inv_batch;
        if (y >= 0 && y < L) atomicAdd(loss_acc, (double)(lse - z[y]) * (double)inv_batch);
      } else {
        for (int l = 0; l < L; ++l) s_dl[tid * 96 + l] = dlogits_in[b * L + l];
      }
      for (int l = 0; l < L; ++l) dlogits[b * L + l] = s_dl[tid * 96 + l];
    } else {
      for (int l = 0; l < L; ++l) s_dl[tid * 96 + l] = 0.f;
    }
  }
  __syncthreads();
  // dz1[r][n] (thread = n)
  for (int r = 0; r < 16; ++r) {
    float s = 0.f;
    for (int l = 0; l < L; ++l) s = fmaf(s_dl[r * 96 + l], __ldg(w2 + l * LS_MLP + tid), s);
    const bool ok = r0 + r < B;
    const float zz = ok ? z1[(r0 + r) * LS_MLP + tid] : 0.f;
    const float d = zz > 0.f ? s : 0.f;
    s_dz[r * LS_MLP + tid] = d;
    if (ok) dz1[(r0 + r) * LS_MLP + tid] = d;
  }
  __syncthreads();
  // dh[r][k]: threads 0..127 rows 0..7, threads 128..255 rows 8..15
  const int k = tid & (LS_H - 1), rh = tid >> 7;
  float acc[8];
#pragma unroll
  for (int r = 0; r < 8; ++r) acc[r] = 0.f;
  for (int n = 0; n < LS_MLP; ++n) {
    const float w = __ldg(w1 + n * LS_H + k);
#pragma unroll
    for (int r = 0; r < 8; ++r) acc[r] = fmaf(w, s_dz[(rh * 8 + r) * LS_MLP + n], acc[r]);
  }
#pragma unroll
  for (int r = 0; r < 8; ++r)
    if (r0 + rh * 8 + r < B) dh[(r0 + rh * 8 + r) * LS_H + k] = acc[r];
}

// =============================================================================================
// backward recurrence (BPTT); leaves the pre-activation gate gradients in `gates`
// =============================================================================================
struct LstmBwdArgs {
  const int64_t* lengths;
  const float* w_hh;      // [512][128]
  const float* c0;
  float* gates;           // in: activated gates, out: d(pre-activation)
  const float* cs;
  const float* dh_head;   // [B][128] gradient at h_n (frame objective), or null
  const float* dh_seq;    // [T][B][128] gradient at every h_t (sequential objective), or null
  int64_t B;
  int T;
};

__global__ void __launch_bounds__(LS_THREADS, 1) lstm_bwd_kernel(const LstmBwdArgs a) {
  extern __shared__ __align__(16) float smem[];
  float* s_da = smem;                          // [512][16]
  float* s_dh = s_da + LS_G * LS_NB;           // [16][128]
  float* s_dc = s_dh + LS_NB * LS_H;           // [16][128]
  float* s_part = s_dc + LS_NB * LS_H;         // [4][16][128]
  __shared__ int s_len[LS_NB];
  const int tid = threadIdx.x;
  const int64_t b0 = (int64_t)blockIdx.x * LS_NB;
  if (tid < LS_NB) s_len[tid] = (b0 + tid < a.B) ? (int)a.lengths[b0 + tid] : 0;
  for (int p = tid; p < LS_NB * LS_H; p += LS_THREADS) s_dh[p] = s_dc[p] = 0.f;
  __syncthreads();
  const int kk = tid & (LS_H - 1), part = tid >> 7;
  for (int t = a.T - 1; t >= 0; --t) {
    for (int p = tid; p < LS_NB * LS_H; p += LS_THREADS) {
      const int b = p / LS_H, u = p - b * LS_H;
      const bool inb = b0 + b < a.B;
      const bool live = t < s_len[b];
      float dht = s_dh[p];
      if (inb && a.dh_head && t == s_len[b] - 1) dht += a.dh_head[(b0 + b) * LS_H + u];
      if (inb && a.dh_seq && live) dht += a.dh_seq[((size_t)t * a.B + b0 + b) * LS_H + u];
      float dai = 0.f, daf = 0.f, dag = 0.f, dao = 0.f;
      if (inb && live) {
        const size_t row = (size_t)t * a.B + b0 + b;
        float* g = a.gates + row * LS_G;
        const float gi = g[u], gf = g[LS_H + u], gg = g[2 * LS_H + u], go = g[3 * LS_H + u];
        const float c_t = a.cs[row * LS_H + u];
        const float c_prev = t > 0 ? a.cs[(row - a.B) * LS_H + u] : a.c0[(b0 + b) * LS_H + u];
        const float tc = tanhf(c_t);
        const float dtc = dht * go * (1.f - tc * tc) + s_dc[p];
        dao = dht * tc * go * (1.f - go);
        dai = dtc * gg * gi * (1.f - gi);
        dag = dtc * gi * (1.f - gg * gg);
        daf = dtc * c_prev * gf * (1.f - gf);
        s_dc[p] = dtc * gf;
        g[u] = dai; g[LS_H + u] = daf; g[2 * LS_H + u] = dag; g[3 * LS_H + u] = dao;
      } else if (inb) {
        float* g = a.gates + ((size_t)t * a.B + b0 + b) * LS_G;
        g[u] = 0.f; g[LS_H + u] = 0.f; g[2 * LS_H + u] = 0.f; g[3 * LS_H + u] = 0.f;
      }
      s_da[u * LS_NB + b] = dai;
      s_da[(LS_H + u) * LS_NB + b] = daf;
      s_da[(2 * LS_H + u) * LS_NB + b] = dag;
      s_da[(3 * LS_H + u) * LS_NB + b] = dao;
      s_dh[p] = live ? 0.f : dht;      // dead steps pass the gradient straight through; live ones go through W_hh below
    }
    __syncthreads();
    // dh_prev[b][k] += sum_j W_hh[j][k] da[b][j]: thread (part, k) covers j in [128 part, 128 part + 128)
    float acc[LS_NB];
#pragma unroll
    for (int b = 0; b < LS_NB; ++b) acc[b] = 0.f;
#pragma unroll 4
    for (int j = part * LS_H; j < (part + 1) * LS_H; ++j) {
      const float w = __ldg(a.w_hh + (size_t)j * LS_H + kk);
      const float4* d4 = reinterpret_cast<const float4*>(s_da + j * LS_NB);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float4 d = d4[q];
        acc[4 * q] = fmaf(w, d.x, acc[4 * q]);
        acc[4 * q + 1] = fmaf(w, d.y, acc[4 * q + 1]);
        acc[4 * q + 2] = fmaf(w, d.z, acc[4 * q + 2]);
        acc[4 * q + 3] = fmaf(w, d.w, acc[4 * q + 3]);
      }
    }
#pragma unroll
    for (int b = 0; b < LS_NB; ++b) s_part[(part * LS_NB + b) * LS_H + kk] = acc[b];
    __syncthreads();
    for (int p = tid; p < LS_NB * LS_H; p += LS_THREADS)
      s_dh[p] += s_part[p] + s_part[LS_NB * LS_H + p] + s_part[2 * LS_NB * LS_H + p] + s_part[3 * LS_NB * LS_H + p];
    __syncthreads();
  }
}


// ---------------------------------------------------------------------------------------------
// backward recurrence, software pipelined (option "lstm_engine" = 1 / 2); same arithmetic and summation order as
// lstm_bwd_kernel.  The saved gates / cell states of step t - 1 are fetched from HBM while step t multiplies (c_t of step t - 1
// is c_prev of step t: carried in a register), the W_hh rows stream through the register double buffer of the forward, the
// four partial sums of dh are added by the thread that consumes them (one barrier less per step), [j][sequence] tile padded to 20.
// ---------------------------------------------------------------------------------------------
template <int RT>   // register tile of the dh product: 1 = (part, k) x 16 sequences, 2 = (part, k and k + 64) x 8 sequences
__global__ void __launch_bounds__(LS_THREADS, 1) lstm_bwd_pipe_kernel(const LstmBwdArgs a) {
  extern __shared__ __align__(16) float smem[];
  float* s_da = smem;                          // [512][20]
  float* s_dh = s_da + LS_G * LS_NBP;          // [16][128]
  float* s_dc = s_dh + LS_NB * LS_H;           // [16][128]
  float* s_part = s_dc + LS_NB * LS_H;         // [4][16][128]
  __shared__ int s_len[LS_NB];
  const int tid = threadIdx.x;
  const int64_t b0 = (int64_t)blockIdx.x * LS_NB;
  if (tid < LS_NB) s_len[tid] = (b0 + tid < a.B) ? (int)a.lengths[b0 + tid] : 0;
  for (int p = tid; p < LS_NB * LS_H; p += LS_THREADS) s_dh[p] = s_dc[p] = 0.f;
  for (int p = tid; p < 4 * LS_NB * LS_H; p += LS_THREADS) s_part[p] = 0.f;
  __syncthreads();
  const int kk = tid & (LS_H - 1), part = tid >> 7;
  // this thread's four (sequence, unit) pairs: sequence part + 4 i, unit kk
  constexpr int NP = LS_NB * LS_H / LS_THREADS;
  float r_g[NP][4], r_ct[NP], r_cp[NP], r_dseq[NP], r_dhead[NP];
  int len[NP];
#pragma unroll
  for (int i = 0; i < NP; ++i) {
    const int b = part + 4 * i;
    len[i] = s_len[b];
    r_dhead[i] = (b0 + b < a.B && a.dh_head) ? a.dh_head[(b0 + b) * LS_H + kk] : 0.f;
    r_ct[i] = r_cp[i] = r_dseq[i] = 0.f;
    r_g[i][0] = r_g[i][1] = r_g[i][2] = r_g[i][3] = 0.f;
  }
  // saved tensors of step t (valid for every t < T and every sequence of the batch, live or not)
  auto fetch = [&](int t, bool first) {
#pragma unroll
    for (int i = 0; i < NP; ++i) {
      const int b = part + 4 * i;
      const bool inb = b0 + b < a.B;
      if (inb && t >= 0) {
        const size_t row = (size_t)t * a.B + b0 + b;
        const float* g = a.gates + row * LS_G;
        r_g[i][0] = g[kk]; r_g[i][1] = g[LS_H + kk]; r_g[i][2] = g[2 * LS_H + kk]; r_g[i][3] = g[3 * LS_H + kk];
        if (first) r_ct[i] = a.cs[row * LS_H + kk];
        r_cp[i] = t > 0 ? a.cs[(row - a.B) * LS_H + kk] : a.c0[(b0 + b) * LS_H + kk];
        r_dseq[i] = a.dh_seq ? a.dh_seq[row * LS_H + kk] : 0.f;
      }
    }
  };
  fetch(a.T - 1, true);
  // product mapping.  RT = 1: column kk, sequences 0..15.  RT = 2: columns kA and kA + 64 (the part's warp q: 16 q .. 16 q + 15), sequences sq0 .. sq0 + 7
  const int kA = RT == 1 ? kk : ((tid >> 5) & 3) * 16 + (tid & 15);
  const int sq0 = RT == 1 ? 0 : 8 * ((tid >> 4) & 1);
  const float* wp = a.w_hh + (size_t)(part * LS_H) * LS_H + kA;
  float w0[RT][LS_WB];
#pragma unroll
  for (int i = 0; i < LS_WB; ++i) {
    w0[0][i] = __ldg(wp + (size_t)i * LS_H);
    if (RT == 2) w0[RT - 1][i] = __ldg(wp + (size_t)i * LS_H + LS_H / 2);
  }
  for (int t = a.T - 1; t >= 0; --t) {
#pragma unroll
    for (int i = 0; i < NP; ++i) {
      const int b = part + 4 * i, p = b * LS_H + kk;
      const bool inb = b0 + b < a.B;
      const bool live = t < len[i];
      // dh_{t} = pass-through part + the four partial products of step t + 1 (same order as the separate reduction pass had)
      float dht = s_dh[p] + (s_part[p] + s_part[LS_NB * LS_H + p] + s_part[2 * LS_NB * LS_H + p] + s_part[3 * LS_NB * LS_H + p]);
      if (inb && a.dh_head && t == len[i] - 1) dht += r_dhead[i];
      if (inb && a.dh_seq && live) dht += r_dseq[i];
      float dai = 0.f, daf = 0.f, dag = 0.f, dao = 0.f;
      if (inb && live) {
        float* g = a.gates + ((size_t)t * a.B + b0 + b) * LS_G;
        const float gi = r_g[i][0], gf = r_g[i][1], gg = r_g[i][2], go = r_g[i][3];
        const float c_t = r_ct[i], c_prev = r_cp[i];
        const float tc = tanhf(c_t);
        const float dtc = dht * go * (1.f - tc * tc) + s_dc[p];
        dao = dht * tc * go * (1.f - go);
        dai = dtc * gg * gi * (1.f - gi);
        dag = dtc * gi * (1.f - gg * gg);
        daf = dtc * c_prev * gf * (1.f - gf);
        s_dc[p] = dtc * gf;
        g[kk] = dai; g[LS_H + kk] = daf; g[2 * LS_H + kk] = dag; g[3 * LS_H + kk] = dao;
      } else if (inb) {
        float* g = a.gates + ((size_t)t * a.B + b0 + b) * LS_G;
        g[kk] = 0.f; g[LS_H + kk] = 0.f; g[2 * LS_H + kk] = 0.f; g[3 * LS_H + kk] = 0.f;
      }
      s_da[kk * LS_NBP + b] = dai;
      s_da[(LS_H + kk) * LS_NBP + b] = daf;
      s_da[(2 * LS_H + kk) * LS_NBP + b] = dag;
      s_da[(3 * LS_H + kk) * LS_NBP + b] = dao;
      s_dh[p] = live ? 0.f : dht;      // dead steps pass the gradient straight through; live ones go through W_hh below
      r_ct[i] = r_cp[i];               // c_{t-1}: the cell state of the next step to be visited
    }
    fetch(t - 1, false);               // in flight during the product below
    __syncthreads();
    // dh_prev[b][k] += sum_j W_hh[j][k] da[b][j]: thread (part, k) covers j in [128 part, 128 part + 128)
    if constexpr (RT == 1) {
      float acc[LS_NB];
#pragma unroll
      for (int b = 0; b < LS_NB; ++b) acc[b] = 0.f;
      lstm_dot_pipelined(acc, wp, LS_H, s_da + (part * LS_H) * LS_NBP, LS_H, w0[0]);
#pragma unroll
      for (int b = 0; b < LS_NB; ++b) s_part[(part * LS_NB + b) * LS_H + kk] = acc[b];
    } else {
      float acc[2][8];
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[0][j] = acc[1][j] = 0.f;
      lstm_dot2x8_pipelined(acc, wp, LS_H, LS_H / 2, s_da + (part * LS_H) * LS_NBP + sq0, LS_H, w0);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        s_part[(part * LS_NB + sq0 + j) * LS_H + kA] = acc[0][j];
        s_part[(part * LS_NB + sq0 + j) * LS_H + kA + LS_H / 2] = acc[1][j];
      }
    }
    __syncthreads();
  }
}

// =============================================================================================
// CTC (nn.CTCLoss(blank), reduction 'mean', on log_softmax(scores); training/run/train.py:253,296-298).
// One warp per sequence: log-space forward variables (kept in the workspace), backward variables on the fly, and the
// gradient with respect to the SCORES (log_softmax folded in):
//   d(-ln P)/d z_t(k) = softmax_t(k) - sum_{s: l'_s = k} exp(alpha_t(s) + beta_t(s) - logp_t(k) - ln P)
// scaled by 1 / (max(U, 1) * batch).  Frames beyond the input length get zero gradient.
// =============================================================================================
__device__ __forceinline__ float log_add(float a, float b) {
  if (a == -INFINITY) return b;
  if (b == -INFINITY) return a;
  const float m = fmaxf(a, b);
  return m + log1pf(expf(-fabsf(a - b)));
}

__global__ void __launch_bounds__(128) ctc_kernel(const float* __restrict__ scores, const int64_t* __restrict__ targets,
                                                  const int64_t* __restrict__ tgt_len, const int64_t* __restrict__ in_len,
                                                  int blank, int T, int64_t B, int L, int Lmax, float* __restrict__ alpha_ws,
                                                  float* __restrict__ dscores, double* __restrict__ loss_acc,
                                                  float inv_batch) {
  __shared__ float s_a[4][2][CTC_SMAX];
  __shared__ float s_logp[4][96];
  __shared__ float s_acc[4][96];
  __shared__ int s_lab[4][CTC_SMAX];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int64_t b = (int64_t)blockIdx.x * 4 + w;
  if (b >= B) return;
  int U = (int)tgt_len[b];
  if (U > CTC_UMAX) U = CTC_UMAX;
  const int S = 2 * U + 1;
  int Tb = (int)in_len[b];
  if (Tb > T) Tb = T;
  for (int s = lane; s < CTC_SMAX; s += 32) s_lab[w][s] = (s < S && (s & 1)) ? (int)targets[b * Lmax + (s >> 1)] : blank;
  __syncwarp();
  float* my_alpha = alpha_ws + (size_t)b * T * CTC_SMAX;
  auto frame_logp = [&](int t) {   // log_softmax of frame t into s_logp
    const float* z = scores + ((size_t)t * B + b) * L;
    float mx = -INFINITY;
    for (int k = lane; k < L; k += 32) mx = fmaxf(mx, z[k]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float se = 0.f;
    for (int k = lane; k < L; k += 32) se += expf(z[k] - mx);
    se = warp_sum(se);
    const float lse = mx + logf(se);
    __syncwarp();
    for (int k = lane; k < L; k += 32) s_logp[w][k] = z[k] - lse;
    __syncwarp();
  };
  // ---- forward
  float logP = -INFINITY;
  if (Tb > 0) {
    frame_logp(0);
    for (int s = lane; s < S; s += 32) {
      const float v = (s < 2) ? s_logp[w][s_lab[w][s]] : -INFINITY;
      s_a[w][0][s] = v;
      my_alpha[s] = v;
    }
    __syncwarp();
    for (int t = 1; t < Tb; ++t) {
      frame_logp(t);
      const float* prev = s_a[w][(t - 1) & 1];
      float* cur = s_a[w][t & 1];
      for (int s = lane; s < S; s += 32) {
        float v = prev[s];
        if (s >= 1) v = log_add(v, prev[s - 1]);
        if (s >= 2 && (s & 1) && s_lab[w][s] != s_lab[w][s - 2]) v = log_add(v, prev[s - 2]);
        v += s_logp[w][s_lab[w][s]];
        cur[s] = v;
        my_alpha[(size_t)t * CTC_SMAX + s] = v;
      }
      __syncwarp();
    }
    const float* last = s_a[w][(Tb - 1) & 1];
    logP = last[S - 1];
    if (S >= 2) logP = log_add(logP, last[S - 2]);
  }
  const float scale = inv_batch / (float)(U > 0 ? U : 1);
  if (lane == 0) atomicAdd(loss_acc, (double)(-logP) * (double)scale);
  // ---- frames beyond the input length: zero gradient
  for (int t = Tb; t < T; ++t)
    for (int k = lane; k < L; k += 32) dscores[((size_t)t * B + b) * L + k] = 0.f;
  // ---- backward + gradient
  float* bcur = s_a[w][0];
  float* bnext = s_a[w][1];
  for (int t = Tb - 1; t >= 0; --t) {
    frame_logp(t);
    for (int s = lane; s < S; s += 32) {
      float v;
      if (t == Tb - 1) {
        v = (s >= S - 2) ? 0.f : -INFINITY;
      } else {
        v = bnext[s];
        if (s + 1 < S) v = log_add(v, bnext[s + 1]);
        if (s + 2 < S && (s & 1) && s_lab[w][s + 2] != s_lab[w][s]) v = log_add(v, bnext[s + 2]);
      }
      bcur[s] = v + s_logp[w][s_lab[w][s]];
    }
    for (int k = lane; k < L; k += 32) s_acc[w][k] = 0.f;
    __syncwarp();
    for (int s = lane; s < S; s += 32) {
      const int k = s_lab[w][s];
      const float e = my_alpha[(size_t)t * CTC_SMAX + s] + bcur[s] - s_logp[w][k] - logP;
      if (e > -80.f) atomicAdd(&s_acc[w][k], expf(e));
    }
    __syncwarp();
    for (int k = lane; k < L; k += 32)
      dscores[((size_t)t * B + b) * L + k] = (expf(s_logp[w][k]) - s_acc[w][k]) * scale;
    __syncwarp();
    float* tmp = bcur;
    bcur = bnext;
    bnext = tmp;
  }
}

// =============================================================================================
// out[m][n] (+)= sum_r A[r][m] * Bm[r][n]   (tall-skinny A^T B, split over r with fp32 atomics)
// 64 x 64 output tile per CTA, 4 x 4 per thread, 32-row slabs staged in shared memory
// =============================================================================================
__global__ void __launch_bounds__(256) lstm_atb_kernel(const float* __restrict__ A, int lda, const float* __restrict__ Bm,
                                                       int ldb, float* __restrict__ out, int ldo, int Mo, int No,
                                                       int64_t R, int64_t rows_per_cta) {
  __shared__ __align__(16) float sA[32][64 + 4];
  __shared__ __align__(16) float sB[32][64 + 4];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.x * 64, n0 = blockIdx.y * 64;
  const int64_t r_begin = (int64_t)blockIdx.z * rows_per_cta;
  int64_t r_end = r_begin + rows_per_cta;
  if (r_end > R) r_end = R;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int64_t r = r_begin; r < r_end; r += 32) {
    for (int p = tid; p < 32 * 64; p += 256) {
      const int rr = p >> 6, c = p & 63;
      const bool okr = r + rr < r_end;
      sA[rr][c] = (okr && m0 + c < Mo) ? A[(r + rr) * lda + m0 + c] : 0.f;
      sB[rr][c] = (okr && n0 + c < No) ? Bm[(r + rr) * ldb + n0 + c] : 0.f;
    }
    __syncthreads();
#pragma unroll 8
    for (int k = 0; k < 32; ++k) {
      const float4 a4 = *reinterpret_cast<const float4*>(&sA[k][ty * 4]);
      const float4 b4 = *reinterpret_cast<const float4*>(&sB[k][tx * 4]);
      const float av[4] = {a4.x, a4.y, a4.z, a4.w}, bv[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int m = m0 + ty * 4 + i, n = n0 + tx * 4 + j;
      if (m < Mo && n < No) atomicAdd(&out[(size_t)m * ldo + n], acc[i][j]);
    }
}

// out[m] (+ out2[m]) += sum_r A[r][m]
__global__ void __launch_bounds__(256) lstm_colsum_kernel(const float* __restrict__ A, int lda, int Mo, int64_t R,
                                                          int64_t rows_per_cta, float* __restrict__ out,
                                                          float* __restrict__ out2) {
  const int m = blockIdx.x * 256 + threadIdx.x;
  const int64_t r_begin = (int64_t)blockIdx.y * rows_per_cta;
  int64_t r_end = r_begin + rows_per_cta;
  if (r_end > R) r_end = R;
  if (m >= Mo) return;
  float s = 0.f;
  for (int64_t r = r_begin; r < r_end; ++r) s += A[r * lda + m];
  atomicAdd(&out[m], s);
  if (out2) atomicAdd(&out2[m], s);
}

__global__ void lstm_loss_kernel(const double* __restrict__ acc, float* __restrict__ loss) { *loss = (float)(*acc); }

// =============================================================================================
// host side
// =============================================================================================
static int lstm_atb(howl_ctx_t* ctx, cudaStream_t st, const float* A, int lda, const float* Bm, int ldb, float* out, int ldo,
                    int Mo, int No, int64_t R) {
  const int gx = (Mo + 63) / 64, gy = (No + 63) / 64;
  int64_t splits = (int64_t)ctx->sm_count * 3 / (gx * gy);
  if (splits < 1) splits = 1;
  int64_t rows = howl_ceil_div(R, splits);
  rows = (rows + 31) / 32 * 32;
  const int64_t gz = howl_ceil_div(R, rows);
  lstm_atb_kernel<<<dim3(gx, gy, (unsigned)gz), 256, 0, st>>>(A, lda, Bm, ldb, out, ldo, Mo, No, R, rows);
  HOWL_LAUNCHED(ctx, "lstm_atb");
  return HOWL_OK;
}

// out[m][n] += sum_r A[r][m] Bm[r][n] on the tensor cores: both sides split into (hi, lo) bf16 operands, three products, fp32 accumulate.
// reuse_x: the X side (A) is already packed in ws.px_* (the gate gradients feed two products)
// colsum / colsum2: the column sums of A (bias gradients) are accumulated while A is packed -- or by lstm_colsum on the FFMA path
static int lstm_colsum(howl_ctx_t* ctx, cudaStream_t st, const float* A, int lda, int Mo, int64_t R, float* out, float* out2);
static int lstm_atb_tc(howl_ctx_t* ctx, cudaStream_t st, const LstmWs& ws, const float* A, int lda, const float* Bm, int ldb, float* out, int ldo,
                       int Mo, int No, int64_t R, bool reuse_x, float* colsum = nullptr, float* colsum2 = nullptr) {
  int rc;
  if (R < 1024) {      // a handful of row tiles: the FFMA kernels, exact fp32
    if (colsum && (rc = lstm_colsum(ctx, st, A, lda, Mo, R, colsum, colsum2))) return rc;
    return lstm_atb(ctx, st, A, lda, Bm, ldb, out, ldo, Mo, No, R);
  }
  if (!reuse_x && (rc = mbn_pack_split(ctx, st, A, lda, R, Mo, ws.px_hi, ws.px_lo, colsum, colsum2))) return rc;
  if ((rc = mbn_pack_split(ctx, st, Bm, ldb, R, No, ws.py_hi, ws.py_lo))) return rc;
  return mbn_atb3_packed(ctx, st, ws.px_hi, ws.px_lo, ws.py_hi, ws.py_lo, out, R, Mo, No, ldo);
}

static int lstm_colsum(howl_ctx_t* ctx, cudaStream_t st, const float* A, int lda, int Mo, int64_t R, float* out, float* out2) {
  const int gx = (Mo + 255) / 256;
  int64_t splits = (int64_t)ctx->sm_count * 2 / gx;
  if (splits < 1) splits = 1;
  const int64_t rows = howl_ceil_div(R, splits);
  lstm_colsum_kernel<<<dim3(gx, (unsigned)howl_ceil_div(R, rows)), 256, 0, st>>>(A, lda, Mo, R, rows, out, out2);
  HOWL_LAUNCHED(ctx, "lstm_colsum");
  return HOWL_OK;
}

static int lstm_check(howl_ctx_t* ctx, int64_t B, int frames, int M, int L, int T, const void* ws) {
  HOWL_REQUIRE(ctx, B >= 1, HOWL_E_INVALID, "lstm: empty batch");
  HOWL_REQUIRE(ctx, M >= 1 && M <= HOWL_MAX_MELS && (M % 4) == 0, HOWL_E_UNSUPPORTED, "lstm: n_mels=%d unsupported", M);
  HOWL_REQUIRE(ctx, L >= 1 && L <= 96, HOWL_E_UNSUPPORTED, "lstm: num_labels=%d outside 1..96", L);
  HOWL_REQUIRE(ctx, T >= 1 && T <= frames, HOWL_E_INVALID, "lstm: max_steps=%d must be in 1..frames=%d", T, frames);
  HOWL_REQUIRE(ctx, ws != nullptr, HOWL_E_WORKSPACE, "lstm: null workspace");
  return HOWL_OK;
}

extern "C" int howl_b200_lstm_fwd(howl_ctx_t* ctx, void* stream, const float* feats, const int64_t* lengths, int64_t B,
                                  int32_t frames, int32_t n_mels, int32_t num_labels, int32_t max_steps,
                                  const float* params, const float* state_in, float* state_out, int sequential,
                                  int train, float* out, void* workspace, size_t workspace_bytes) {
  if (!ctx) return HOWL_E_INVALID;
  HOWL_REQUIRE(ctx, feats && lengths && params && out, HOWL_E_INVALID, "lstm_fwd: null pointer");
  int rc = lstm_check(ctx, B, frames, n_mels, num_labels, max_steps, workspace);
  if (rc) return rc;
  const int M = n_mels, L = num_labels, T = max_steps, K = M + LS_H;
  LstmWs ws = lstm_carve(workspace, B, T, M, L, train, sequential);
  HOWL_REQUIRE(ctx, ws.bytes <= workspace_bytes, HOWL_E_WORKSPACE, "lstm_fwd: workspace %zu < required %zu", workspace_bytes,
               ws.bytes);
  cudaStream_t st = (cudaStream_t)stream;
  const LstmParams v = lstm_views(params, M, L);
  lstm_prep_kernel<<<64, 256, 0, st>>>(v, M, ws.wt, ws.bsum, ws.w1t);
  HOWL_LAUNCHED(ctx, "lstm_prep");
  if (state_in) {
    HOWL_CUDA(ctx, cudaMemcpyAsync(ws.h0, state_in, sizeof(float) * B * LS_H, cudaMemcpyDeviceToDevice, st));
    HOWL_CUDA(ctx, cudaMemcpyAsync(ws.c0, state_in + B * LS_H, sizeof(float) * B * LS_H, cudaMemcpyDeviceToDevice, st));
  } else {
    HOWL_CUDA(ctx, cudaMemsetAsync(ws.h0, 0, sizeof(float) * B * LS_H, st));
    HOWL_CUDA(ctx, cudaMemsetAsync(ws.c0, 0, sizeof(float) * B * LS_H, st));
  }
  LstmFwdArgs a;
  a.feats = feats; a.lengths = lengths; a.wt = ws.wt; a.bsum = ws.bsum; a.h0 = ws.h0; a.c0 = ws.c0;
  a.gates = train ? ws.gates : nullptr; a.cs = ws.cs; a.xh = train ? ws.xh : nullptr;
  a.hfin = ws.hfin; a.cfin = ws.cfin; a.hseq = sequential ? ws.hseq : nullptr;
  a.B = B; a.F = frames; a.M = M; a.T = T;
  if (ctx->lstm_engine == 0) {
    const size_t smem = sizeof(float) * ((size_t)K * LS_NB + LS_NB * LS_G + LS_NB * LS_H);
    HOWL_CUDA(ctx, cudaFuncSetAttribute(lstm_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    lstm_fwd_kernel<<<(unsigned)howl_ceil_div(B, LS_NB), LS_THREADS, smem, st>>>(a);
  } else {
    const size_t smem = sizeof(float) * ((size_t)K * LS_NBP + LS_NB * LS_G + LS_NB * LS_H);
    if (ctx->lstm_engine == 2) {
      HOWL_CUDA(ctx, cudaFuncSetAttribute(lstm_fwd_pipe_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      lstm_fwd_pipe_kernel<2><<<(unsigned)howl_ceil_div(B, LS_NB), LS_THREADS, smem, st>>>(a);
    } else {
      HOWL_CUDA(ctx, cudaFuncSetAttribute(lstm_fwd_pipe_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      lstm_fwd_pipe_kernel<1><<<(unsigned)howl_ceil_div(B, LS_NB), LS_THREADS, smem, st>>>(a);
    }
  }
  HOWL_LAUNCHED(ctx, "lstm_fwd");
  if (state_out) {
    HOWL_CUDA(ctx, cudaMemcpyAsync(state_out, ws.hfin, sizeof(float) * B * LS_H, cudaMemcpyDeviceToDevice, st));
    HOWL_CUDA(ctx, cudaMemcpyAsync(state_out + B * LS_H, ws.cfin, sizeof(float) * B * LS_H, cudaMemcpyDeviceToDevice, st));
  }
  const int64_t rows = sequential ? (int64_t)T * B : B;
  if (train && rows >= LH_TC_ROWS) {
    // z1 = relu(h W1^T + b1) on the tensor cores, then the L-wide scores
    int rc;
    if ((rc = mbn_weight_operand3(ctx, st, v.w1, LS_MLP, LS_H, LS_H, 0, ws.wop_w1))) return rc;
    if ((rc = mbn_pack3(ctx, st, sequential ? ws.hseq : ws.hfin, LS_H, rows, LS_H, ws.p3))) return rc;
    if ((rc = mbn_gemm_nt3_f32(ctx, st, ws.p3, ws.wop_w1, ws.z1, LS_MLP, rows, LS_H, LS_MLP, v.b1, 1))) return rc;
    lstm_scores_kernel<<<(unsigned)howl_ceil_div(rows, 8), 256, 0, st>>>(ws.z1, rows, v.w2, v.b2, out, ws.logits, L);
    HOWL_LAUNCHED(ctx, "lstm_scores");
    return HOWL_OK;
  }
  lstm_head_fwd_kernel<<<(unsigned)howl_ceil_div(rows, 16), 256, 0, st>>>(sequential ? ws.hseq : ws.hfin, rows, ws.w1t, v.b1,
                                                                          v.w2, v.b2, train ? ws.z1 : nullptr, out,
                                                                          (sequential && !train) ? nullptr : ws.logits, L);
  HOWL_LAUNCHED(ctx, "lstm_head_fwd");
  return HOWL_OK;
}

static int lstm_bwd_impl(howl_ctx_t* ctx, void* stream, const int64_t* lengths, const int64_t* labels,
                         const float* dlogits_in, int sequential, const int64_t* targets, const int64_t* tgt_len,
                         int max_target_len, int blank, int64_t B, int32_t frames, int32_t n_mels, int32_t num_labels,
                         int32_t max_steps, int64_t loss_scale_batch, const float* params, float* grads, float* loss,
                         void* workspace, size_t workspace_bytes) {
  int rc = lstm_check(ctx, B, frames, n_mels, num_labels, max_steps, workspace);
  if (rc) return rc;
  const int M = n_mels, L = num_labels, T = max_steps, K = M + LS_H;
  LstmWs ws = lstm_carve(workspace, B, T, M, L, 1, sequential);
  HOWL_REQUIRE(ctx, ws.bytes <= workspace_bytes, HOWL_E_WORKSPACE, "lstm_bwd: workspace %zu < required %zu", workspace_bytes,
               ws.bytes);
  cudaStream_t st = (cudaStream_t)stream;
  const LstmParams v = lstm_views(params, M, L);
  const int64_t nparam = howl_b200_lstm_param_count(L, M);
  const int64_t rows = sequential ? (int64_t)T * B : B;
  float* g_w_ih = grads;
  float* g_w_hh = g_w_ih + (size_t)LS_G * M;
  float* g_b_ih = g_w_hh + (size_t)LS_G * LS_H;
  float* g_b_hh = g_b_ih + LS_G;
  float* g_w1 = g_b_hh + LS_G;
  float* g_b1 = g_w1 + (size_t)LS_MLP * LS_H;
  float* g_w2 = g_b1 + LS_MLP;
  float* g_b2 = g_w2 + (size_t)L * LS_MLP;
  HOWL_CUDA(ctx, cudaMemsetAsync(grads, 0, sizeof(float) * nparam, st));
  HOWL_CUDA(ctx, cudaMemsetAsync(ws.loss_acc, 0, sizeof(double) * 2, st));
  const float inv_batch = 1.f / (float)loss_scale_batch;
  if (sequential && targets) {
    HOWL_REQUIRE(ctx, max_target_len >= 1 && max_target_len <= CTC_UMAX, HOWL_E_UNSUPPORTED,
                 "ctc: max target length %d outside 1..%d", max_target_len, CTC_UMAX);
    HOWL_REQUIRE(ctx, blank >= 0 && blank < L, HOWL_E_INVALID, "ctc: blank %d outside the %d labels", blank, L);
    ctc_kernel<<<(unsigned)howl_ceil_div(B, 4), 128, 0, st>>>(ws.logits, targets, tgt_len, lengths, blank, T, B, L,
                                                               max_target_len, ws.alpha, ws.dlogits, ws.loss_acc, inv_batch);
    HOWL_LAUNCHED(ctx, "ctc");
    dlogits_in = ws.dlogits;
  }
  if (rows >= LH_TC_ROWS) {
    // dlogits, dz1 (L-wide), then dh = dz1 W1 on the tensor cores
    lstm_dz1_kernel<<<(unsigned)howl_ceil_div(rows, 8), 256, 0, st>>>(ws.logits, labels, dlogits_in, ws.z1, v.w2, ws.dlogits, ws.dz1, ws.loss_acc,
                                                                      rows, L, inv_batch);
    HOWL_LAUNCHED(ctx, "lstm_dz1");
    if ((rc = mbn_weight_operand3(ctx, st, v.w1, LS_H, LS_MLP, LS_H, 1, ws.wop_w1t))) return rc;
    if ((rc = mbn_pack3(ctx, st, ws.dz1, LS_MLP, rows, LS_MLP, ws.p3))) return rc;
    if ((rc = mbn_gemm_nt3_f32(ctx, st, ws.p3, ws.wop_w1t, ws.dh, LS_H, rows, LS_MLP, LS_H, nullptr, 0))) return rc;
  } else {
    lstm_head_bwd_kernel<<<(unsigned)howl_ceil_div(rows, 16), 256, 0, st>>>(ws.logits, labels, dlogits_in, ws.z1, v.w1, v.w2,
                                                                            ws.dlogits, ws.dz1, ws.dh, ws.loss_acc, rows, L,
                                                                            inv_batch);
    HOWL_LAUNCHED(ctx, "lstm_head_bwd");
  }
  if (loss) {
    lstm_loss_kernel<<<1, 1, 0, st>>>(ws.loss_acc, loss);
    HOWL_LAUNCHED(ctx, "lstm_loss");
  }
  // head parameter gradients
  // (L output rows fill a sliver of a 128-row MMA tile: the FFMA kernel reads z1 once and is faster here)
  if ((rc = lstm_atb(ctx, st, ws.dlogits, L, ws.z1, LS_MLP, g_w2, LS_MLP, L, LS_MLP, rows))) return rc;
  if ((rc = lstm_colsum(ctx, st, ws.dlogits, L, L, rows, g_b2, nullptr))) return rc;
  if ((rc = lstm_atb_tc(ctx, st, ws, ws.dz1, LS_MLP, sequential ? ws.hseq : ws.hfin, LS_H, g_w1, LS_H, LS_MLP, LS_H, rows, false, g_b1))) return rc;
  // BPTT
  LstmBwdArgs a;
  a.lengths = lengths; a.w_hh = v.w_hh; a.c0 = ws.c0; a.gates = ws.gates; a.cs = ws.cs;
  a.dh_head = sequential ? nullptr : ws.dh; a.dh_seq = sequential ? ws.dh : nullptr; a.B = B; a.T = T;
  if (ctx->lstm_engine == 0) {
    const size_t smem = sizeof(float) * ((size_t)LS_G * LS_NB + 2 * LS_NB * LS_H + 4 * LS_NB * LS_H);
    HOWL_CUDA(ctx, cudaFuncSetAttribute(lstm_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    lstm_bwd_kernel<<<(unsigned)howl_ceil_div(B, LS_NB), LS_THREADS, smem, st>>>(a);
  } else {
    const size_t smem = sizeof(float) * ((size_t)LS_G * LS_NBP + 2 * LS_NB * LS_H + 4 * LS_NB * LS_H);
    if (ctx->lstm_engine == 2) {
      HOWL_CUDA(ctx, cudaFuncSetAttribute(lstm_bwd_pipe_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      lstm_bwd_pipe_kernel<2><<<(unsigned)howl_ceil_div(B, LS_NB), LS_THREADS, smem, st>>>(a);
    } else {
      HOWL_CUDA(ctx, cudaFuncSetAttribute(lstm_bwd_pipe_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      lstm_bwd_pipe_kernel<1><<<(unsigned)howl_ceil_div(B, LS_NB), LS_THREADS, smem, st>>>(a);
    }
  }
  HOWL_LAUNCHED(ctx, "lstm_bwd");
  // weight gradients over all (t, b) rows: [W_ih | W_hh] from xh = [x_t | h_{t-1}]
  const int64_t R = (int64_t)T * B;
  if ((rc = lstm_atb_tc(ctx, st, ws, ws.gates, LS_G, ws.xh, K, g_w_ih, M, LS_G, M, R, false, g_b_ih, g_b_hh))) return rc;
  if ((rc = lstm_atb_tc(ctx, st, ws, ws.gates, LS_G, ws.xh + M, K, g_w_hh, LS_H, LS_G, LS_H, R, true))) return rc;
  return HOWL_OK;
}

extern "C" int howl_b200_lstm_bwd(howl_ctx_t* ctx, void* stream, const int64_t* lengths, const int64_t* labels, int64_t B,
                                  int32_t frames, int32_t n_mels, int32_t num_labels, int32_t max_steps,
                                  int64_t loss_scale_batch, const float* params, float* grads, float* loss,
                                  void* workspace, size_t workspace_bytes) {
  if (!ctx) return HOWL_E_INVALID;
  HOWL_REQUIRE(ctx, lengths && labels && params && grads && loss, HOWL_E_INVALID, "lstm_bwd: null pointer");
  HOWL_REQUIRE(ctx, loss_scale_batch >= 1, HOWL_E_INVALID, "lstm_bwd: loss_scale_batch must be >= 1");
  return lstm_bwd_impl(ctx, stream, lengths, labels, nullptr, 0, nullptr, nullptr, 0, 0, B, frames, n_mels, num_labels,
                       max_steps, loss_scale_batch, params, grads, loss, workspace, workspace_bytes);
}

extern "C" int howl_b200_lstm_bwd_dlogits(howl_ctx_t* ctx, void* stream, const int64_t* lengths, const float* dlogits,
                                          int sequential, int64_t B, int32_t frames, int32_t n_mels, int32_t num_labels,
                                          int32_t max_steps, const float* params, float* grads, void* workspace,
                                          size_t workspace_bytes) {
  if (!ctx) return HOWL_E_INVALID;
  HOWL_REQUIRE(ctx, lengths && dlogits && params && grads, HOWL_E_INVALID, "lstm_bwd_dlogits: null pointer");
  return lstm_bwd_impl(ctx, stream, lengths, nullptr, dlogits, sequential, nullptr, nullptr, 0, 0, B, frames, n_mels,
                       num_labels, max_steps, 1, params, grads, nullptr, workspace, workspace_bytes);
}

extern "C" int howl_b200_lstm_ctc_bwd(howl_ctx_t* ctx, void* stream, const int64_t* lengths, const int64_t* targets,
                                      const int64_t* target_lengths, int32_t max_target_len, int32_t blank, int64_t B,
                                      int32_t frames, int32_t n_mels, int32_t num_labels, int32_t max_steps,
                                      int64_t loss_scale_batch, const float* params, float* grads, float* loss,
                                      void* workspace, size_t workspace_bytes) {
  if (!ctx) return HOWL_E_INVALID;
  HOWL_REQUIRE(ctx, lengths && targets && target_lengths && params && grads && loss, HOWL_E_INVALID, "lstm_ctc_bwd: null pointer");
  HOWL_REQUIRE(ctx, loss_scale_batch >= 1, HOWL_E_INVALID, "lstm_ctc_bwd: loss_scale_batch must be >= 1");
  return lstm_bwd_impl(ctx, stream, lengths, nullptr, nullptr, 1, targets, target_lengths, max_target_len, blank, B, frames,
                       n_mels, num_labels, max_steps, loss_scale_batch, params, grads, loss, workspace, workspace_bytes);
}

extern "C" int howl_b200_seq_lstm_ctc_train_step(howl_ctx_t* ctx, void* stream, const float* pcm, const int64_t* targets,
                                                 const int64_t* target_lengths, int32_t max_target_len, int32_t blank,
                                                 const int64_t* lengths, int64_t B, int64_t T, const float* fb,
                                                 float zmuv_mean, float zmuv_std, int32_t num_labels, int32_t max_steps,
                                                 float* params, float* state, float* grads, float* exp_avg,
                                                 float* exp_avg_sq, int64_t step, float lr, float weight_decay,
                                                 float* loss, float* scores, void* workspace, size_t workspace_bytes) {
  if (!ctx) return HOWL_E_INVALID;
  HOWL_REQUIRE(ctx, workspace, HOWL_E_WORKSPACE, "seq_lstm_ctc_train_step: null workspace");
  const int M = ctx->fe.n_mels;
  const int64_t F = howl_b200_num_frames(T, ctx->fe.hop);
  HOWL_REQUIRE(ctx, F > 0 && F < (1 << 20), HOWL_E_INVALID, "seq_lstm_ctc_train_step: bad clip length %lld", (long long)T);
  const size_t feat_bytes = howl_align_up(sizeof(float) * (size_t)B * F * M, 256);
  HOWL_REQUIRE(ctx, workspace_bytes > feat_bytes, HOWL_E_WORKSPACE, "seq_lstm_ctc_train_step: workspace too small");
  float* feats = (float*)workspace;
  void* ws = (char*)workspace + feat_bytes;
  const size_t ws_bytes = workspace_bytes - feat_bytes;
  int rc = howl_b200_frontend_fwd(ctx, stream, pcm, B, T, fb, zmuv_mean, zmuv_std, nullptr, HOWL_FE_TIME_MAJOR | HOWL_FE_ZMUV,
                                  feats);
  if (rc) return rc;
  // streaming: the (detached) state of the previous call is the initial state and is replaced (rnn.py:62-68)
  rc = howl_b200_lstm_fwd(ctx, stream, feats, lengths, B, (int)F, M, num_labels, max_steps, params, state, state, 1, 1, scores,
                          ws, ws_bytes);
  if (rc) return rc;
  rc = howl_b200_lstm_ctc_bwd(ctx, stream, lengths, targets, target_lengths, max_target_len, blank, B, (int)F, M, num_labels,
                              max_steps, B, params, grads, loss, ws, ws_bytes);
  if (rc) return rc;
  return howl_b200_adamw(ctx, stream, params, grads, exp_avg, exp_avg_sq, howl_b200_lstm_param_count(num_labels, M), step, lr,
                         0.9f, 0.999f, 1e-8f, weight_decay);
}

extern "C" int howl_b200_lstm_train_step(howl_ctx_t* ctx, void* stream, const float* pcm, const int64_t* labels,
                                         const int64_t* lengths, int64_t B, int64_t T, const float* fb, float zmuv_mean,
                                         float zmuv_std, int32_t num_labels, int32_t max_steps, float* params,
                                         float* grads, float* exp_avg, float* exp_avg_sq, int64_t step, float lr,
                                         float weight_decay, float* loss, float* logits, void* workspace,
                                         size_t workspace_bytes) {
  if (!ctx) return HOWL_E_INVALID;
  HOWL_REQUIRE(ctx, workspace, HOWL_E_WORKSPACE, "lstm_train_step: null workspace");
  const int M = ctx->fe.n_mels;
  const int64_t F = howl_b200_num_frames(T, ctx->fe.hop);
  HOWL_REQUIRE(ctx, F > 0 && F < (1 << 20), HOWL_E_INVALID, "lstm_train_step: bad clip length %lld", (long long)T);
  const size_t feat_bytes = howl_align_up(sizeof(float) * (size_t)B * F * M, 256);
  HOWL_REQUIRE(ctx, workspace_bytes > feat_bytes, HOWL_E_WORKSPACE, "lstm_train_step: workspace too small");
  float* feats = (float*)workspace;
  void* ws = (char*)workspace + feat_bytes;
  const size_t ws_bytes = workspace_bytes - feat_bytes;
  int rc = howl_b200_frontend_fwd(ctx, stream, pcm, B, T, fb, zmuv_mean, zmuv_std, nullptr, HOWL_FE_TIME_MAJOR | HOWL_FE_ZMUV,
                                  feats);
  if (rc) return rc;
  rc = howl_b200_lstm_fwd(ctx, stream, feats, lengths, B, (int)F, M, num_labels, max_steps, params, nullptr, nullptr, 0, 1,
                          logits, ws, ws_bytes);
  if (rc) return rc;
  rc = howl_b200_lstm_bwd(ctx, stream, lengths, labels, B, (int)F, M, num_labels, max_steps, B, params, grads, loss, ws,
                          ws_bytes);
  if (rc) return rc;
  return howl_b200_adamw(ctx, stream, params, grads, exp_avg, exp_avg_sq, howl_b200_lstm_param_count(num_labels, M), step, lr,
                         0.9f, 0.999f, 1e-8f, weight_decay);
}
