// LASClassifier forward (howl/model/rnn.py:133-215) for sm_100a, exact fp32:
//   encoder: Conv2d(3, 8, 3, padding=2) + BatchNorm2d + ReLU + MaxPool2d((1, 2)), Conv2d(8, 8, 3, padding=2) + BatchNorm2d + ReLU + MaxPool2d((1, 2)),
//            [F'', B, 8 * 44] -> bidirectional LSTM(352 -> 96) over each clip's own length (pack_padded_sequence semantics, any order)
//   FixedAttentionModule: 4 heads, a fixed context vector scores the value projection, softmax over time (padding masked with -100),
//            the scores average the key projection;  fc: Linear(192 -> 256) + ReLU + Dropout (identity here) + Linear(256 -> L).
// and the autograd backward of all of it (CrossEntropyLoss(mean) or a caller's dlogits -> gradients in the flat parameter layout).
// Small convolutions and the attention are plain CUDA-core kernels; the recurrence is batch-parallel like K5 (lstm.cu): one CTA owns 16
// sequences of one direction, thread = gate row, [x_t | h] in shared memory, weights streamed from L2.  The backward keeps the recurrence
// kernel down to dh_{t-1} = da_t W_hh and hands the batched products (dx = da W_ih, dW_ih = da^T x, dW_hh = da^T h_{t-1}, the head's
// weight gradients) to one tiled fp32 GEMM kernel; reductions over the batch go through that GEMM or per-CTA partial sums + atomics.
#include <math.h>

#include <algorithm>

#include "common.cuh"
#include "mbn_common.cuh"

#define LA_C 8              // latent channels
#define LA_H 96             // LSTM hidden size
#define LA_G (4 * LA_H)     // gate rows (i, f, g, o)
#define LA_D (2 * LA_H)     // encoder output width
#define LA_HEADS 4
#define LA_DNN 256
#define LA_NB 16            // sequences per CTA of the recurrence
#define LA_EPS 1e-5
#define LA_MOM 0.1

struct LasDims {
  int M, F;                 // input mels x frames
  int h1, w1, w1p;          // conv1 output (M + 2) x (F + 2), pooled width
  int h2, w2, w2p;          // conv2 output (M + 4) x (w1p + 2), pooled width = LSTM steps
  int in;                   // LSTM input width 8 * h2
};
static LasDims las_dims(int n_mels, int frames) {
  LasDims d;
  d.M = n_mels; d.F = frames;
  d.h1 = n_mels + 2; d.w1 = frames + 2; d.w1p = d.w1 / 2;
  d.h2 = d.h1 + 2; d.w2 = d.w1p + 2; d.w2p = d.w2 / 2;
  d.in = LA_C * d.h2;
  return d;
}

// flat parameter layout = state_dict order of the trainable tensors (SURVEY App. B / tests/golden/las.npz):
//   encoder.conv1.{weight [8,3,3,3], bias}, encoder.conv2.{weight [8,8,3,3], bias}, encoder.conv_encoder.1.{weight, bias},
//   encoder.conv_encoder.5.{weight, bias}, encoder.lstm_encoder.{weight_ih_l0 [384,in], weight_hh_l0 [384,96], bias_ih_l0, bias_hh_l0,
//   *_reverse}, attn.context_vec [192], attn.v_proj.{weight [192,192], bias}, attn.k_proj.{weight, bias}, fc.0.{weight [256,192], bias},
//   fc.3.{weight [L,256], bias}
struct LasParams {
  const float *c1w, *c1b, *c2w, *c2b, *bn1g, *bn1b, *bn2g, *bn2b;
  const float *wih[2], *whh[2], *bih[2], *bhh[2];
  const float *cvec, *vw, *vb, *kw, *kb, *f0w, *f0b, *f3w, *f3b;
  int64_t total;
};
static LasParams las_params(const float* p, int in, int L) {
  LasParams q;
  const float* s = p;
  auto take = [&](size_t n) { const float* r = p; p += n; return r; };
  q.c1w = take(LA_C * 3 * 9); q.c1b = take(LA_C); q.c2w = take(LA_C * LA_C * 9); q.c2b = take(LA_C);
  q.bn1g = take(LA_C); q.bn1b = take(LA_C); q.bn2g = take(LA_C); q.bn2b = take(LA_C);
  for (int d = 0; d < 2; ++d) {
    q.wih[d] = take((size_t)LA_G * in); q.whh[d] = take((size_t)LA_G * LA_H); q.bih[d] = take(LA_G); q.bhh[d] = take(LA_G);
  }
  q.cvec = take(LA_D); q.vw = take(LA_D * LA_D); q.vb = take(LA_D); q.kw = take(LA_D * LA_D); q.kb = take(LA_D);
  q.f0w = take(LA_DNN * LA_D); q.f0b = take(LA_DNN); q.f3w = take((size_t)L * LA_DNN); q.f3b = take(L);
  q.total = p - s;
  return q;
}

struct LasWs {
  float* raw1;      // [B, 8, h1, w1]
  float* pool1;     // [B, 8, h1, w1p]
  float* raw2;      // [B, 8, h2, w2]
  float* x;         // [w2p, B, 8 * h2]   LSTM input (time major)
  float* wt;        // [2][in + 96][384]  transposed [W_ih | W_hh] per direction
  float* bsum;      // [2][384]
  float* hseq;      // [w2p, B, 192]
  double* stats;    // [2 layers][2][8]
  float* bn;        // [2 layers][4][8] scale, shift, mean, rstd
  float* u;         // [4, 192] + [4]: v_proj folded into the context vector (las_attn_prep_kernel)
  // ---- kept by the train-mode forward for the backward
  float* gates;     // [w2p, B, 2, 384] activated gates (i, f, g, o) of the live steps
  float* cseq;      // [w2p, B, 2, 96]  cell state after each live step
  float* scores;    // [B, w2p, 4]      attention scores
  float* ctxs;      // [B, 192]         attention context
  float* hidd;      // [B, 256]         fc hidden after ReLU and dropout
  float* logits;    // [B, L]
  // ---- backward scratch
  float* dlogits;   // [B, L]
  float* dctx;      // [B, 192]
  float* hbar;      // [B, 4, 192]  sum_t scores[t][h] h_t
  float* dhid;      // [B, 256]
  float* du;        // [B, 4, 192]
  float* dsum;      // [B, 4]
  float* red;       // [4 * 192 + 4]  batch sums of du, dsum
  float* dhseq;     // [w2p, B, 192]
  float* dgates;    // [w2p, B, 2, 384] gradients of the gate pre-activations
  float* dx;        // [w2p, B, 8 * h2]
  float* draw2;     // [B, 8, h2, w2]
  float* dpool1;    // [B, 8, h1, w1p]
  float* draw1;     // [B, 8, h1, w1]
  double* bstats;   // [2 layers][2][8]  sum dy, sum dy * xhat
  double* loss_acc; // [2]
  __nv_bfloat16 *px_hi, *px_lo;   // (hi, lo) bf16 operands of the LSTM weight-gradient products: [w2p * B][384] ...
  __nv_bfloat16 *py_hi, *py_lo;   // ... and [w2p * B][8 * h2]
  __nv_bfloat16* p3;              // [w2p * B][3 * 384]: [hi | hi | lo] of one direction's gate gradients (dx = da W_ih on the tensor cores)
  __nv_bfloat16* wop3;            // W_ih^T operand [8 * h2][3 * 384] / W_ih operand [384][3 * 8 * h2]
  float* xproj;                   // [w2p * B][2][384]: input projections of all steps (forward, large batches)
  size_t bytes;
};
static LasWs las_carve(void* base, int64_t B, const LasDims& d, int train = 0, int L = 0) {
  LasWs w;
  size_t off = 0;
  char* p = (char*)base;
  auto take = [&](size_t bytes) {
    void* r = p ? (void*)(p + off) : nullptr;
    off += howl_align_up(bytes, 256);
    return r;
  };
  w.raw1 = (float*)take(sizeof(float) * B * LA_C * d.h1 * d.w1);
  w.pool1 = (float*)take(sizeof(float) * B * LA_C * d.h1 * d.w1p);
  w.raw2 = (float*)take(sizeof(float) * B * LA_C * d.h2 * d.w2);
  w.x = (float*)take(sizeof(float) * (size_t)d.w2p * B * d.in);
  w.wt = (float*)take(sizeof(float) * 2 * (size_t)(d.in + LA_H) * LA_G);
  w.bsum = (float*)take(sizeof(float) * 2 * LA_G);
  w.hseq = (float*)take(sizeof(float) * (size_t)d.w2p * B * LA_D);
  w.stats = (double*)take(sizeof(double) * 2 * 2 * LA_C);
  w.bn = (float*)take(sizeof(float) * 2 * 4 * LA_C);
  w.u = (float*)take(sizeof(float) * (LA_HEADS * LA_D + LA_HEADS));
  if (train) {
    const size_t TB = (size_t)d.w2p * B;
    w.gates = (float*)take(sizeof(float) * TB * 2 * LA_G);
    w.cseq = (float*)take(sizeof(float) * TB * 2 * LA_H);
    w.scores = (float*)take(sizeof(float) * TB * LA_HEADS);
    w.ctxs = (float*)take(sizeof(float) * B * LA_D);
    w.hidd = (float*)take(sizeof(float) * B * LA_DNN);
    w.logits = (float*)take(sizeof(float) * B * L);
    w.dlogits = (float*)take(sizeof(float) * B * L);
    w.dctx = (float*)take(sizeof(float) * B * LA_D);
    w.hbar = (float*)take(sizeof(float) * B * LA_HEADS * LA_D);
    w.dhid = (float*)take(sizeof(float) * B * LA_DNN);
    w.du = (float*)take(sizeof(float) * B * LA_HEADS * LA_D);
    w.dsum = (float*)take(sizeof(float) * B * LA_HEADS);
    w.red = (float*)take(sizeof(float) * (LA_HEADS * LA_D + LA_HEADS));
    w.dhseq = (float*)take(sizeof(float) * TB * LA_D);
    w.dgates = (float*)take(sizeof(float) * TB * 2 * LA_G);
    w.dx = (float*)take(sizeof(float) * TB * d.in);
    w.draw2 = (float*)take(sizeof(float) * B * LA_C * d.h2 * d.w2);
    w.dpool1 = (float*)take(sizeof(float) * B * LA_C * d.h1 * d.w1p);
    w.draw1 = (float*)take(sizeof(float) * B * LA_C * d.h1 * d.w1);
    w.bstats = (double*)take(sizeof(double) * 2 * 2 * LA_C);
    w.loss_acc = (double*)take(sizeof(double) * 2);
    w.px_hi = (__nv_bfloat16*)take(mbn_tmo_bytes((int64_t)TB, LA_G));
    w.px_lo = (__nv_bfloat16*)take(mbn_tmo_bytes((int64_t)TB, LA_G));
    w.py_hi = (__nv_bfloat16*)take(mbn_tmo_bytes((int64_t)TB, d.in));
    w.py_lo = (__nv_bfloat16*)take(mbn_tmo_bytes((int64_t)TB, d.in));
    w.p3 = (__nv_bfloat16*)take(mbn_tmo_bytes((int64_t)TB, 3 * LA_G));
    w.wop3 = (__nv_bfloat16*)take(mbn_weight_operand3_bytes(d.in, LA_G));
    w.xproj = (float*)take(sizeof(float) * TB * 2 * LA_G);
  } else {
    w.px_hi = w.px_lo = w.py_hi = w.py_lo = w.p3 = w.wop3 = nullptr;
    w.xproj = nullptr;
    w.gates = w.cseq = w.scores = w.ctxs = w.hidd = w.logits = w.dlogits = w.dctx = w.hbar = w.dhid = w.du = w.dsum = w.red = nullptr;
    w.dhseq = w.dgates = w.dx = w.draw2 = w.dpool1 = w.draw1 = nullptr;
    w.bstats = w.loss_acc = nullptr;
  }
  w.bytes = off;
  return w;
}

// ---------------------------------------------------------------------------------------------
// Conv2d(CIN, 8, 3, padding=2) + bias; thread = one output pixel, all 8 channels.  train: per-channel sum / sum of squares.
// ---------------------------------------------------------------------------------------------
template <int CIN>
__global__ void __launch_bounds__(256) las_conv_kernel(const float* __restrict__ in, const float* __restrict__ w, const float* __restrict__ bias,
                                                       float* __restrict__ out, int64_t B, int hi, int wi, double* __restrict__ stats) {
  __shared__ float s_w[LA_C * CIN * 9];
  __shared__ float s_part[8][2 * LA_C];
  for (int i = threadIdx.x; i < LA_C * CIN * 9; i += blockDim.x) s_w[i] = w[i];
  __syncthreads();
  const int ho = hi + 2, wo = wi + 2;
  const int64_t n = B * ho * wo;
  float part[2 * LA_C];
#pragma unroll
  for (int i = 0; i < 2 * LA_C; ++i) part[i] = 0.f;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (int64_t)gridDim.x * blockDim.x) {
    const int x = (int)(idx % wo), y = (int)((idx / wo) % ho);
    const int64_t b = idx / ((int64_t)wo * ho);
    float acc[LA_C];
#pragma unroll
    for (int o = 0; o < LA_C; ++o) acc[o] = bias[o];
    for (int c = 0; c < CIN; ++c) {
#pragma unroll
      for (int ky = 0; ky < 3; ++ky) {
        const int yy = y + ky - 2;
        if (yy < 0 || yy >= hi) continue;
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          const int xx = x + kx - 2;
          if (xx < 0 || xx >= wi) continue;
          const float v = __ldg(in + ((b * CIN + c) * hi + yy) * (int64_t)wi + xx);
#pragma unroll
          for (int o = 0; o < LA_C; ++o) acc[o] = fmaf(s_w[(o * CIN + c) * 9 + ky * 3 + kx], v, acc[o]);
        }
      }
    }
#pragma unroll
    for (int o = 0; o < LA_C; ++o) {
      out[((b * LA_C + o) * ho + y) * (int64_t)wo + x] = acc[o];
      part[o] += acc[o];
      part[LA_C + o] = fmaf(acc[o], acc[o], part[LA_C + o]);
    }
  }
  if (stats) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < 2 * LA_C; ++i) {
      const float s = warp_sum(part[i]);
      if (lane == 0) s_part[warp][i] = s;
    }
    __syncthreads();
    if (threadIdx.x < 2 * LA_C) {
      double t = 0.0;
      for (int wv = 0; wv < 8; ++wv) t += (double)s_part[wv][threadIdx.x];
      atomicAdd(stats + threadIdx.x, t);
    }
  }
}

// statistics (train) or running statistics (eval) -> fused scale / shift (+ mean, rstd for the backward); train also updates the running statistics
__global__ void las_bn_finalize_kernel(const double* __restrict__ stats, double count, const float* __restrict__ gamma, const float* __restrict__ beta,
                                       float* __restrict__ run_mean, float* __restrict__ run_var, int64_t* __restrict__ nbt, int train,
                                       float* __restrict__ out) {
  const int c = threadIdx.x;
  if (c == 0 && train && nbt) *nbt += 1;
  if (c >= LA_C) return;
  double mean, var;
  if (train) {
    mean = stats[c] / count;
    var = stats[LA_C + c] / count - mean * mean;
    if (var < 0.0) var = 0.0;
    const double unbiased = count > 1.0 ? var * count / (count - 1.0) : var;
    run_mean[c] = (float)((1.0 - LA_MOM) * run_mean[c] + LA_MOM * mean);
    run_var[c] = (float)((1.0 - LA_MOM) * run_var[c] + LA_MOM * unbiased);
  } else {
    mean = run_mean[c];
    var = run_var[c];
  }
  const double r = 1.0 / sqrt(var + LA_EPS);
  out[c] = (float)(gamma[c] * r);
  out[LA_C + c] = (float)(beta[c] - mean * gamma[c] * r);
  out[2 * LA_C + c] = (float)mean;
  out[3 * LA_C + c] = (float)r;
}

// BatchNorm + ReLU + MaxPool(1, 2).  time_major = 0: out [B, 8, h, wp];  1: out [wp, B, 8 * h] (the LSTM input, x.permute(3,0,1,2).view)
__global__ void __launch_bounds__(256) las_bn_relu_pool_kernel(const float* __restrict__ raw, const float* __restrict__ bn, float* __restrict__ out,
                                                               int64_t B, int h, int w, int wp, int time_major) {
  const int64_t n = B * LA_C * h * wp;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (int64_t)gridDim.x * blockDim.x) {
    const int xp = (int)(idx % wp), y = (int)((idx / wp) % h), c = (int)((idx / ((int64_t)wp * h)) % LA_C);
    const int64_t b = idx / ((int64_t)wp * h * LA_C);
    const float* r = raw + ((b * LA_C + c) * h + y) * (int64_t)w + 2 * xp;
    const float sc = bn[c], sh = bn[LA_C + c];
    const float v = fmaxf(fmaxf(fmaf(r[0], sc, sh), 0.f), fmaxf(fmaf(r[1], sc, sh), 0.f));
    if (time_major) out[((int64_t)xp * B + b) * (LA_C * h) + c * h + y] = v;
    else out[idx] = v;
  }
}

// [W_ih | W_hh] of both directions transposed to [k][gate row] + summed biases
__global__ void las_lstm_prep_kernel(LasParams q, int in, float* __restrict__ wt, float* __restrict__ bsum) {
  const int K = in + LA_H;
  const int64_t n = 2 * (int64_t)K * LA_G;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int j = (int)(i % LA_G), k = (int)((i / LA_G) % K), d = (int)(i / ((int64_t)LA_G * K));
    wt[i] = k < in ? q.wih[d][(size_t)j * in + k] : q.whh[d][(size_t)j * LA_H + (k - in)];
  }
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < 2 * LA_G) bsum[t] = q.bih[t / LA_G][t % LA_G] + q.bhh[t / LA_G][t % LA_G];
}

// one direction of the recurrence for LA_NB sequences: grid = (ceil(B / LA_NB), 2), block = 384 threads (thread j = gate row j)
__global__ void __launch_bounds__(LA_G, 1) las_lstm_kernel(const float* __restrict__ x, const int64_t* __restrict__ lengths, const float* __restrict__ wt,
                                                          const float* __restrict__ bsum, float* __restrict__ hseq, int64_t B, int T, int in,
                                                          float* __restrict__ gsave, float* __restrict__ csave,
                                                          const float* __restrict__ xproj) {
  // xproj != null: [T, B, 2, 384] = W_ih x_t + b_ih + b_hh of every step, computed up front as one tensor-core GEMM per direction; the loop
  // then only carries the 96-wide recurrent product
  extern __shared__ __align__(16) float smem[];
  const int K = in + LA_H, dir = blockIdx.y, j = threadIdx.x;
  float* s_xh = smem;                         // [K][LA_NB]
  float* s_g = s_xh + (size_t)K * LA_NB;      // [LA_NB][LA_G] activated gates
  float* s_c = s_g + LA_NB * LA_G;            // [LA_NB][LA_H]
  __shared__ int s_len[LA_NB];
  const int64_t b0 = (int64_t)blockIdx.x * LA_NB;
  if (j < LA_NB) s_len[j] = (b0 + j < B) ? (int)(lengths[b0 + j] < (int64_t)T ? lengths[b0 + j] : (int64_t)T) : 0;
  for (int i = j; i < LA_NB * LA_H; i += LA_G) s_c[i] = 0.f;
  for (int i = j; i < LA_H * LA_NB; i += LA_G) s_xh[(size_t)in * LA_NB + i] = 0.f;      // h_{-1} = 0
  __syncthreads();
  const float* w = wt + (size_t)dir * K * LA_G;
  const float bias = bsum[dir * LA_G + j];
  for (int step = 0; step < T; ++step) {
    const int t = dir ? T - 1 - step : step;
    float acc[LA_NB];
    if (xproj) {
#pragma unroll
      for (int bb = 0; bb < LA_NB; ++bb)
        acc[bb] = (b0 + bb < B) ? __ldg(xproj + (((int64_t)t * B + b0 + bb) * 2 + dir) * LA_G + j) : 0.f;
    } else {
      // x_t of the 16 sequences -> s_xh[k][b]
      for (int i = j; i < in * LA_NB; i += LA_G) {
        const int bb = i / in, k = i - bb * in;
        s_xh[(size_t)k * LA_NB + bb] = (b0 + bb < B) ? __ldg(x + ((int64_t)t * B + b0 + bb) * in + k) : 0.f;
      }
      __syncthreads();
#pragma unroll
      for (int bb = 0; bb < LA_NB; ++bb) acc[bb] = bias;
    }
    for (int k = xproj ? in : 0; k < K; ++k) {
      const float wv = __ldg(w + (size_t)k * LA_G + j);
      const float4* xv = reinterpret_cast<const float4*>(s_xh + (size_t)k * LA_NB);
#pragma unroll
      for (int q4 = 0; q4 < LA_NB / 4; ++q4) {
        const float4 v = xv[q4];
        acc[4 * q4] = fmaf(wv, v.x, acc[4 * q4]);
        acc[4 * q4 + 1] = fmaf(wv, v.y, acc[4 * q4 + 1]);
        acc[4 * q4 + 2] = fmaf(wv, v.z, acc[4 * q4 + 2]);
        acc[4 * q4 + 3] = fmaf(wv, v.w, acc[4 * q4 + 3]);
      }
    }
    const bool tanh_gate = (j >= 2 * LA_H) && (j < 3 * LA_H);
#pragma unroll
    for (int bb = 0; bb < LA_NB; ++bb) {
      const float a = tanh_gate ? tanhf(acc[bb]) : 1.f / (1.f + expf(-acc[bb]));
      s_g[bb * LA_G + j] = a;
      if (gsave && b0 + bb < B) gsave[(((int64_t)t * B + b0 + bb) * 2 + dir) * LA_G + j] = a;
    }
    __syncthreads();
    // cell update on (sequence, unit) pairs; a sequence only advances while t < its length (the reverse direction starts there)
    for (int i = j; i < LA_NB * LA_H; i += LA_G) {
      const int bb = i / LA_H, u = i - bb * LA_H;
      const bool live = t < s_len[bb];
      float hval = 0.f;
      if (live) {
        const float* g = s_g + bb * LA_G;
        const float c = g[LA_H + u] * s_c[i] + g[u] * g[2 * LA_H + u];
        s_c[i] = c;
        if (csave) csave[(((int64_t)t * B + b0 + bb) * 2 + dir) * LA_H + u] = c;
        hval = g[3 * LA_H + u] * tanhf(c);
        s_xh[(size_t)(in + u) * LA_NB + bb] = hval;
      }
      if (b0 + bb < B) hseq[((int64_t)t * B + b0 + bb) * LA_D + dir * LA_H + u] = hval;
    }
    __syncthreads();
  }
}

// u[h][k] = sum_l context_vec[l * 4 + h] * v_proj.weight[h * 48 + l][k], u0[h] = sum_l context_vec[l * 4 + h] * v_proj.bias[h * 48 + l]
// (stored after u): the attention logit of head h at step t is u[h] . h_t + u0[h]
__global__ void las_attn_prep_kernel(LasParams q, float* __restrict__ u) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= LA_HEADS * LA_D + LA_HEADS) return;
  constexpr int HD = LA_D / LA_HEADS;
  float acc = 0.f;
  if (i < LA_HEADS * LA_D) {
    const int h = i / LA_D, k = i - h * LA_D;
    for (int l = 0; l < HD; ++l) acc = fmaf(q.cvec[l * LA_HEADS + h], q.vw[(h * HD + l) * LA_D + k], acc);
  } else {
    const int h = i - LA_HEADS * LA_D;
    for (int l = 0; l < HD; ++l) acc = fmaf(q.cvec[l * LA_HEADS + h], q.vb[h * HD + l], acc);
  }
  u[i] = acc;
}

// Dropout(p) of the fc hidden layer: counter-based hash of (seed, utterance, unit)
__device__ __forceinline__ bool las_keep(unsigned long long seed, int64_t b, int r, float p) {
  if (p <= 0.f) return true;
  unsigned long long z = seed + 0x9E3779B97F4A7C15ull * (unsigned long long)(b * LA_DNN + r + 1);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z ^= z >> 31;
  return (float)(z >> 40) * (1.f / 16777216.f) >= p;
}

// attention + MLP head: one CTA per utterance.  train: scores, context and the dropped hidden layer are kept for the backward
__global__ void __launch_bounds__(256) las_head_kernel(const float* __restrict__ hseq, const int64_t* __restrict__ lengths, LasParams q, int64_t B,
                                                       int T, int L, float* __restrict__ logits, float drop_p, unsigned long long seed,
                                                       const float* __restrict__ u, float* __restrict__ sc_save, float* __restrict__ ctx_save,
                                                       float* __restrict__ hid_save, float* __restrict__ logits_save) {
  extern __shared__ __align__(16) float smem[];
  float* s_h = smem;                       // [T][192]
  float* s_sc = s_h + (size_t)T * LA_D;    // [T][4] attention logits -> scores
  float* s_ctx = s_sc + T * LA_HEADS;      // [192]
  float* s_hid = s_ctx + LA_D;             // [256]
  float* s_hbar = s_hid + LA_DNN;          // [4][192]
  const int64_t b = blockIdx.x;
  const int tid = threadIdx.x, len = (int)(lengths[b] < (int64_t)T ? lengths[b] : (int64_t)T);
  for (int i = tid; i < T * LA_D; i += 256) s_h[i] = hseq[((int64_t)(i / LA_D) * B + b) * LA_D + (i % LA_D)];
  __syncthreads();
  // attention logits: logits[t][h] = sum_l values[t][h * 48 + l] * context_vec.view(48, 4)[l][h], values = v_proj(h_t)
  //                                = u[h] . h_t + u0[h]   (u, u0 from las_attn_prep_kernel: the projection folded into the context vector)
  for (int i = tid; i < T * LA_HEADS; i += 256) {
    const int t = i / LA_HEADS, h = i - t * LA_HEADS;
    float acc = u[LA_HEADS * LA_D + h];
    for (int k = 0; k < LA_D; ++k) acc = fmaf(__ldg(u + h * LA_D + k), s_h[t * LA_D + k], acc);
    s_sc[i] = acc + (t < len ? 0.f : -100.f);
  }
  __syncthreads();
  if (tid < LA_HEADS) {       // softmax over time
    float mx = -INFINITY;
    for (int t = 0; t < T; ++t) mx = fmaxf(mx, s_sc[t * LA_HEADS + tid]);
    float se = 0.f;
    for (int t = 0; t < T; ++t) {
      const float e = expf(s_sc[t * LA_HEADS + tid] - mx);
      s_sc[t * LA_HEADS + tid] = e;
      se += e;
    }
    for (int t = 0; t < T; ++t) s_sc[t * LA_HEADS + tid] /= se;
  }
  __syncthreads();
  if (sc_save)
    for (int i = tid; i < T * LA_HEADS; i += 256) sc_save[b * T * LA_HEADS + i] = s_sc[i];
  // context[h * 48 + l] = sum_t scores[t][h] * keys[t][h * 48 + l] = k_proj( sum_t scores[t][h] h_t )[row] (+ bias: the scores sum to 1)
  for (int i = tid; i < LA_HEADS * LA_D; i += 256) {
    const int h = i / LA_D, k = i - h * LA_D;
    float hb = 0.f;
    for (int t = 0; t < T; ++t) hb = fmaf(s_sc[t * LA_HEADS + h], s_h[t * LA_D + k], hb);
    s_hbar[i] = hb;
  }
  __syncthreads();
  for (int row = tid; row < LA_D; row += 256) {
    const float* hb = s_hbar + (row / (LA_D / LA_HEADS)) * LA_D;
    float acc = q.kb[row];
    for (int k = 0; k < LA_D; ++k) acc = fmaf(__ldg(q.kw + row * LA_D + k), hb[k], acc);
    s_ctx[row] = acc;
    if (ctx_save) ctx_save[b * LA_D + row] = acc;
  }
  __syncthreads();
  for (int r = tid; r < LA_DNN; r += 256) {
    float acc = q.f0b[r];
    for (int k = 0; k < LA_D; ++k) acc = fmaf(__ldg(q.f0w + r * LA_D + k), s_ctx[k], acc);
    acc = fmaxf(acc, 0.f);
    if (drop_p > 0.f) acc = las_keep(seed, b, r, drop_p) ? acc / (1.f - drop_p) : 0.f;
    s_hid[r] = acc;
    if (hid_save) hid_save[b * LA_DNN + r] = acc;
  }
  __syncthreads();
  for (int l = tid; l < L; l += 256) {
    float acc = q.f3b[l];
    for (int k = 0; k < LA_DNN; ++k) acc = fmaf(__ldg(q.f3w + (size_t)l * LA_DNN + k), s_hid[k], acc);
    logits[b * L + l] = acc;
    if (logits_save) logits_save[b * L + l] = acc;
  }
}

// =============================================================================================
// backward
// =============================================================================================
// softmax cross-entropy: dlogits = (p - onehot) / batch (or the caller's dlogits), loss
__global__ void las_ce_kernel(const float* __restrict__ logits, const int64_t* __restrict__ labels, const float* __restrict__ dlogits_in,
                              float* __restrict__ dlogits, double* __restrict__ loss_acc, int64_t B, int L, float inv_batch) {
  const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  double nll = 0.0;
  if (b < B) {
    if (!labels) {
      for (int l = 0; l < L; ++l) dlogits[b * L + l] = dlogits_in[b * L + l];
    } else {
      const float* z = logits + b * L;
      float mx = z[0];
      for (int l = 1; l < L; ++l) mx = fmaxf(mx, z[l]);
      float se = 0.f;
      for (int l = 0; l < L; ++l) se += expf(z[l] - mx);
      const float lse = mx + logf(se);
      const int64_t y = labels[b];
      for (int l = 0; l < L; ++l) dlogits[b * L + l] = (expf(z[l] - lse) - (l == y ? 1.f : 0.f)) * inv_batch;
      if (y >= 0 && y < L) nll = (double)(lse - z[y]);
    }
  }
  nll = warp_sum(nll);
  if ((threadIdx.x & 31) == 0 && nll != 0.0) atomicAdd(loss_acc, nll * (double)inv_batch);
}
__global__ void las_loss_out_kernel(const double* __restrict__ loss_acc, float* __restrict__ loss) { *loss = (float)*loss_acc; }

// head backward, one CTA per utterance: dlogits -> fc -> attention -> dh_t, plus the per-utterance factors of the weight gradients
struct LasHeadBwd {
  const float *hseq, *scores, *hidd, *dlogits, *u;
  float *dctx, *hbar, *dhid, *du, *dsum, *dhseq;
  int64_t B;
  int T, L;
  float drop_p;
};
__global__ void __launch_bounds__(256) las_head_bwd_kernel(const LasHeadBwd a, LasParams q) {
  extern __shared__ __align__(16) float smem[];
  const int T = a.T, L = a.L, tid = threadIdx.x;
  constexpr int HD = LA_D / LA_HEADS;
  float* s_h = smem;                          // [T][192]
  float* s_sc = s_h + (size_t)T * LA_D;       // [T][4] scores
  float* s_dl = s_sc + T * LA_HEADS;          // [T][4] dsc -> dlogit
  float* s_dhid = s_dl + T * LA_HEADS;        // [256]
  float* s_dctx = s_dhid + LA_DNN;            // [192]
  float* s_dhbar = s_dctx + LA_D;             // [4][192]
  float* s_u = s_dhbar + LA_HEADS * LA_D;     // [4][192]
  float* s_dlog = s_u + LA_HEADS * LA_D;      // [L]
  const int64_t b = blockIdx.x, B = a.B;
  for (int i = tid; i < T * LA_D; i += 256) s_h[i] = a.hseq[((int64_t)(i / LA_D) * B + b) * LA_D + (i % LA_D)];
  for (int i = tid; i < T * LA_HEADS; i += 256) s_sc[i] = a.scores[b * T * LA_HEADS + i];
  for (int i = tid; i < LA_HEADS * LA_D; i += 256) s_u[i] = a.u[i];
  for (int i = tid; i < L; i += 256) s_dlog[i] = a.dlogits[b * L + i];
  __syncthreads();
  // fc.3 and the ReLU / dropout mask (hidd > 0 <=> unit active and kept)
  for (int r = tid; r < LA_DNN; r += 256) {
    float acc = 0.f;
    for (int l = 0; l < L; ++l) acc = fmaf(s_dlog[l], __ldg(q.f3w + (size_t)l * LA_DNN + r), acc);
    const float d = a.hidd[b * LA_DNN + r] > 0.f ? acc / (1.f - a.drop_p) : 0.f;
    s_dhid[r] = d;
    a.dhid[b * LA_DNN + r] = d;
  }
  // hbar[h][k] = sum_t scores[t][h] h_t[k]
  for (int i = tid; i < LA_HEADS * LA_D; i += 256) {
    const int h = i / LA_D, k = i - h * LA_D;
    float acc = 0.f;
    for (int t = 0; t < T; ++t) acc = fmaf(s_sc[t * LA_HEADS + h], s_h[t * LA_D + k], acc);
    a.hbar[b * LA_HEADS * LA_D + i] = acc;
  }
  __syncthreads();
  for (int k = tid; k < LA_D; k += 256) {      // fc.0
    float acc = 0.f;
    for (int r = 0; r < LA_DNN; ++r) acc = fmaf(s_dhid[r], __ldg(q.f0w + r * LA_D + k), acc);
    s_dctx[k] = acc;
    a.dctx[b * LA_D + k] = acc;
  }
  __syncthreads();
  // context[row] = k_proj.bias[row] + sum_k k_proj.weight[row][k] hbar[head(row)][k]
  for (int i = tid; i < LA_HEADS * LA_D; i += 256) {
    const int h = i / LA_D, k = i - h * LA_D;
    float acc = 0.f;
    for (int l = 0; l < HD; ++l) acc = fmaf(s_dctx[h * HD + l], __ldg(q.kw + (h * HD + l) * LA_D + k), acc);
    s_dhbar[i] = acc;
  }
  __syncthreads();
  for (int i = tid; i < T * LA_HEADS; i += 256) {       // d scores
    const int t = i / LA_HEADS, h = i - t * LA_HEADS;
    float acc = 0.f;
    for (int k = 0; k < LA_D; ++k) acc = fmaf(s_dhbar[h * LA_D + k], s_h[t * LA_D + k], acc);
    s_dl[i] = acc;
  }
  __syncthreads();
  if (tid < LA_HEADS) {                                  // softmax over time
    float dot = 0.f;
    for (int t = 0; t < T; ++t) dot = fmaf(s_sc[t * LA_HEADS + tid], s_dl[t * LA_HEADS + tid], dot);
    float sum = 0.f;
    for (int t = 0; t < T; ++t) {
      const float d = s_sc[t * LA_HEADS + tid] * (s_dl[t * LA_HEADS + tid] - dot);
      s_dl[t * LA_HEADS + tid] = d;
      sum += d;
    }
    a.dsum[b * LA_HEADS + tid] = sum;
  }
  __syncthreads();
  for (int i = tid; i < LA_HEADS * LA_D; i += 256) {    // du[h][k] = sum_t dlogit[t][h] h_t[k]
    const int h = i / LA_D, k = i - h * LA_D;
    float acc = 0.f;
    for (int t = 0; t < T; ++t) acc = fmaf(s_dl[t * LA_HEADS + h], s_h[t * LA_D + k], acc);
    a.du[b * LA_HEADS * LA_D + i] = acc;
  }
  for (int i = tid; i < T * LA_D; i += 256) {           // dh_t
    const int t = i / LA_D, k = i - t * LA_D;
    float acc = 0.f;
#pragma unroll
    for (int h = 0; h < LA_HEADS; ++h)
      acc += s_sc[t * LA_HEADS + h] * s_dhbar[h * LA_D + k] + s_dl[t * LA_HEADS + h] * s_u[h * LA_D + k];
    a.dhseq[((int64_t)t * B + b) * LA_D + k] = acc;
  }
}

// v_proj / context_vec gradients from the batch sums DU [4][192], DS [4] (red = [DU | DS])
__global__ void las_attn_grads_kernel(LasParams q, const float* __restrict__ red, float* __restrict__ d_vw, float* __restrict__ d_vb,
                                      float* __restrict__ d_cvec) {
  constexpr int HD = LA_D / LA_HEADS;
  const float* DU = red;
  const float* DS = red + LA_HEADS * LA_D;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < LA_D * LA_D) {
    const int row = i / LA_D, k = i - row * LA_D, h = row / HD, l = row - h * HD;
    d_vw[i] = q.cvec[l * LA_HEADS + h] * DU[h * LA_D + k];
  }
  if (i < LA_D) {
    const int h = i / HD, l = i - h * HD;          // i = v_proj row
    d_vb[i] = q.cvec[l * LA_HEADS + h] * DS[h];
    float acc = q.vb[i] * DS[h];
    for (int k = 0; k < LA_D; ++k) acc = fmaf(DU[h * LA_D + k], q.vw[i * LA_D + k], acc);
    d_cvec[l * LA_HEADS + h] = acc;
  }
}

// C[m][n] (row stride ldc) += sum_k A(m, k) B(k, n) with element strides (sa_m, sa_k), (sb_k, sb_n); gridDim.z splits K; atomic accumulate
#define LG_BM 64
#define LG_BN 64
#define LG_BK 16
__global__ void __launch_bounds__(256) las_gemm_kernel(const float* __restrict__ A, int64_t sa_m, int64_t sa_k, const float* __restrict__ Bm,
                                                       int64_t sb_k, int64_t sb_n, float* __restrict__ C, int64_t ldc, int64_t M, int N, int64_t K,
                                                       int atomic) {
  __shared__ float s_a[LG_BK][LG_BM + 4];
  __shared__ float s_b[LG_BK][LG_BN + 4];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int64_t m0 = (int64_t)blockIdx.x * LG_BM;
  const int n0 = blockIdx.y * LG_BN;
  const int64_t kchunk = (K + gridDim.z - 1) / gridDim.z;
  const int64_t k_lo = (int64_t)blockIdx.z * kchunk, k_hi = k_lo + kchunk < K ? k_lo + kchunk : K;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  const bool a_kfast = sa_k == 1, b_kfast = sb_k == 1;
  for (int64_t k0 = k_lo; k0 < k_hi; k0 += LG_BK) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int idx = tid + 256 * e;
      const int am = a_kfast ? idx / LG_BK : idx % LG_BM, ak = a_kfast ? idx % LG_BK : idx / LG_BM;
      s_a[ak][am] = (m0 + am < M && k0 + ak < k_hi) ? __ldg(A + (m0 + am) * sa_m + (k0 + ak) * sa_k) : 0.f;
      const int bn = b_kfast ? idx / LG_BK : idx % LG_BN, bk = b_kfast ? idx % LG_BK : idx / LG_BN;
      s_b[bk][bn] = (n0 + bn < N && k0 + bk < k_hi) ? __ldg(Bm + (k0 + bk) * sb_k + (int64_t)(n0 + bn) * sb_n) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < LG_BK; ++kk) {
      const float4 av = *reinterpret_cast<const float4*>(&s_a[kk][ty * 4]);
      const float4 bv = *reinterpret_cast<const float4*>(&s_b[kk][tx * 4]);
      const float ar[4] = {av.x, av.y, av.z, av.w}, br[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(ar[i], br[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int64_t m = m0 + ty * 4 + i;
      const int n = n0 + tx * 4 + j;
      if (m < M && n < N) {
        if (atomic) atomicAdd(C + m * ldc + n, acc[i][j]);
        else C[m * ldc + n] = acc[i][j];
      }
    }
}

// out[c] += sum_r A[r * ld + c]  (column sums over `rows` rows): grid.x = column blocks of 256, grid.y splits the rows
__global__ void __launch_bounds__(256) las_colsum_kernel(const float* __restrict__ A, int64_t rows, int cols, int64_t ld, float* __restrict__ out,
                                                         float* __restrict__ out2) {
  const int c = blockIdx.x * 256 + threadIdx.x;
  if (c >= cols) return;
  const int64_t per = (rows + gridDim.y - 1) / gridDim.y;
  const int64_t lo = (int64_t)blockIdx.y * per, hi = lo + per < rows ? lo + per : rows;
  float acc = 0.f;
  for (int64_t r = lo; r < hi; ++r) acc += A[r * ld + c];
  atomicAdd(out + c, acc);
  if (out2) atomicAdd(out2 + c, acc);
}

// BPTT of one direction for LA_NB sequences: grid = (ceil(B / LA_NB), 2), block = 384.  Writes the gate pre-activation gradients da_t
// (zero on padded steps); only dh_{t-1} = da_t W_hh stays in the loop.
__global__ void __launch_bounds__(LA_G, 1) las_lstm_bwd_kernel(const float* __restrict__ gates, const float* __restrict__ cseq,
                                                              const float* __restrict__ dhseq, const int64_t* __restrict__ lengths,
                                                              const float* __restrict__ whh_f, const float* __restrict__ whh_r,
                                                              float* __restrict__ dgates, int64_t B, int T) {
  __shared__ float s_da[LA_NB][LA_G];
  __shared__ float s_dh[LA_NB][LA_H];
  __shared__ float s_dc[LA_NB][LA_H];
  __shared__ int s_len[LA_NB];
  const int dir = blockIdx.y, j = threadIdx.x;
  const int64_t b0 = (int64_t)blockIdx.x * LA_NB;
  const float* whh = dir ? whh_r : whh_f;
  if (j < LA_NB) s_len[j] = (b0 + j < B) ? (int)(lengths[b0 + j] < (int64_t)T ? lengths[b0 + j] : (int64_t)T) : 0;
  for (int i = j; i < LA_NB * LA_H; i += LA_G) (&s_dh[0][0])[i] = (&s_dc[0][0])[i] = 0.f;
  __syncthreads();
  for (int step = 0; step < T; ++step) {
    const int t = dir ? step : T - 1 - step;             // reverse of the forward order
    const int tprev = dir ? t + 1 : t - 1;               // the step whose state this one consumed
    for (int i = j; i < LA_NB * LA_H; i += LA_G) {
      const int bb = i / LA_H, u = i - bb * LA_H;
      const bool live = t < s_len[bb];
      float da_i = 0.f, da_f = 0.f, da_g = 0.f, da_o = 0.f;
      const int64_t row = ((int64_t)t * B + b0 + bb) * 2 + dir;
      if (live) {
        const float* g = gates + row * LA_G;
        const float gi = g[u], gf = g[LA_H + u], gg = g[2 * LA_H + u], go = g[3 * LA_H + u];
        const float c = cseq[row * LA_H + u];
        const bool has_prev = tprev >= 0 && tprev < s_len[bb];
        const float cp = has_prev ? cseq[(((int64_t)tprev * B + b0 + bb) * 2 + dir) * LA_H + u] : 0.f;
        const float dh = dhseq[((int64_t)t * B + b0 + bb) * LA_D + dir * LA_H + u] + s_dh[bb][u];
        const float tc = tanhf(c);
        const float dc = s_dc[bb][u] + dh * go * (1.f - tc * tc);
        s_dc[bb][u] = dc * gf;
        da_i = dc * gg * gi * (1.f - gi);
        da_f = dc * cp * gf * (1.f - gf);
        da_g = dc * gi * (1.f - gg * gg);
        da_o = dh * tc * go * (1.f - go);
      }
      s_da[bb][u] = da_i; s_da[bb][LA_H + u] = da_f; s_da[bb][2 * LA_H + u] = da_g; s_da[bb][3 * LA_H + u] = da_o;
      if (b0 + bb < B) {
        float* dg = dgates + row * LA_G;
        dg[u] = da_i; dg[LA_H + u] = da_f; dg[2 * LA_H + u] = da_g; dg[3 * LA_H + u] = da_o;
      }
    }
    __syncthreads();
    {   // dh_{prev}[bb][u] = sum_j da[bb][j] W_hh[j][u]: thread -> unit u, four sequences
      const int u = j % LA_H, qd = j / LA_H;
      float acc[4] = {0.f, 0.f, 0.f, 0.f};
      for (int k = 0; k < LA_G; ++k) {
        const float w = __ldg(whh + (size_t)k * LA_H + u);
#pragma unroll
        for (int s = 0; s < 4; ++s) acc[s] = fmaf(s_da[qd * 4 + s][k], w, acc[s]);
      }
#pragma unroll
      for (int s = 0; s < 4; ++s) s_dh[qd * 4 + s][u] = acc[s];
    }
    __syncthreads();
  }
}

// BatchNorm + ReLU + MaxPool(1, 2) backward, pass 1: per-channel sum dy, sum dy * xhat (dy = gradient at the BatchNorm output).
// g: gradient of the pooled tensor, [B, 8, h, wp] or time major [wp, B, 8 * h].  grid = (blocks, 8 channels)
__device__ __forceinline__ float las_pool_grad(const float* __restrict__ r, float sc, float sh, float g, int which) {
  const float v0 = fmaf(r[0], sc, sh), v1 = fmaf(r[1], sc, sh);
  const int sel = v1 > v0 ? 1 : 0;                       // the first maximum wins a tie (torch max_pool2d)
  return (sel == which && (sel ? v1 : v0) > 0.f) ? g : 0.f;
}
__global__ void __launch_bounds__(256) las_bn_bwd_stats_kernel(const float* __restrict__ raw, const float* __restrict__ bn, const float* __restrict__ g,
                                                               int64_t B, int h, int w, int wp, int time_major, double* __restrict__ bstats) {
  __shared__ float s_part[8][2];
  const int c = blockIdx.y;
  const float sc = bn[c], sh = bn[LA_C + c], mean = bn[2 * LA_C + c], rstd = bn[3 * LA_C + c];
  const int64_t n = B * h * wp;
  float s0 = 0.f, s1 = 0.f;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (int64_t)gridDim.x * blockDim.x) {
    const int xp = (int)(idx % wp), y = (int)((idx / wp) % h);
    const int64_t b = idx / ((int64_t)wp * h);
    const float* r = raw + ((b * LA_C + c) * h + y) * (int64_t)w + 2 * xp;
    const float gv = time_major ? g[((int64_t)xp * B + b) * (LA_C * h) + c * h + y] : g[((b * LA_C + c) * h + y) * (int64_t)wp + xp];
    const float d0 = las_pool_grad(r, sc, sh, gv, 0), d1 = las_pool_grad(r, sc, sh, gv, 1);
    s0 += d0 + d1;
    s1 += d0 * (r[0] - mean) * rstd + d1 * (r[1] - mean) * rstd;
  }
  s0 = warp_sum(s0); s1 = warp_sum(s1);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) { s_part[warp][0] = s0; s_part[warp][1] = s1; }
  __syncthreads();
  if (threadIdx.x < 2) {
    double t = 0.0;
    for (int wv = 0; wv < 8; ++wv) t += (double)s_part[wv][threadIdx.x];
    atomicAdd(bstats + threadIdx.x * LA_C + c, t);
  }
}
// pass 2: d raw = gamma * rstd * (dy - mean(dy) - xhat * mean(dy * xhat)) over the whole raw tensor; block 0 writes dgamma, dbeta
__global__ void __launch_bounds__(256) las_bn_bwd_apply_kernel(const float* __restrict__ raw, const float* __restrict__ bn, const float* __restrict__ g,
                                                               const float* __restrict__ gamma, int64_t B, int h, int w, int wp, int time_major,
                                                               const double* __restrict__ bstats, double count, float* __restrict__ draw,
                                                               float* __restrict__ dgamma, float* __restrict__ dbeta) {
  if (blockIdx.x == 0 && threadIdx.x < LA_C) {
    dbeta[threadIdx.x] = (float)bstats[threadIdx.x];
    dgamma[threadIdx.x] = (float)bstats[LA_C + threadIdx.x];
  }
  const int64_t n = B * LA_C * h * w;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (int64_t)gridDim.x * blockDim.x) {
    const int x = (int)(idx % w), y = (int)((idx / w) % h), c = (int)((idx / ((int64_t)w * h)) % LA_C);
    const int64_t b = idx / ((int64_t)w * h * LA_C);
    const float sc = bn[c], sh = bn[LA_C + c], mean = bn[2 * LA_C + c], rstd = bn[3 * LA_C + c];
    const int xp = x >> 1;
    float dy = 0.f;
    if (xp < wp) {
      const float* r = raw + ((b * LA_C + c) * h + y) * (int64_t)w + 2 * xp;
      const float gv = time_major ? g[((int64_t)xp * B + b) * (LA_C * h) + c * h + y] : g[((b * LA_C + c) * h + y) * (int64_t)wp + xp];
      dy = las_pool_grad(r, sc, sh, gv, x & 1);
    }
    const float xhat = (raw[idx] - mean) * rstd;
    const float m0 = (float)(bstats[c] / count), m1 = (float)(bstats[LA_C + c] / count);
    draw[idx] = gamma[c] * rstd * (dy - m0 - xhat * m1);
  }
}

// conv2 data gradient: din[b, c, y, x] = sum_{o, ky, kx} dout[b, o, y + 2 - ky, x + 2 - kx] w[o, c, ky, kx]  (padding 2: always in range)
__global__ void __launch_bounds__(256) las_conv_dgrad_kernel(const float* __restrict__ dout, const float* __restrict__ w, float* __restrict__ din,
                                                             int64_t B, int hi, int wi) {
  __shared__ float s_w[LA_C * LA_C * 9];
  for (int i = threadIdx.x; i < LA_C * LA_C * 9; i += blockDim.x) s_w[i] = w[i];
  __syncthreads();
  const int ho = hi + 2, wo = wi + 2;
  const int64_t n = B * hi * wi;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (int64_t)gridDim.x * blockDim.x) {
    const int x = (int)(idx % wi), y = (int)((idx / wi) % hi);
    const int64_t b = idx / ((int64_t)wi * hi);
    float acc[LA_C];
#pragma unroll
    for (int c = 0; c < LA_C; ++c) acc[c] = 0.f;
    for (int o = 0; o < LA_C; ++o) {
#pragma unroll
      for (int ky = 0; ky < 3; ++ky) {
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          const float v = __ldg(dout + ((b * LA_C + o) * ho + (y + 2 - ky)) * (int64_t)wo + (x + 2 - kx));
#pragma unroll
          for (int c = 0; c < LA_C; ++c) acc[c] = fmaf(s_w[(o * LA_C + c) * 9 + ky * 3 + kx], v, acc[c]);
        }
      }
    }
#pragma unroll
    for (int c = 0; c < LA_C; ++c) din[((b * LA_C + c) * hi + y) * (int64_t)wi + x] = acc[c];
  }
}

// conv weight / bias gradients: block = LW_ROWS x (8 * CIN * 3) threads, thread = (staged row, o, c, ky) with the three kx taps in
// registers over a sliding input window (2 shared-memory loads per 3 FMAs); LW_ROWS output rows (b, yo) are staged per iteration, input
// rows with two zero columns on each side so the taps need no bounds test
#define LW_ROWS 4
template <int CIN>
__global__ void __launch_bounds__(LW_ROWS * LA_C * CIN * 3) las_conv_wgrad_kernel(const float* __restrict__ in, const float* __restrict__ dout,
                                                                                   int64_t B, int hi, int wi, float* __restrict__ dw,
                                                                                   float* __restrict__ dbias) {
  extern __shared__ __align__(16) float smem[];
  const int ho = hi + 2, wo = wi + 2, wpad = wi + 4;
  float* s_d = smem;                               // [LW_ROWS][8][wo]
  float* s_in = s_d + LW_ROWS * LA_C * wo;         // [LW_ROWS][CIN][3][wi + 4]
  constexpr int ROLES = LA_C * CIN * 3;
  const int tid = threadIdx.x, nth = LW_ROWS * ROLES;
  const int r = tid / ROLES, role = tid % ROLES;
  const int o = role / (CIN * 3), c = (role / 3) % CIN, ky = role % 3;
  const int64_t rows = B * ho;
  float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, accb = 0.f;
  for (int64_t row0 = (int64_t)blockIdx.x * LW_ROWS; row0 < rows; row0 += (int64_t)gridDim.x * LW_ROWS) {
    __syncthreads();
    // staging by row segments: a warp copies one contiguous run (a channel row of the gradient, or one padded input row) per turn, lanes
    // along x -- no per-element index arithmetic
    for (int seg = tid >> 5; seg < LW_ROWS * (LA_C + CIN * 3); seg += nth >> 5) {
      const int rr = seg / (LA_C + CIN * 3), sidx = seg - rr * (LA_C + CIN * 3);
      const int64_t row = row0 + rr;
      const bool ok = row < rows;
      const int yo = ok ? (int)(row % ho) : 0;
      const int64_t b = ok ? row / ho : 0;
      const int lane = tid & 31;
      if (sidx < LA_C) {
        const float* src = dout + ((b * LA_C + sidx) * ho + yo) * (int64_t)wo;
        float* dst = s_d + (rr * LA_C + sidx) * wo;
        for (int x = lane; x < wo; x += 32) dst[x] = ok ? src[x] : 0.f;
      } else {
        const int cc = (sidx - LA_C) / 3, kr = (sidx - LA_C) - cc * 3;
        const int yy = yo + kr - 2;
        const bool valid = ok && yy >= 0 && yy < hi;
        const float* src = in + ((b * CIN + cc) * hi + (valid ? yy : 0)) * (int64_t)wi;
        float* dst = s_in + ((rr * CIN + cc) * 3 + kr) * wpad;
        for (int x = lane; x < wpad; x += 32) dst[x] = (valid && x >= 2 && x < wi + 2) ? src[x - 2] : 0.f;
      }
    }
    __syncthreads();
    const float* dr = s_d + (r * LA_C + o) * wo;
    const float* ir = s_in + ((r * CIN + c) * 3 + ky) * wpad;        // input column xo + kx - 2 -> padded index xo + kx
    float w0 = ir[0], w1 = ir[1];
    for (int xo = 0; xo < wo; ++xo) {
      const float dv = dr[xo], w2 = ir[xo + 2];
      acc0 = fmaf(dv, w0, acc0);
      acc1 = fmaf(dv, w1, acc1);
      acc2 = fmaf(dv, w2, acc2);
      w0 = w1;
      w1 = w2;
      if (role % (CIN * 3) == 0) accb += dv;
    }
  }
  float* out = dw + ((o * CIN + c) * 3 + ky) * 3;
  atomicAdd(out, acc0);
  atomicAdd(out + 1, acc1);
  atomicAdd(out + 2, acc2);
  if (role % (CIN * 3) == 0) atomicAdd(dbias + o, accb);
}

// =============================================================================================
extern "C" int64_t howl_b200_las_param_count(int32_t num_labels, int32_t n_mels) {
  if (num_labels < 1 || n_mels < 1) return -1;
  return las_params(nullptr, LA_C * (n_mels + 4), num_labels).total;
}
extern "C" int64_t howl_b200_las_workspace_bytes(int64_t B, int32_t frames, int32_t n_mels, int32_t num_labels, int train) {
  if (B < 1 || frames < 4 || n_mels < 1 || num_labels < 1) return -1;
  return (int64_t)las_carve(nullptr, B, las_dims(n_mels, frames), train, num_labels).bytes;
}
// LASEncoder.forward's length arithmetic (rnn.py:163-168), host, float floor at every step as the reference
extern "C" int howl_b200_las_lengths(const int64_t* lengths, int64_t n, int64_t* out) {
  if (!lengths || !out || n < 0) return HOWL_E_INVALID;
  for (int64_t i = 0; i < n; ++i) {
    float l = floorf(((float)lengths[i] - 3.f + 4.f) / 1.f + 1.f);
    l = floorf(l / 2.f);
    l = floorf((l - 3.f + 4.f) / 1.f + 1.f);
    l = floorf(l / 2.f);
    out[i] = (int64_t)l;
  }
  return HOWL_OK;
}

static size_t las_head_smem(const LasDims& d) { return sizeof(float) * ((size_t)d.w2p * LA_D + d.w2p * LA_HEADS + LA_D + LA_DNN + LA_HEADS * LA_D); }
static size_t las_head_bwd_smem(const LasDims& d, int L) {
  return sizeof(float) * ((size_t)d.w2p * LA_D + 2 * d.w2p * LA_HEADS + LA_DNN + LA_D + 2 * LA_HEADS * LA_D + L);
}

extern "C" int howl_b200_las_fwd(howl_ctx_t* ctx, void* stream, const float* feats, const int64_t* enc_lengths, int64_t B, int32_t frames,
                                 int32_t n_mels, int32_t num_labels, const float* params, float* bn_running, int64_t* num_batches_tracked,
                                 int train, float dropout_p, uint64_t seed, float* logits, void* workspace, size_t workspace_bytes) {
  if (!ctx) return HOWL_E_INVALID;
  HOWL_REQUIRE(ctx, feats && enc_lengths && params && bn_running && logits && workspace, HOWL_E_INVALID, "las_fwd: null pointer");
  HOWL_REQUIRE(ctx, B >= 1 && frames >= 4 && n_mels >= 1 && num_labels >= 1 && num_labels <= 256, HOWL_E_INVALID, "las_fwd: bad shape");
  HOWL_REQUIRE(ctx, dropout_p >= 0.f && dropout_p < 1.f, HOWL_E_INVALID, "las_fwd: dropout_p %g outside [0, 1)", (double)dropout_p);
  const LasDims d = las_dims(n_mels, frames);
  LasWs ws = las_carve(workspace, B, d, train, num_labels);
  HOWL_REQUIRE(ctx, ws.bytes <= workspace_bytes, HOWL_E_WORKSPACE, "las_fwd: workspace %zu < required %zu", workspace_bytes, ws.bytes);
  const size_t lstm_smem = sizeof(float) * ((size_t)(d.in + LA_H) * LA_NB + LA_NB * LA_G + LA_NB * LA_H);
  const size_t head_smem = las_head_smem(d);
  HOWL_REQUIRE(ctx, lstm_smem <= 200 * 1024 && head_smem <= 200 * 1024, HOWL_E_UNSUPPORTED, "las_fwd: %d mels x %d frames exceed the shared-memory tiles", n_mels, frames);
  cudaStream_t st = (cudaStream_t)stream;
  const LasParams q = las_params(params, d.in, num_labels);
  const int blocks = ctx->sm_count * 8;
  if (train) HOWL_CUDA(ctx, cudaMemsetAsync(ws.stats, 0, sizeof(double) * 4 * LA_C, st));
  // bn_running: [2 layers][2][8] (mean, var), num_batches_tracked [2]
  las_conv_kernel<3><<<blocks, 256, 0, st>>>(feats, q.c1w, q.c1b, ws.raw1, B, d.M, d.F, train ? ws.stats : nullptr);
  HOWL_LAUNCHED(ctx, "las_conv1");
  las_bn_finalize_kernel<<<1, 32, 0, st>>>(ws.stats, (double)B * d.h1 * d.w1, q.bn1g, q.bn1b, bn_running, bn_running + LA_C,
                                           num_batches_tracked, train, ws.bn);
  HOWL_LAUNCHED(ctx, "las_bn_finalize");
  las_bn_relu_pool_kernel<<<blocks, 256, 0, st>>>(ws.raw1, ws.bn, ws.pool1, B, d.h1, d.w1, d.w1p, 0);
  HOWL_LAUNCHED(ctx, "las_bn_relu_pool");
  las_conv_kernel<LA_C><<<blocks, 256, 0, st>>>(ws.pool1, q.c2w, q.c2b, ws.raw2, B, d.h1, d.w1p, train ? ws.stats + 2 * LA_C : nullptr);
  HOWL_LAUNCHED(ctx, "las_conv2");
  las_bn_finalize_kernel<<<1, 32, 0, st>>>(ws.stats + 2 * LA_C, (double)B * d.h2 * d.w2, q.bn2g, q.bn2b, bn_running + 2 * LA_C,
                                           bn_running + 3 * LA_C, num_batches_tracked ? num_batches_tracked + 1 : nullptr, train,
                                           ws.bn + 4 * LA_C);
  HOWL_LAUNCHED(ctx, "las_bn_finalize");
  las_bn_relu_pool_kernel<<<blocks, 256, 0, st>>>(ws.raw2, ws.bn + 4 * LA_C, ws.x, B, d.h2, d.w2, d.w2p, 1);
  HOWL_LAUNCHED(ctx, "las_bn_relu_pool");
  las_lstm_prep_kernel<<<blocks, 256, 0, st>>>(q, d.in, ws.wt, ws.bsum);
  HOWL_LAUNCHED(ctx, "las_lstm_prep");
  HOWL_CUDA(ctx, cudaFuncSetAttribute(las_lstm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lstm_smem));
  const float* xproj = nullptr;
  if (train && (int64_t)d.w2p * B >= 1024) {
    // input projections of every step and both directions on the tensor cores: xproj[:, dir] = x W_ih[dir]^T + b_ih + b_hh
    const int64_t TB = (int64_t)d.w2p * B;
    int rc;
    if ((rc = mbn_pack3(ctx, st, ws.x, d.in, TB, d.in, ws.p3))) return rc;
    for (int dir = 0; dir < 2; ++dir) {
      if ((rc = mbn_weight_operand3(ctx, st, q.wih[dir], LA_G, d.in, d.in, 0, ws.wop3))) return rc;
      if ((rc = mbn_gemm_nt3_f32(ctx, st, ws.p3, ws.wop3, ws.xproj + dir * LA_G, 2 * LA_G, TB, d.in, LA_G, ws.bsum + dir * LA_G, 0))) return rc;
    }
    xproj = ws.xproj;
  }
  las_lstm_kernel<<<dim3((unsigned)howl_ceil_div(B, LA_NB), 2), LA_G, lstm_smem, st>>>(ws.x, enc_lengths, ws.wt, ws.bsum, ws.hseq, B, d.w2p, d.in,
                                                                                       ws.gates, ws.cseq, xproj);
  HOWL_LAUNCHED(ctx, "las_lstm");
  las_attn_prep_kernel<<<(LA_HEADS * LA_D + LA_HEADS + 255) / 256, 256, 0, st>>>(q, ws.u);
  HOWL_LAUNCHED(ctx, "las_attn_prep");
  HOWL_CUDA(ctx, cudaFuncSetAttribute(las_head_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)head_smem));
  las_head_kernel<<<(unsigned)B, 256, head_smem, st>>>(ws.hseq, enc_lengths, q, B, d.w2p, num_labels, logits, train ? dropout_p : 0.f, seed,
                                                      ws.u, ws.scores, ws.ctxs, ws.hidd, ws.logits);
  HOWL_LAUNCHED(ctx, "las_head");
  return HOWL_OK;
}

static int las_gemm(howl_ctx_t* ctx, cudaStream_t st, const float* A, int64_t sa_m, int64_t sa_k, const float* Bm, int64_t sb_k, int64_t sb_n,
                    float* C, int64_t ldc, int64_t M, int N, int64_t K, bool accumulate) {
  if (M <= 0 || N <= 0 || K <= 0) return HOWL_OK;
  const int64_t tiles = howl_ceil_div(M, LG_BM) * howl_ceil_div(N, LG_BN);
  int split = 1;
  if (accumulate) {      // reductions over the batch: few output tiles, long K -> split K over the machine
    split = (int)std::min<int64_t>(std::max<int64_t>(1, (4 * ctx->sm_count) / tiles), std::max<int64_t>(1, K / 256));
    split = std::min(split, 512);
  }
  las_gemm_kernel<<<dim3((unsigned)howl_ceil_div(M, LG_BM), (unsigned)howl_ceil_div(N, LG_BN), split), 256, 0, st>>>(A, sa_m, sa_k, Bm, sb_k, sb_n, C,
                                                                                                                      ldc, M, N, K, accumulate ? 1 : 0);
  HOWL_LAUNCHED(ctx, "las_gemm");
  return HOWL_OK;
}
static int las_colsum(howl_ctx_t* ctx, cudaStream_t st, const float* A, int64_t rows, int cols, int64_t ld, float* out, float* out2) {
  const int ysplit = (int)std::min<int64_t>(std::max<int64_t>(1, rows / 64), 4 * ctx->sm_count);
  las_colsum_kernel<<<dim3((unsigned)howl_ceil_div(cols, 256), ysplit), 256, 0, st>>>(A, rows, cols, ld, out, out2);
  HOWL_LAUNCHED(ctx, "las_colsum");
  return HOWL_OK;
}

static int las_bwd_impl(howl_ctx_t* ctx, void* stream, const float* feats, const int64_t* enc_lengths, const int64_t* labels,
                        const float* dlogits_in, int64_t B, int32_t frames, int32_t n_mels, int32_t num_labels, int64_t loss_scale_batch,
                        const float* params, float* grads, float dropout_p, float* loss, void* workspace, size_t workspace_bytes) {
  HOWL_REQUIRE(ctx, B >= 1 && frames >= 4 && n_mels >= 1 && num_labels >= 1 && num_labels <= 256, HOWL_E_INVALID, "las_bwd: bad shape");
  HOWL_REQUIRE(ctx, dropout_p >= 0.f && dropout_p < 1.f, HOWL_E_INVALID, "las_bwd: dropout_p %g outside [0, 1)", (double)dropout_p);
  const LasDims d = las_dims(n_mels, frames);
  LasWs ws = las_carve(workspace, B, d, 1, num_labels);
  HOWL_REQUIRE(ctx, ws.bytes <= workspace_bytes, HOWL_E_WORKSPACE, "las_bwd: workspace %zu < required %zu", workspace_bytes, ws.bytes);
  const int L = num_labels, T = d.w2p, in = d.in;
  const int64_t TB = (int64_t)T * B;
  const size_t hb_smem = las_head_bwd_smem(d, L);
  HOWL_REQUIRE(ctx, hb_smem <= 200 * 1024, HOWL_E_UNSUPPORTED, "las_bwd: %d frames exceed the shared-memory tiles", frames);
  cudaStream_t st = (cudaStream_t)stream;
  const LasParams q = las_params(params, in, L);
  // gradient pointers: same offsets as the parameters
  auto G = [&](const float* p) { return grads + (p - params); };
  const int blocks = ctx->sm_count * 8;
  int rc;
  HOWL_CUDA(ctx, cudaMemsetAsync(grads, 0, sizeof(float) * q.total, st));
  HOWL_CUDA(ctx, cudaMemsetAsync(ws.loss_acc, 0, sizeof(double) * 2, st));
  HOWL_CUDA(ctx, cudaMemsetAsync(ws.bstats, 0, sizeof(double) * 4 * LA_C, st));
  HOWL_CUDA(ctx, cudaMemsetAsync(ws.red, 0, sizeof(float) * (LA_HEADS * LA_D + LA_HEADS), st));
  // ---- loss and head
  las_ce_kernel<<<(unsigned)howl_ceil_div(B, 128), 128, 0, st>>>(ws.logits, labels, dlogits_in, ws.dlogits, ws.loss_acc, B, L,
                                                                 1.f / (float)loss_scale_batch);
  HOWL_LAUNCHED(ctx, "las_ce");
  if (loss) {
    las_loss_out_kernel<<<1, 1, 0, st>>>(ws.loss_acc, loss);
    HOWL_LAUNCHED(ctx, "las_loss_out");
  }
  LasHeadBwd hb;      // ws.u: kept from the forward (same parameters)
  hb.hseq = ws.hseq; hb.scores = ws.scores; hb.hidd = ws.hidd; hb.dlogits = ws.dlogits; hb.u = ws.u;
  hb.dctx = ws.dctx; hb.hbar = ws.hbar; hb.dhid = ws.dhid; hb.du = ws.du; hb.dsum = ws.dsum; hb.dhseq = ws.dhseq;
  hb.B = B; hb.T = T; hb.L = L; hb.drop_p = dropout_p;
  HOWL_CUDA(ctx, cudaFuncSetAttribute(las_head_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)hb_smem));
  las_head_bwd_kernel<<<(unsigned)B, 256, hb_smem, st>>>(hb, q);
  HOWL_LAUNCHED(ctx, "las_head_bwd");
  // fc.3: dW[l][r] = sum_b dlogits[b][l] hidd[b][r];  fc.0: dW[r][k] = sum_b dhid[b][r] ctx[b][k];  k_proj per head
  if ((rc = las_gemm(ctx, st, ws.dlogits, 1, L, ws.hidd, LA_DNN, 1, G(q.f3w), LA_DNN, L, LA_DNN, B, true))) return rc;
  if ((rc = las_colsum(ctx, st, ws.dlogits, B, L, L, G(q.f3b), nullptr))) return rc;
  if ((rc = las_gemm(ctx, st, ws.dhid, 1, LA_DNN, ws.ctxs, LA_D, 1, G(q.f0w), LA_D, LA_DNN, LA_D, B, true))) return rc;
  if ((rc = las_colsum(ctx, st, ws.dhid, B, LA_DNN, LA_DNN, G(q.f0b), nullptr))) return rc;
  constexpr int HD = LA_D / LA_HEADS;
  for (int h = 0; h < LA_HEADS; ++h)
    if ((rc = las_gemm(ctx, st, ws.dctx + h * HD, 1, LA_D, ws.hbar + h * LA_D, LA_HEADS * LA_D, 1, G(q.kw) + (size_t)h * HD * LA_D, LA_D, HD, LA_D, B,
                       true)))
      return rc;
  if ((rc = las_colsum(ctx, st, ws.dctx, B, LA_D, LA_D, G(q.kb), nullptr))) return rc;
  if ((rc = las_colsum(ctx, st, ws.du, B, LA_HEADS * LA_D, LA_HEADS * LA_D, ws.red, nullptr))) return rc;
  if ((rc = las_colsum(ctx, st, ws.dsum, B, LA_HEADS, LA_HEADS, ws.red + LA_HEADS * LA_D, nullptr))) return rc;
  las_attn_grads_kernel<<<(LA_D * LA_D + 255) / 256, 256, 0, st>>>(q, ws.red, G(q.vw), G(q.vb), G(q.cvec));
  HOWL_LAUNCHED(ctx, "las_attn_grads");
  // ---- recurrence
  las_lstm_bwd_kernel<<<dim3((unsigned)howl_ceil_div(B, LA_NB), 2), LA_G, 0, st>>>(ws.gates, ws.cseq, ws.dhseq, enc_lengths, q.whh[0], q.whh[1],
                                                                                   ws.dgates, B, T);
  HOWL_LAUNCHED(ctx, "las_lstm_bwd");
  for (int dir = 0; dir < 2; ++dir) {
    const float* dg = ws.dgates + dir * LA_G;        // [TB][2 * 384], this direction's columns
    // dx (+)= da W_ih: large batches on the tensor cores (C[r][k] = sum_j da[r][j] W_ih[j][k]: the weight is stored [reduction][output])
    if (TB >= 1024) {
      if ((rc = mbn_weight_operand3(ctx, st, q.wih[dir], in, LA_G, in, 1, ws.wop3))) return rc;
      if ((rc = mbn_pack3(ctx, st, dg, 2 * LA_G, TB, LA_G, ws.p3))) return rc;
      if ((rc = mbn_gemm_nt3_f32(ctx, st, ws.p3, ws.wop3, ws.dx, in, TB, LA_G, in, nullptr, 0, dir == 1))) return rc;
    } else if ((rc = las_gemm(ctx, st, dg, 2 * LA_G, 1, q.wih[dir], in, 1, ws.dx, in, TB, in, LA_G, dir == 1))) {
      return rc;
    }
    // dW_ih = da^T x and dW_hh = da^T h_prev.  The previous step of the forward direction is t - 1, of the reverse direction t + 1
    // (rows shifted by B).  Large batches: tensor cores, (hi, lo) bf16 operands, three products each, fp32 accumulate.
    const float* da = dir ? dg : dg + (size_t)B * 2 * LA_G;
    const float* hp = dir ? ws.hseq + (size_t)B * LA_D + LA_H : ws.hseq;
    if (TB >= 1024) {
      if ((rc = mbn_pack_split(ctx, st, dg, 2 * LA_G, TB, LA_G, ws.px_hi, ws.px_lo, G(q.bih[dir]), G(q.bhh[dir])))) return rc;      // + bias gradients
      if ((rc = mbn_pack_split(ctx, st, ws.x, in, TB, in, ws.py_hi, ws.py_lo))) return rc;
      if ((rc = mbn_atb3_packed(ctx, st, ws.px_hi, ws.px_lo, ws.py_hi, ws.py_lo, G(q.wih[dir]), TB, LA_G, in, in))) return rc;
      if (T > 1) {
        if ((rc = mbn_pack_split(ctx, st, da, 2 * LA_G, TB - B, LA_G, ws.px_hi, ws.px_lo))) return rc;
        if ((rc = mbn_pack_split(ctx, st, hp, LA_D, TB - B, LA_H, ws.py_hi, ws.py_lo))) return rc;
        if ((rc = mbn_atb3_packed(ctx, st, ws.px_hi, ws.px_lo, ws.py_hi, ws.py_lo, G(q.whh[dir]), TB - B, LA_G, LA_H, LA_H))) return rc;
      }
    } else {
      if ((rc = las_gemm(ctx, st, dg, 1, 2 * LA_G, ws.x, in, 1, G(q.wih[dir]), in, LA_G, in, TB, true))) return rc;
      if (T > 1 && (rc = las_gemm(ctx, st, da, 1, 2 * LA_G, hp, LA_D, 1, G(q.whh[dir]), LA_H, LA_G, LA_H, TB - B, true))) return rc;
    }
    if (TB < 1024 && (rc = las_colsum(ctx, st, dg, TB, LA_G, 2 * LA_G, G(q.bih[dir]), G(q.bhh[dir])))) return rc;
  }
  // ---- BatchNorm 2 + ReLU + MaxPool, conv2
  las_bn_bwd_stats_kernel<<<dim3(ctx->sm_count, LA_C), 256, 0, st>>>(ws.raw2, ws.bn + 4 * LA_C, ws.dx, B, d.h2, d.w2, d.w2p, 1, ws.bstats + 2 * LA_C);
  HOWL_LAUNCHED(ctx, "las_bn_bwd_stats");
  las_bn_bwd_apply_kernel<<<blocks, 256, 0, st>>>(ws.raw2, ws.bn + 4 * LA_C, ws.dx, q.bn2g, B, d.h2, d.w2, d.w2p, 1, ws.bstats + 2 * LA_C,
                                                  (double)B * d.h2 * d.w2, ws.draw2, G(q.bn2g), G(q.bn2b));
  HOWL_LAUNCHED(ctx, "las_bn_bwd_apply");
  {
    const size_t smem = sizeof(float) * LW_ROWS * ((size_t)LA_C * d.w2 + LA_C * 3 * (d.w1p + 4));
    HOWL_CUDA(ctx, cudaFuncSetAttribute(las_conv_wgrad_kernel<LA_C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    las_conv_wgrad_kernel<LA_C><<<ctx->sm_count * 2, LW_ROWS * LA_C * LA_C * 3, smem, st>>>(ws.pool1, ws.draw2, B, d.h1, d.w1p, G(q.c2w), G(q.c2b));
    HOWL_LAUNCHED(ctx, "las_conv_wgrad");
  }
  las_conv_dgrad_kernel<<<blocks, 256, 0, st>>>(ws.draw2, q.c2w, ws.dpool1, B, d.h1, d.w1p);
  HOWL_LAUNCHED(ctx, "las_conv_dgrad");
  // ---- BatchNorm 1 + ReLU + MaxPool, conv1
  las_bn_bwd_stats_kernel<<<dim3(ctx->sm_count, LA_C), 256, 0, st>>>(ws.raw1, ws.bn, ws.dpool1, B, d.h1, d.w1, d.w1p, 0, ws.bstats);
  HOWL_LAUNCHED(ctx, "las_bn_bwd_stats");
  las_bn_bwd_apply_kernel<<<blocks, 256, 0, st>>>(ws.raw1, ws.bn, ws.dpool1, q.bn1g, B, d.h1, d.w1, d.w1p, 0, ws.bstats, (double)B * d.h1 * d.w1,
                                                  ws.draw1, G(q.bn1g), G(q.bn1b));
  HOWL_LAUNCHED(ctx, "las_bn_bwd_apply");
  {
    const size_t smem = sizeof(float) * LW_ROWS * ((size_t)LA_C * d.w1 + 3 * 3 * (d.F + 4));
    HOWL_CUDA(ctx, cudaFuncSetAttribute(las_conv_wgrad_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    las_conv_wgrad_kernel<3><<<ctx->sm_count * 6, LW_ROWS * LA_C * 3 * 3, smem, st>>>(feats, ws.draw1, B, d.M, d.F, G(q.c1w), G(q.c1b));
    HOWL_LAUNCHED(ctx, "las_conv_wgrad");
  }
  return HOWL_OK;
}

extern "C" int howl_b200_las_bwd(howl_ctx_t* ctx, void* stream, const float* feats, const int64_t* enc_lengths, const int64_t* labels, int64_t B,
                                 int32_t frames, int32_t n_mels, int32_t num_labels, int64_t loss_scale_batch, const float* params, float* grads,
                                 float dropout_p, float* loss, void* workspace, size_t workspace_bytes) {
  if (!ctx) return HOWL_E_INVALID;
  HOWL_REQUIRE(ctx, feats && enc_lengths && labels && params && grads && workspace, HOWL_E_INVALID, "las_bwd: null pointer");
  HOWL_REQUIRE(ctx, loss_scale_batch >= 1, HOWL_E_INVALID, "las_bwd: loss_scale_batch must be >= 1");
  return las_bwd_impl(ctx, stream, feats, enc_lengths, labels, nullptr, B, frames, n_mels, num_labels, loss_scale_batch, params, grads, dropout_p,
                      loss, workspace, workspace_bytes);
}
extern "C" int howl_b200_las_bwd_dlogits(howl_ctx_t* ctx, void* stream, const float* feats, const int64_t* enc_lengths, const float* dlogits,
                                         int64_t B, int32_t frames, int32_t n_mels, int32_t num_labels, const float* params, float* grads,
                                         float dropout_p, void* workspace, size_t workspace_bytes) {
  if (!ctx) return HOWL_E_INVALID;
  HOWL_REQUIRE(ctx, feats && enc_lengths && dlogits && params && grads && workspace, HOWL_E_INVALID, "las_bwd_dlogits: null pointer");
  return las_bwd_impl(ctx, stream, feats, enc_lengths, nullptr, dlogits, B, frames, n_mels, num_labels, 1, params, grads, dropout_p, nullptr,
                      workspace, workspace_bytes);
}
