// LASClassifier forward (howl/model/rnn.py:133-215) for sm_100a, exact fp32:
//   encoder: Conv2d(3, 8, 3, padding=2) + BatchNorm2d + ReLU + MaxPool2d((1, 2)), Conv2d(8, 8, 3, padding=2) + BatchNorm2d + ReLU + MaxPool2d((1, 2)),
//            [F'', B, 8 * 44] -> bidirectional LSTM(352 -> 96) over each clip's own length (pack_padded_sequence semantics, any order)
//   FixedAttentionModule: 4 heads, a fixed context vector scores the value projection, softmax over time (padding masked with -100),
//            the scores average the key projection;  fc: Linear(192 -> 256) + ReLU + Dropout (identity here) + Linear(256 -> L).
// Forward only (inference, and the batch-statistics forward of train mode with the running-stat update); the backward of this model is not
// built.  Small convolutions and the attention are plain CUDA-core kernels; the recurrence is batch-parallel like K5
// (lstm.cu): one CTA owns 16 sequences of one direction, thread = gate row, [x_t | h] in shared memory, weights streamed from L2.
#include <math.h>

#include "common.cuh"

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
  float* bn;        // [2 layers][2][8] scale, shift
  size_t bytes;
};
static LasWs las_carve(void* base, int64_t B, const LasDims& d) {
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
  w.bn = (float*)take(sizeof(float) * 2 * 2 * LA_C);
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

// statistics (train) or running statistics (eval) -> fused scale / shift; train also updates the running statistics
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
                                                          const float* __restrict__ bsum, float* __restrict__ hseq, int64_t B, int T, int in) {
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
    // x_t of the 16 sequences -> s_xh[k][b]
    for (int i = j; i < in * LA_NB; i += LA_G) {
      const int bb = i / in, k = i - bb * in;
      s_xh[(size_t)k * LA_NB + bb] = (b0 + bb < B) ? __ldg(x + ((int64_t)t * B + b0 + bb) * in + k) : 0.f;
    }
    __syncthreads();
    float acc[LA_NB];
#pragma unroll
    for (int bb = 0; bb < LA_NB; ++bb) acc[bb] = bias;
    for (int k = 0; k < K; ++k) {
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
    for (int bb = 0; bb < LA_NB; ++bb) s_g[bb * LA_G + j] = tanh_gate ? tanhf(acc[bb]) : 1.f / (1.f + expf(-acc[bb]));
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
        hval = g[3 * LA_H + u] * tanhf(c);
        s_xh[(size_t)(in + u) * LA_NB + bb] = hval;
      }
      if (b0 + bb < B) hseq[((int64_t)t * B + b0 + bb) * LA_D + dir * LA_H + u] = hval;
    }
    __syncthreads();
  }
}

// attention + MLP head: one CTA per utterance
__global__ void __launch_bounds__(256) las_head_kernel(const float* __restrict__ hseq, const int64_t* __restrict__ lengths, LasParams q, int64_t B,
                                                       int T, int L, float* __restrict__ logits) {
  extern __shared__ __align__(16) float smem[];
  float* s_h = smem;                       // [T][192]
  float* s_sc = s_h + (size_t)T * LA_D;    // [T][4] attention logits -> scores
  float* s_ctx = s_sc + T * LA_HEADS;      // [192]
  float* s_hid = s_ctx + LA_D;             // [256]
  const int64_t b = blockIdx.x;
  const int tid = threadIdx.x, len = (int)(lengths[b] < (int64_t)T ? lengths[b] : (int64_t)T);
  for (int i = tid; i < T * LA_D; i += 256) s_h[i] = hseq[((int64_t)(i / LA_D) * B + b) * LA_D + (i % LA_D)];
  __syncthreads();
  // attention logits: logits[t][h] = sum_l values[t][h * 48 + l] * context_vec.view(48, 4)[l][h], values = v_proj(h_t)
  for (int i = tid; i < T * LA_HEADS; i += 256) {
    const int t = i / LA_HEADS, h = i - t * LA_HEADS;
    float acc = 0.f;
    for (int l = 0; l < LA_D / LA_HEADS; ++l) {
      const int row = h * (LA_D / LA_HEADS) + l;
      float v = q.vb[row];
      for (int k = 0; k < LA_D; ++k) v = fmaf(__ldg(q.vw + row * LA_D + k), s_h[t * LA_D + k], v);
      acc = fmaf(v, q.cvec[l * LA_HEADS + h], acc);
    }
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
  // context[h * 48 + l] = sum_t scores[t][h] * keys[t][h * 48 + l] = k_proj( sum_t scores[t][h] h_t )[row] (+ bias: the scores sum to 1)
  for (int row = tid; row < LA_D; row += 256) {
    const int h = row / (LA_D / LA_HEADS);
    float acc = q.kb[row];
    for (int k = 0; k < LA_D; ++k) {
      float hb = 0.f;
      for (int t = 0; t < T; ++t) hb = fmaf(s_sc[t * LA_HEADS + h], s_h[t * LA_D + k], hb);
      acc = fmaf(__ldg(q.kw + row * LA_D + k), hb, acc);
    }
    s_ctx[row] = acc;
  }
  __syncthreads();
  for (int r = tid; r < LA_DNN; r += 256) {
    float acc = q.f0b[r];
    for (int k = 0; k < LA_D; ++k) acc = fmaf(__ldg(q.f0w + r * LA_D + k), s_ctx[k], acc);
    s_hid[r] = fmaxf(acc, 0.f);
  }
  __syncthreads();
  for (int l = tid; l < L; l += 256) {
    float acc = q.f3b[l];
    for (int k = 0; k < LA_DNN; ++k) acc = fmaf(__ldg(q.f3w + (size_t)l * LA_DNN + k), s_hid[k], acc);
    logits[b * L + l] = acc;
  }
}

// =============================================================================================
extern "C" int64_t howl_b200_las_param_count(int32_t num_labels, int32_t n_mels) {
  if (num_labels < 1 || n_mels < 1) return -1;
  return las_params(nullptr, LA_C * (n_mels + 4), num_labels).total;
}
extern "C" int64_t howl_b200_las_workspace_bytes(int64_t B, int32_t frames, int32_t n_mels) {
  if (B < 1 || frames < 4 || n_mels < 1) return -1;
  return (int64_t)las_carve(nullptr, B, las_dims(n_mels, frames)).bytes;
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

extern "C" int howl_b200_las_fwd(howl_ctx_t* ctx, void* stream, const float* feats, const int64_t* enc_lengths, int64_t B, int32_t frames,
                                 int32_t n_mels, int32_t num_labels, const float* params, float* bn_running, int64_t* num_batches_tracked,
                                 int train, float* logits, void* workspace, size_t workspace_bytes) {
  if (!ctx) return HOWL_E_INVALID;
  HOWL_REQUIRE(ctx, feats && enc_lengths && params && bn_running && logits && workspace, HOWL_E_INVALID, "las_fwd: null pointer");
  HOWL_REQUIRE(ctx, B >= 1 && frames >= 4 && n_mels >= 1 && num_labels >= 1 && num_labels <= 256, HOWL_E_INVALID, "las_fwd: bad shape");
  const LasDims d = las_dims(n_mels, frames);
  LasWs ws = las_carve(workspace, B, d);
  HOWL_REQUIRE(ctx, ws.bytes <= workspace_bytes, HOWL_E_WORKSPACE, "las_fwd: workspace %zu < required %zu", workspace_bytes, ws.bytes);
  const size_t lstm_smem = sizeof(float) * ((size_t)(d.in + LA_H) * LA_NB + LA_NB * LA_G + LA_NB * LA_H);
  const size_t head_smem = sizeof(float) * ((size_t)d.w2p * LA_D + d.w2p * LA_HEADS + LA_D + LA_DNN);
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
                                           ws.bn + 2 * LA_C);
  HOWL_LAUNCHED(ctx, "las_bn_finalize");
  las_bn_relu_pool_kernel<<<blocks, 256, 0, st>>>(ws.raw2, ws.bn + 2 * LA_C, ws.x, B, d.h2, d.w2, d.w2p, 1);
  HOWL_LAUNCHED(ctx, "las_bn_relu_pool");
  las_lstm_prep_kernel<<<blocks, 256, 0, st>>>(q, d.in, ws.wt, ws.bsum);
  HOWL_LAUNCHED(ctx, "las_lstm_prep");
  HOWL_CUDA(ctx, cudaFuncSetAttribute(las_lstm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lstm_smem));
  las_lstm_kernel<<<dim3((unsigned)howl_ceil_div(B, LA_NB), 2), LA_G, lstm_smem, st>>>(ws.x, enc_lengths, ws.wt, ws.bsum, ws.hseq, B, d.w2p, d.in);
  HOWL_LAUNCHED(ctx, "las_lstm");
  HOWL_CUDA(ctx, cudaFuncSetAttribute(las_head_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)head_smem));
  las_head_kernel<<<(unsigned)B, 256, head_smem, st>>>(ws.hseq, enc_lengths, q, B, d.w2p, num_labels, logits);
  HOWL_LAUNCHED(ctx, "las_head");
  return HOWL_OK;
}
