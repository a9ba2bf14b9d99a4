// MobileNetClassifier (howl/model/cnn.py:15-29: Conv2d(1,3,3,pad=(1,3)) + BatchNorm + ReLU + MaxPool(1,2), then torchvision's MobileNetV2
// width 1.0 and a Linear head) -- forward, backward and the fused train step for sm_100a.
//
// bf16 activations / gradients with fp32 master weights, fp32 BatchNorm statistics and fp32 accumulation (BASELINE.json configs[2]).
// Every tensor between layers is in the tile-major operand format of mbn_common.cuh, so
//   * the 35 pointwise (1x1) convolutions and the 3x3 stride-2 entry convolution (as an im2col GEMM) -- >90 % of the flops -- and their
//     data / weight gradients run on the tensor cores (mbn_gemm.cu, tcgen05 + TMEM, operands landed by TMA);
//   * the depthwise 3x3 convolutions are stencils on 16-byte (8-channel) vectors over image tiles staged in shared memory (register-
//     resident images for the small late layers), with the producer's BatchNorm + ReLU6 applied while staging, so the expanded tensors
//     are stored once (raw) and never in normalised form;
//   * BatchNorm needs batch statistics before anything can be normalised: [conv -> statistics -> finalize -> consumer applies].
// Layer order, parameter order and names follow the reference's state_dict (SURVEY App. B.2 / BASELINE.md: 2,262,338 parameters at
// 30 labels).
#include <math.h>

#include <vector>

#include "mbn_common.cuh"
#include "../../include/howl_b200_debug.h"

#define MB_EPS 1e-5
#define MB_MOM 0.1
#define MB_LAST 1280
#define MB_MAXC 1280

// ---------------------------------------------------------------------------------------------
// network description (host)
// ---------------------------------------------------------------------------------------------
struct MbConv {
  int kind;            // 0 = stem (1->3, 3x3, bias), 1 = entry 3x3 stride 2 (im2col GEMM), 2 = pointwise, 3 = depthwise
  int cin, cout, stride;
  int hin, win, hout, wout;
  int act;             // 0 none, 1 ReLU6 (stem: ReLU + max pool, handled by its own kernels)
  int block;           // inverted-residual block index (0-based) or -1
  int residual;        // projection of a block with a skip connection
  size_t w_off, bias_off, g_off, b_off;   // offsets into the flat parameter buffer (floats)
  int bn_off;          // channel offset into the running statistics
  int bn_index;        // index into num_batches_tracked
};

struct MbNet {
  std::vector<MbConv> convs;
  size_t cls_w, cls_b, n_params;
  int n_bn_channels, n_bn;
  int h0, w0;          // spatial size after the stem (n_mels x (frames + 4) / 2)
};

static const int kMbSetting[7][4] = {{1, 16, 1, 1}, {6, 24, 2, 2}, {6, 32, 3, 2}, {6, 64, 4, 2}, {6, 96, 3, 1}, {6, 160, 3, 2}, {6, 320, 1, 1}};

static MbNet mb_build(int n_mels, int frames, int num_labels) {
  MbNet net;
  size_t off = 0;
  int bn_off = 0, bn_idx = 0;
  auto add = [&](int kind, int cin, int cout, int stride, int hin, int win, int act, int block, int residual) {
    MbConv c;
    memset(&c, 0, sizeof(c));
    c.kind = kind; c.cin = cin; c.cout = cout; c.stride = stride; c.hin = hin; c.win = win; c.act = act; c.block = block; c.residual = residual;
    if (kind == 0) {
      c.hout = hin; c.wout = (win + 4) / 2;       // conv pad (1,3): W + 4, then MaxPool(1,2)
    } else if (stride == 2) {
      c.hout = (hin - 1) / 2 + 1; c.wout = (win - 1) / 2 + 1;
    } else {
      c.hout = hin; c.wout = win;
    }
    c.w_off = off;
    off += (kind == 0) ? 27 : (kind == 1 ? (size_t)cout * cin * 9 : (kind == 2 ? (size_t)cout * cin : (size_t)cout * 9));
    if (kind == 0) { c.bias_off = off; off += 3; }
    c.g_off = off; off += cout;
    c.b_off = off; off += cout;
    c.bn_off = bn_off; bn_off += cout;
    c.bn_index = bn_idx++;
    net.convs.push_back(c);
    return c;
  };
  MbConv c = add(0, 1, 3, 1, n_mels, frames, 1, -1, 0);
  net.h0 = c.hout; net.w0 = c.wout;
  c = add(1, 3, 32, 2, c.hout, c.wout, 1, -1, 0);
  int inp = 32, h = c.hout, w = c.wout, blk = 0;
  for (int g = 0; g < 7; ++g) {
    for (int i = 0; i < kMbSetting[g][2]; ++i, ++blk) {
      const int t = kMbSetting[g][0], oup = kMbSetting[g][1], stride = (i == 0) ? kMbSetting[g][3] : 1, hidden = inp * t;
      if (t != 1) add(2, inp, hidden, 1, h, w, 1, blk, 0);
      c = add(3, hidden, hidden, stride, h, w, 1, blk, 0);
      h = c.hout; w = c.wout;
      add(2, hidden, oup, 1, h, w, 0, blk, (stride == 1 && inp == oup) ? 1 : 0);
      inp = oup;
    }
  }
  add(2, inp, MB_LAST, 1, h, w, 1, -1, 0);
  net.cls_w = off; off += (size_t)num_labels * MB_LAST;
  net.cls_b = off; off += num_labels;
  net.n_params = off;
  net.n_bn_channels = bn_off;
  net.n_bn = bn_idx;
  return net;
}

// ---------------------------------------------------------------------------------------------
// device helpers: 8 bf16 <-> 8 floats
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void mb_unpack(const uint4& q, float* v) {
  const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    v[2 * i] = __uint_as_float(w[i] << 16);
    v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
  }
}
__device__ __forceinline__ uint4 mb_pack(const float* v) {
  uint32_t o[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
    o[i] = *reinterpret_cast<const uint32_t*>(&h);
  }
  return make_uint4(o[0], o[1], o[2], o[3]);
}
__device__ __forceinline__ float mb_bf16_round(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

// block-wide sum of `n` per-thread fp32 partials per channel -> atomicAdd(double); n <= 16 values per thread, blockDim 256
template <int N>
__device__ __forceinline__ void mb_block_accumulate(const float* part, double* dst, const bool* ok) {
  __shared__ float sh[8][N];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < N; ++i) {
    const float s = warp_sum(part[i]);
    if (lane == 0) sh[warp][i] = s;
  }
  __syncthreads();
  if (threadIdx.x < N && ok[threadIdx.x]) {
    double t = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += (double)sh[w][threadIdx.x];
    atomicAdd(dst + threadIdx.x, t);
  }
  __syncthreads();
}

// per-channel BatchNorm constants of one layer as the kernels use them
struct MbBn {
  const float* scale;   // gamma * rstd
  const float* shift;   // beta - mean * gamma * rstd
  const float* mean;
  const float* rstd;
  const float* gamma;
};

// ---------------------------------------------------------------------------------------------
// per-channel sum / sum of squares of a TMO tensor (valid rows).  grid = (row blocks, chunks); stats double [2][cp]
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) mbn_stats_kernel(const uint4* __restrict__ x, int64_t rows, int c8, double* __restrict__ stats, int cp) {
  const int chunk = blockIdx.y;
  float part[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) part[i] = 0.f;
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += (int64_t)gridDim.x * blockDim.x) {
    float v[8];
    mb_unpack(x[mbn_vec(r, chunk, c8)], v);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      part[j] += v[j];
      part[8 + j] = fmaf(v[j], v[j], part[8 + j]);
    }
  }
  __shared__ float sh[8][16];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const float s = warp_sum(part[i]);
    if (lane == 0) sh[warp][i] = s;
  }
  __syncthreads();
  if (threadIdx.x < 16) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += (double)sh[w][threadIdx.x];
    const int which = threadIdx.x >> 3, j = threadIdx.x & 7;
    atomicAdd(stats + (size_t)which * cp + chunk * 8 + j, t);
  }
}

// statistics -> mean / rstd (+ running statistics, momentum 0.1, unbiased variance) -> fused scale / shift.  eval: running statistics.
__global__ void mbn_bn_finalize_kernel(const double* __restrict__ stats, double count, int c, int cp, const float* __restrict__ gamma,
                                       const float* __restrict__ beta, float* __restrict__ running_mean, float* __restrict__ running_var,
                                       int64_t* __restrict__ nbt, int train, float* __restrict__ out /* [5][cp]: scale, shift, mean, rstd, gamma */) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0 && train && nbt) *nbt += 1;
  if (i >= cp) return;
  float sc = 0.f, sh = 0.f, mu = 0.f, rs = 0.f, g = 0.f;
  if (i < c) {
    double mean, var;
    if (train) {
      mean = stats[i] / count;
      var = stats[cp + i] / count - mean * mean;
      if (var < 0.0) var = 0.0;
      const double unbiased = count > 1.0 ? var * count / (count - 1.0) : var;
      running_mean[i] = (float)((1.0 - MB_MOM) * running_mean[i] + MB_MOM * mean);
      running_var[i] = (float)((1.0 - MB_MOM) * running_var[i] + MB_MOM * unbiased);
    } else {
      mean = running_mean[i];
      var = running_var[i];
    }
    const double r = 1.0 / sqrt(var + MB_EPS);
    g = gamma[i];
    mu = (float)mean;
    rs = (float)r;
    sc = (float)(g * r);
    sh = (float)(beta[i] - mean * g * r);
  }
  out[i] = sc;
  out[cp + i] = sh;
  out[2 * cp + i] = mu;
  out[3 * cp + i] = rs;
  out[4 * cp + i] = g;
}

// out = act(raw * scale + shift) (+ residual); pad rows -> 0.  grid = (row blocks, chunks)
__global__ void __launch_bounds__(256) mbn_apply_kernel(const uint4* __restrict__ raw, const float* __restrict__ bn, int cp, int act,
                                                        const uint4* __restrict__ residual, uint4* __restrict__ out, int64_t rows,
                                                        int64_t rows_pad) {
  const int chunk = blockIdx.y, c8 = cp / 8;
  float sc[8], sh[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    sc[j] = bn[chunk * 8 + j];
    sh[j] = bn[cp + chunk * 8 + j];
  }
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < rows_pad; r += (int64_t)gridDim.x * blockDim.x) {
    const size_t idx = mbn_vec(r, chunk, c8);
    float v[8];
    if (r < rows) {
      mb_unpack(raw[idx], v);
      float res[8];
      if (residual) mb_unpack(residual[idx], res);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float y = fmaf(v[j], sc[j], sh[j]);
        if (act) y = fminf(fmaxf(y, 0.f), 6.f);
        if (residual) y += res[j];
        v[j] = y;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = 0.f;
    }
    out[idx] = mb_pack(v);
  }
}

// ---------------------------------------------------------------------------------------------
// depthwise 3x3 (pad 1, stride s).  All three kernels work on IMAGE TILES: a CTA owns one 8-channel chunk and walks groups of `ni`
// consecutive utterances; the tensor the stencil reads (the activated input, or the output gradient) is staged ONCE per group in shared
// memory as fp32 with a one-pixel zero border (two float4 planes -> conflict-free 128-bit reads), so BatchNorm + ReLU6 of the producer
// is applied once per element instead of once per tap, the taps need no bounds tests, and the nine reads per output hit shared memory.
// Pixel indices are decoded with multiply-high divisions by host-precomputed reciprocals.
// ---------------------------------------------------------------------------------------------
#define DW_PX 1024          // padded pixels of one staged tile (2 planes x 16 B x 1024 = 32 KB)

struct MbFastDiv { unsigned m, d; };
static MbFastDiv mb_fastdiv(int d) {
  MbFastDiv f;
  f.d = (unsigned)d;
  f.m = d > 1 ? (unsigned)((1ull << 32) / (unsigned)d + 1) : 0u;      // exact for x < 2^32 / d
  return f;
}
__device__ __forceinline__ int mb_div(int x, const MbFastDiv f) { return f.d == 1 ? x : (int)__umulhi((unsigned)x, f.m); }

struct MbDwGeom {
  int ni, ph, pw;                  // utterances per group, padded tile height / width
  MbFastDiv fd_tile, fd_pw;        // / (ph * pw), / pw: staged pixel -> (image, row, column)
  MbFastDiv fd_hw, fd_w;           // / (h * w), / w of the tensor the compute loop walks
};
// staged tensor hs x ws (padded by one pixel), walked tensor hwalk x wwalk
static MbDwGeom mb_dw_geom(int64_t B, int hs, int ws, int hwalk, int wwalk) {
  MbDwGeom g;
  g.ph = hs + 2; g.pw = ws + 2;
  g.ni = (int)std::max<int64_t>(1, std::min<int64_t>(B, DW_PX / (g.ph * g.pw)));
  g.fd_tile = mb_fastdiv(g.ph * g.pw); g.fd_pw = mb_fastdiv(g.pw);
  g.fd_hw = mb_fastdiv(hwalk * wwalk); g.fd_w = mb_fastdiv(wwalk);
  return g;
}
static bool mb_dw_fits(int hs, int ws) { return (hs + 2) * (ws + 2) <= DW_PX; }

struct MbDwArgs {
  const uint4* in;       // raw producer output, TMO [B * hin * win][cp]
  const float* bn_in;    // producer's [5][cp]
  const float* w;        // [c][9] fp32
  uint4* out;            // raw TMO [B * hout * wout][cp]
  double* stats;         // [2][cp] of the output (bf16-rounded), or null
  int64_t B;
  int c, cp, hin, win, hout, wout, stride;
  MbDwGeom g;
};

// A tile holds at most DW_PX = 4 x blockDim pixels.  mb_dw_load issues this thread's (up to four) global loads of the padded tile of the
// group starting at utterance b0 (zero outside the h x w image; bit e of the result = pixel e is inside); mb_dw_store converts them to
// the fp32 planes, applying the producer's BatchNorm + ReLU6 when s_bn is given.  The kernels issue the loads of the NEXT group before
// they compute the current one, so the DRAM latency overlaps the stencil arithmetic.
__device__ __forceinline__ unsigned mb_dw_load(uint4 (&raw)[4], const uint4* __restrict__ in, const MbDwGeom& g, int b0, int nb, int h, int w, int chunk,
                                               int c8) {
  const int tile = g.ph * g.pw, hw = h * w;
  unsigned inside = 0u;
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const int i = threadIdx.x + e * 256;
    const int img = mb_div(i, g.fd_tile), q = i - img * tile;
    const int py = mb_div(q, g.fd_pw), px = q - py * g.pw;
    raw[e] = make_uint4(0u, 0u, 0u, 0u);
    if (i < nb * tile && py >= 1 && py <= h && px >= 1 && px <= w) {
      raw[e] = __ldg(in + mbn_vec((int64_t)(b0 + img) * hw + (py - 1) * w + (px - 1), chunk, c8));
      inside |= 1u << e;
    }
  }
  return inside;
}
__device__ __forceinline__ void mb_dw_store(float4* s_lo, float4* s_hi, const uint4 (&raw)[4], unsigned inside, int count, const float* s_bn) {
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const int i = threadIdx.x + e * 256;
    if (i >= count) break;
    float v[8];
    mb_unpack(raw[e], v);
    if (s_bn && (inside >> e & 1u)) {
      const float4 sc0 = *reinterpret_cast<const float4*>(s_bn), sc1 = *reinterpret_cast<const float4*>(s_bn + 4);
      const float4 sh0 = *reinterpret_cast<const float4*>(s_bn + 8), sh1 = *reinterpret_cast<const float4*>(s_bn + 12);
      const float sc[8] = {sc0.x, sc0.y, sc0.z, sc0.w, sc1.x, sc1.y, sc1.z, sc1.w}, sh[8] = {sh0.x, sh0.y, sh0.z, sh0.w, sh1.x, sh1.y, sh1.z, sh1.w};
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = fminf(fmaxf(fmaf(v[j], sc[j], sh[j]), 0.f), 6.f);
    }
    s_lo[i] = make_float4(v[0], v[1], v[2], v[3]);
    s_hi[i] = make_float4(v[4], v[5], v[6], v[7]);
  }
}

__global__ void __launch_bounds__(256, 2) mbn_dw_fwd_kernel(const MbDwArgs a) {
  const int chunk = blockIdx.y, c8 = a.cp / 8;
  __shared__ float4 s_lo[DW_PX], s_hi[DW_PX];
  __shared__ __align__(16) float s_w[9][8];       // this chunk's taps: read as warp-wide broadcasts
  if (threadIdx.x < 72) {
    const int k = threadIdx.x >> 3, c = chunk * 8 + (threadIdx.x & 7);
    s_w[k][threadIdx.x & 7] = c < a.c ? a.w[c * 9 + k] : 0.f;
  }
  __shared__ __align__(16) float s_bn[16];        // producer's scale[8], shift[8]
  if (threadIdx.x >= 96 && threadIdx.x < 112) {
    const int j = threadIdx.x - 96;
    s_bn[j] = a.bn_in[(j >> 3) * a.cp + chunk * 8 + (j & 7)];
  }
  const int hw_o = a.hout * a.wout, tile = a.g.ph * a.g.pw;
  const int groups = (int)((a.B + a.g.ni - 1) / a.g.ni);
  float part[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) part[i] = 0.f;
  uint4 raw[4];
  unsigned inside = 0u;
  if ((int)blockIdx.x < groups)
    inside = mb_dw_load(raw, a.in, a.g, blockIdx.x * a.g.ni, (int)min((int64_t)a.g.ni, a.B - (int64_t)blockIdx.x * a.g.ni), a.hin, a.win, chunk, c8);
  for (int grp = blockIdx.x; grp < groups; grp += gridDim.x) {
    const int b0 = grp * a.g.ni, nb = (int)min((int64_t)a.g.ni, a.B - b0);
    __syncthreads();
    mb_dw_store(s_lo, s_hi, raw, inside, nb * tile, s_bn);
    __syncthreads();
    const int nxt = grp + gridDim.x;
    if (nxt < groups) inside = mb_dw_load(raw, a.in, a.g, nxt * a.g.ni, (int)min((int64_t)a.g.ni, a.B - (int64_t)nxt * a.g.ni), a.hin, a.win, chunk, c8);
    for (int o = threadIdx.x; o < nb * hw_o; o += blockDim.x) {
      const int img = mb_div(o, a.g.fd_hw), p = o - img * hw_o;
      const int yo = mb_div(p, a.g.fd_w), xo = p - yo * a.wout;
      const int base = img * tile + yo * a.stride * a.g.pw + xo * a.stride;
      float acc[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] = 0.f;
#pragma unroll
      for (int ky = 0; ky < 3; ++ky)
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          const float4 lo = s_lo[base + ky * a.g.pw + kx], hi = s_hi[base + ky * a.g.pw + kx];
          const float4 w0 = *reinterpret_cast<const float4*>(&s_w[ky * 3 + kx][0]), w1 = *reinterpret_cast<const float4*>(&s_w[ky * 3 + kx][4]);
          acc[0] = fmaf(lo.x, w0.x, acc[0]); acc[1] = fmaf(lo.y, w0.y, acc[1]); acc[2] = fmaf(lo.z, w0.z, acc[2]); acc[3] = fmaf(lo.w, w0.w, acc[3]);
          acc[4] = fmaf(hi.x, w1.x, acc[4]); acc[5] = fmaf(hi.y, w1.y, acc[5]); acc[6] = fmaf(hi.z, w1.z, acc[6]); acc[7] = fmaf(hi.w, w1.w, acc[7]);
        }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        acc[j] = mb_bf16_round(acc[j]);
        part[j] += acc[j];
        part[8 + j] = fmaf(acc[j], acc[j], part[8 + j]);
      }
      a.out[mbn_vec((int64_t)b0 * hw_o + o, chunk, c8)] = mb_pack(acc);
    }
  }
  if (blockIdx.x == 0) {        // pad rows of the last tile
    const int64_t rows = a.B * hw_o, rows_pad = mbn_tiles(rows) * MBN_TILE;
    for (int64_t r = rows + threadIdx.x; r < rows_pad; r += blockDim.x) a.out[mbn_vec(r, chunk, c8)] = make_uint4(0u, 0u, 0u, 0u);
  }
  if (a.stats) {
    __shared__ float s_part[8][16];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const float s = warp_sum(part[i]);
      if (lane == 0) s_part[warp][i] = s;
    }
    __syncthreads();
    if (threadIdx.x < 16) {
      double t = 0.0;
      for (int wv = 0; wv < 8; ++wv) t += (double)s_part[wv][threadIdx.x];
      atomicAdd(a.stats + (size_t)(threadIdx.x >> 3) * a.cp + chunk * 8 + (threadIdx.x & 7), t);
    }
  }
}

// data gradient of the depthwise convolution: d(in_act)[b, yi, xi, c] = sum_{ky,kx} dOut[b, yo, xo, c] w[c][ky][kx], yo * s - 1 + ky = yi.
// The output gradient is the staged tensor (border = the rows / columns yo = -1, hout that some taps address).
struct MbDwBwdArgs {
  const uint4* dout;     // gradient at the raw depthwise output, TMO [B * hout * wout][cp]
  const float* w;
  uint4* din;            // gradient w.r.t. the depthwise input (post-activation), TMO [B * hin * win][cp]
  int64_t B;
  int c, cp, hin, win, hout, wout, stride;
  MbDwGeom g;
  // BatchNorm-backward statistics of the PRODUCER convolution, fused (tiled kernel only): din is exactly the gradient at the producer's
  // activated output, so S1 = sum dn and S2 = sum dn * xhat are accumulated here and the separate pass over (din, raw) is skipped
  const uint4* raw_prev; // producer's raw output, or null
  const float* bn_prev;  // producer's [5][cp]
  double* bstats;        // [2][cp], zeroed by the host
  int act_prev;
};

__global__ void __launch_bounds__(256, 2) mbn_dw_bwd_data_kernel(const MbDwBwdArgs a) {
  const int chunk = blockIdx.y, c8 = a.cp / 8;
  __shared__ float4 s_lo[DW_PX], s_hi[DW_PX];
  __shared__ __align__(16) float s_w[9][8];
  __shared__ __align__(16) float s_bn[32];        // producer's scale, shift, mean, rstd of this chunk
  if (threadIdx.x < 72) {
    const int k = threadIdx.x >> 3, c = chunk * 8 + (threadIdx.x & 7);
    s_w[k][threadIdx.x & 7] = c < a.c ? a.w[c * 9 + k] : 0.f;
  }
  if (a.bstats && threadIdx.x >= 96 && threadIdx.x < 128) {
    const int j = threadIdx.x - 96;
    s_bn[j] = a.bn_prev[(j >> 3) * a.cp + chunk * 8 + (j & 7)];
  }
  float part[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) part[i] = 0.f;
  const int hw_i = a.hin * a.win, tile = a.g.ph * a.g.pw;
  const int groups = (int)((a.B + a.g.ni - 1) / a.g.ni);
  uint4 raw[4];
  unsigned inside = 0u;
  if ((int)blockIdx.x < groups)
    inside = mb_dw_load(raw, a.dout, a.g, blockIdx.x * a.g.ni, (int)min((int64_t)a.g.ni, a.B - (int64_t)blockIdx.x * a.g.ni), a.hout, a.wout, chunk, c8);
  for (int grp = blockIdx.x; grp < groups; grp += gridDim.x) {
    const int b0 = grp * a.g.ni, nb = (int)min((int64_t)a.g.ni, a.B - b0);
    __syncthreads();
    mb_dw_store(s_lo, s_hi, raw, inside, nb * tile, nullptr);
    __syncthreads();
    const int nxt = grp + gridDim.x;
    if (nxt < groups) inside = mb_dw_load(raw, a.dout, a.g, nxt * a.g.ni, (int)min((int64_t)a.g.ni, a.B - (int64_t)nxt * a.g.ni), a.hout, a.wout, chunk, c8);
    for (int o = threadIdx.x; o < nb * hw_i; o += blockDim.x) {
      const int img = mb_div(o, a.g.fd_hw), p = o - img * hw_i;
      const int yi = mb_div(p, a.g.fd_w), xi = p - yi * a.win;
      const size_t oidx = mbn_vec((int64_t)b0 * hw_i + o, chunk, c8);
      uint4 xq = make_uint4(0u, 0u, 0u, 0u);
      if (a.bstats) xq = __ldg(a.raw_prev + oidx);
      float acc[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] = 0.f;
#pragma unroll
      for (int ky = 0; ky < 3; ++ky) {
        const int ty = yi + 1 - ky;                   // = yo * stride
        if (a.stride == 2 && (ty & 1)) continue;
        const int yo = a.stride == 2 ? ty >> 1 : ty;  // -1 .. hout: inside the padded tile
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          const int tx = xi + 1 - kx;
          if (a.stride == 2 && (tx & 1)) continue;
          const int xo = a.stride == 2 ? tx >> 1 : tx;
          const int idx = img * tile + (yo + 1) * a.g.pw + (xo + 1);
          const float4 lo = s_lo[idx], hi = s_hi[idx];
          const float4 w0 = *reinterpret_cast<const float4*>(&s_w[ky * 3 + kx][0]), w1 = *reinterpret_cast<const float4*>(&s_w[ky * 3 + kx][4]);
          acc[0] = fmaf(lo.x, w0.x, acc[0]); acc[1] = fmaf(lo.y, w0.y, acc[1]); acc[2] = fmaf(lo.z, w0.z, acc[2]); acc[3] = fmaf(lo.w, w0.w, acc[3]);
          acc[4] = fmaf(hi.x, w1.x, acc[4]); acc[5] = fmaf(hi.y, w1.y, acc[5]); acc[6] = fmaf(hi.z, w1.z, acc[6]); acc[7] = fmaf(hi.w, w1.w, acc[7]);
        }
      }
      const uint4 packed = mb_pack(acc);
      a.din[oidx] = packed;
      if (a.bstats) {
        float g[8], x[8];
        mb_unpack(packed, g);                     // the statistics see the gradient as stored (bf16), like the separate pass did
        mb_unpack(xq, x);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float dn = g[j];
          if (a.act_prev) {
            const float n = fmaf(x[j], s_bn[j], s_bn[8 + j]);
            if (!(n > 0.f && n < 6.f)) dn = 0.f;
          }
          part[j] += dn;
          part[8 + j] = fmaf(dn, (x[j] - s_bn[16 + j]) * s_bn[24 + j], part[8 + j]);
        }
      }
    }
  }
  if (blockIdx.x == 0) {
    const int64_t rows = a.B * hw_i, rows_pad = mbn_tiles(rows) * MBN_TILE;
    for (int64_t r = rows + threadIdx.x; r < rows_pad; r += blockDim.x) a.din[mbn_vec(r, chunk, c8)] = make_uint4(0u, 0u, 0u, 0u);
  }
  if (a.bstats) {
    __shared__ float s_part[8][16];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const float t = warp_sum(part[i]);
      if (lane == 0) s_part[warp][i] = t;
    }
    __syncthreads();
    if (threadIdx.x < 16) {
      double t = 0.0;
      for (int wv = 0; wv < 8; ++wv) t += (double)s_part[wv][threadIdx.x];
      atomicAdd(a.bstats + (size_t)(threadIdx.x >> 3) * a.cp + chunk * 8 + (threadIdx.x & 7), t);
    }
  }
}

// weight gradient of the depthwise convolution: dW[c][k] = sum_rows dOut[row][c] * in_act[row's tap k][c]
struct MbDwWgArgs {
  const uint4* dout;
  const uint4* in;       // raw producer output
  const float* bn_in;
  float* dw;             // [c][9] fp32, accumulated
  int64_t B;
  int c, cp, hin, win, hout, wout, stride;
  MbDwGeom g;
};

__global__ void __launch_bounds__(256, 2) mbn_dw_bwd_weight_kernel(const MbDwWgArgs a) {
  const int chunk = blockIdx.y, c8 = a.cp / 8;
  __shared__ float4 s_lo[DW_PX], s_hi[DW_PX];
  __shared__ __align__(16) float s_bn[16];
  if (threadIdx.x < 16) s_bn[threadIdx.x] = a.bn_in[(threadIdx.x >> 3) * a.cp + chunk * 8 + (threadIdx.x & 7)];
  float acc[8][9];
#pragma unroll
  for (int j = 0; j < 8; ++j)
#pragma unroll
    for (int k = 0; k < 9; ++k) acc[j][k] = 0.f;
  const int hw_o = a.hout * a.wout, tile = a.g.ph * a.g.pw;
  const int groups = (int)((a.B + a.g.ni - 1) / a.g.ni);
  uint4 raw[4];
  unsigned inside = 0u;
  if ((int)blockIdx.x < groups)
    inside = mb_dw_load(raw, a.in, a.g, blockIdx.x * a.g.ni, (int)min((int64_t)a.g.ni, a.B - (int64_t)blockIdx.x * a.g.ni), a.hin, a.win, chunk, c8);
  for (int grp = blockIdx.x; grp < groups; grp += gridDim.x) {
    const int b0 = grp * a.g.ni, nb = (int)min((int64_t)a.g.ni, a.B - b0);
    uint4 gq[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int o = threadIdx.x + e * 256;
      gq[e] = make_uint4(0u, 0u, 0u, 0u);
      if (o < nb * hw_o) gq[e] = __ldg(a.dout + mbn_vec((int64_t)b0 * hw_o + o, chunk, c8));
    }
    __syncthreads();
    mb_dw_store(s_lo, s_hi, raw, inside, nb * tile, s_bn);
    __syncthreads();
    const int nxt = grp + gridDim.x;
    if (nxt < groups) inside = mb_dw_load(raw, a.in, a.g, nxt * a.g.ni, (int)min((int64_t)a.g.ni, a.B - (int64_t)nxt * a.g.ni), a.hin, a.win, chunk, c8);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int o = threadIdx.x + e * 256;
      if (o >= nb * hw_o) break;
      const int img = mb_div(o, a.g.fd_hw), p = o - img * hw_o;
      const int yo = mb_div(p, a.g.fd_w), xo = p - yo * a.wout;
      const int base = img * tile + yo * a.stride * a.g.pw + xo * a.stride;
      float g[8];
      mb_unpack(gq[e], g);
#pragma unroll
      for (int ky = 0; ky < 3; ++ky)
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          const float4 lo = s_lo[base + ky * a.g.pw + kx], hi = s_hi[base + ky * a.g.pw + kx];
          const int k = ky * 3 + kx;
          acc[0][k] = fmaf(g[0], lo.x, acc[0][k]); acc[1][k] = fmaf(g[1], lo.y, acc[1][k]); acc[2][k] = fmaf(g[2], lo.z, acc[2][k]);
          acc[3][k] = fmaf(g[3], lo.w, acc[3][k]); acc[4][k] = fmaf(g[4], hi.x, acc[4][k]); acc[5][k] = fmaf(g[5], hi.y, acc[5][k]);
          acc[6][k] = fmaf(g[6], hi.z, acc[6][k]); acc[7][k] = fmaf(g[7], hi.w, acc[7][k]);
        }
    }
  }
  __shared__ float s_acc[72];
  if (threadIdx.x < 72) s_acc[threadIdx.x] = 0.f;
  __syncthreads();
#pragma unroll
  for (int j = 0; j < 8; ++j)
#pragma unroll
    for (int k = 0; k < 9; ++k) {
      const float s = warp_sum(acc[j][k]);
      if ((threadIdx.x & 31) == 0) atomicAdd(&s_acc[j * 9 + k], s);
    }
  __syncthreads();
  if (threadIdx.x < 72) {
    const int c = chunk * 8 + threadIdx.x / 9;
    if (c < a.c) atomicAdd(a.dw + c * 9 + threadIdx.x % 9, s_acc[threadIdx.x]);
  }
}

// ---------------------------------------------------------------------------------------------
// depthwise 3x3 on SMALL images (<= 3 x 3 pixels: the last ten depthwise layers at 1 s clips).  thread = (utterance, 8-channel chunk):
// the whole image lives in registers, every load is issued up front, taps are resolved at compile time -- no shared-memory tile, no
// barriers, no index arithmetic per pixel.  grid = (ceil(B / 128), chunks), block = 128.
// ---------------------------------------------------------------------------------------------
#define DWS_THREADS 128
template <int HIN, int WIN, int S>
struct MbSmall {
  static constexpr int HOUT = (HIN - 1) / S + 1, WOUT = (WIN - 1) / S + 1, NIN = HIN * WIN, NOUT = HOUT * WOUT;
};

template <int HIN, int WIN, int S>
__global__ void __launch_bounds__(DWS_THREADS) mbn_dw_small_fwd_kernel(const MbDwArgs a) {
  using G = MbSmall<HIN, WIN, S>;
  const int chunk = blockIdx.y, c8 = a.cp / 8;
  __shared__ __align__(16) float s_w[9][8];
  __shared__ float s_part[DWS_THREADS / 32][16];
  if (threadIdx.x < 72) {
    const int k = threadIdx.x >> 3, c = chunk * 8 + (threadIdx.x & 7);
    s_w[k][threadIdx.x & 7] = c < a.c ? a.w[c * 9 + k] : 0.f;
  }
  float sc[8], sh[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    sc[j] = a.bn_in[chunk * 8 + j];
    sh[j] = a.bn_in[a.cp + chunk * 8 + j];
  }
  __syncthreads();
  const int64_t b = (int64_t)blockIdx.x * DWS_THREADS + threadIdx.x;
  float part[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) part[i] = 0.f;
  if (b < a.B) {
    uint4 raw[G::NIN];
#pragma unroll
    for (int p = 0; p < G::NIN; ++p) raw[p] = __ldg(a.in + mbn_vec(b * G::NIN + p, chunk, c8));
    float x[G::NIN][8];
#pragma unroll
    for (int p = 0; p < G::NIN; ++p) {
      mb_unpack(raw[p], x[p]);
#pragma unroll
      for (int j = 0; j < 8; ++j) x[p][j] = fminf(fmaxf(fmaf(x[p][j], sc[j], sh[j]), 0.f), 6.f);
    }
#pragma unroll
    for (int yo = 0; yo < G::HOUT; ++yo)
#pragma unroll
      for (int xo = 0; xo < G::WOUT; ++xo) {
        float acc[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = 0.f;
#pragma unroll
        for (int ky = 0; ky < 3; ++ky)
#pragma unroll
          for (int kx = 0; kx < 3; ++kx) {
            const int yi = yo * S - 1 + ky, xi = xo * S - 1 + kx;
            if (yi < 0 || yi >= HIN || xi < 0 || xi >= WIN) continue;
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] = fmaf(x[yi * WIN + xi][j], s_w[ky * 3 + kx][j], acc[j]);
          }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          acc[j] = mb_bf16_round(acc[j]);
          part[j] += acc[j];
          part[8 + j] = fmaf(acc[j], acc[j], part[8 + j]);
        }
        a.out[mbn_vec(b * G::NOUT + yo * G::WOUT + xo, chunk, c8)] = mb_pack(acc);
      }
  }
  if (blockIdx.x == 0) {
    const int64_t rows = a.B * G::NOUT, rows_pad = mbn_tiles(rows) * MBN_TILE;
    for (int64_t r = rows + threadIdx.x; r < rows_pad; r += blockDim.x) a.out[mbn_vec(r, chunk, c8)] = make_uint4(0u, 0u, 0u, 0u);
  }
  if (a.stats) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const float t = warp_sum(part[i]);
      if (lane == 0) s_part[warp][i] = t;
    }
    __syncthreads();
    if (threadIdx.x < 16) {
      double t = 0.0;
      for (int wv = 0; wv < DWS_THREADS / 32; ++wv) t += (double)s_part[wv][threadIdx.x];
      atomicAdd(a.stats + (size_t)(threadIdx.x >> 3) * a.cp + chunk * 8 + (threadIdx.x & 7), t);
    }
  }
}

template <int HIN, int WIN, int S>
__global__ void __launch_bounds__(DWS_THREADS) mbn_dw_small_bwd_data_kernel(const MbDwBwdArgs a) {
  using G = MbSmall<HIN, WIN, S>;
  const int chunk = blockIdx.y, c8 = a.cp / 8;
  __shared__ __align__(16) float s_w[9][8];
  if (threadIdx.x < 72) {
    const int k = threadIdx.x >> 3, c = chunk * 8 + (threadIdx.x & 7);
    s_w[k][threadIdx.x & 7] = c < a.c ? a.w[c * 9 + k] : 0.f;
  }
  __syncthreads();
  const int64_t b = (int64_t)blockIdx.x * DWS_THREADS + threadIdx.x;
  if (b < a.B) {
    uint4 raw[G::NOUT];
#pragma unroll
    for (int p = 0; p < G::NOUT; ++p) raw[p] = __ldg(a.dout + mbn_vec(b * G::NOUT + p, chunk, c8));
    float g[G::NOUT][8];
#pragma unroll
    for (int p = 0; p < G::NOUT; ++p) mb_unpack(raw[p], g[p]);
#pragma unroll
    for (int yi = 0; yi < HIN; ++yi)
#pragma unroll
      for (int xi = 0; xi < WIN; ++xi) {
        float acc[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = 0.f;
#pragma unroll
        for (int ky = 0; ky < 3; ++ky)
#pragma unroll
          for (int kx = 0; kx < 3; ++kx) {
            const int ty = yi + 1 - ky, tx = xi + 1 - kx;
            if (ty < 0 || tx < 0 || ty % S || tx % S) continue;
            const int yo = ty / S, xo = tx / S;
            if (yo >= G::HOUT || xo >= G::WOUT) continue;
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] = fmaf(g[yo * G::WOUT + xo][j], s_w[ky * 3 + kx][j], acc[j]);
          }
        a.din[mbn_vec(b * G::NIN + yi * WIN + xi, chunk, c8)] = mb_pack(acc);
      }
  }
  if (blockIdx.x == 0) {
    const int64_t rows = a.B * G::NIN, rows_pad = mbn_tiles(rows) * MBN_TILE;
    for (int64_t r = rows + threadIdx.x; r < rows_pad; r += blockDim.x) a.din[mbn_vec(r, chunk, c8)] = make_uint4(0u, 0u, 0u, 0u);
  }
}

template <int HIN, int WIN, int S>
__global__ void __launch_bounds__(DWS_THREADS) mbn_dw_small_bwd_weight_kernel(const MbDwWgArgs a) {
  using G = MbSmall<HIN, WIN, S>;
  const int chunk = blockIdx.y, c8 = a.cp / 8;
  __shared__ float s_acc[72];
  if (threadIdx.x < 72) s_acc[threadIdx.x] = 0.f;
  float sc[8], sh[8], acc[8][9];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    sc[j] = a.bn_in[chunk * 8 + j];
    sh[j] = a.bn_in[a.cp + chunk * 8 + j];
#pragma unroll
    for (int k = 0; k < 9; ++k) acc[j][k] = 0.f;
  }
  __syncthreads();
  for (int64_t b = (int64_t)blockIdx.x * DWS_THREADS + threadIdx.x; b < a.B; b += (int64_t)gridDim.x * DWS_THREADS) {
    uint4 rin[G::NIN], rg[G::NOUT];
#pragma unroll
    for (int p = 0; p < G::NIN; ++p) rin[p] = __ldg(a.in + mbn_vec(b * G::NIN + p, chunk, c8));
#pragma unroll
    for (int p = 0; p < G::NOUT; ++p) rg[p] = __ldg(a.dout + mbn_vec(b * G::NOUT + p, chunk, c8));
    float x[G::NIN][8];
#pragma unroll
    for (int p = 0; p < G::NIN; ++p) {
      mb_unpack(rin[p], x[p]);
#pragma unroll
      for (int j = 0; j < 8; ++j) x[p][j] = fminf(fmaxf(fmaf(x[p][j], sc[j], sh[j]), 0.f), 6.f);
    }
#pragma unroll
    for (int yo = 0; yo < G::HOUT; ++yo)
#pragma unroll
      for (int xo = 0; xo < G::WOUT; ++xo) {
        float g[8];
        mb_unpack(rg[yo * G::WOUT + xo], g);
#pragma unroll
        for (int ky = 0; ky < 3; ++ky)
#pragma unroll
          for (int kx = 0; kx < 3; ++kx) {
            const int yi = yo * S - 1 + ky, xi = xo * S - 1 + kx;
            if (yi < 0 || yi >= HIN || xi < 0 || xi >= WIN) continue;
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j][ky * 3 + kx] = fmaf(g[j], x[yi * WIN + xi][j], acc[j][ky * 3 + kx]);
          }
      }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j)
#pragma unroll
    for (int k = 0; k < 9; ++k) {
      const float t = warp_sum(acc[j][k]);
      if ((threadIdx.x & 31) == 0) atomicAdd(&s_acc[j * 9 + k], t);
    }
  __syncthreads();
  if (threadIdx.x < 72) {
    const int c = chunk * 8 + threadIdx.x / 9;
    if (c < a.c) atomicAdd(a.dw + c * 9 + threadIdx.x % 9, s_acc[threadIdx.x]);
  }
}

// shape dispatch of the small-image kernels; false = not a small shape (the tiled kernels take it)
static bool mb_dw_small_fwd(cudaStream_t st, const MbDwArgs& d, int c8) {
  const dim3 grid((unsigned)howl_ceil_div(d.B, DWS_THREADS), c8);
  if (d.hin == 3 && d.win == 3 && d.stride == 1) mbn_dw_small_fwd_kernel<3, 3, 1><<<grid, DWS_THREADS, 0, st>>>(d);
  else if (d.hin == 3 && d.win == 3 && d.stride == 2) mbn_dw_small_fwd_kernel<3, 3, 2><<<grid, DWS_THREADS, 0, st>>>(d);
  else if (d.hin == 2 && d.win == 2 && d.stride == 1) mbn_dw_small_fwd_kernel<2, 2, 1><<<grid, DWS_THREADS, 0, st>>>(d);
  else return false;
  return true;
}
static bool mb_dw_small_bwd_data(cudaStream_t st, const MbDwBwdArgs& d, int c8) {
  const dim3 grid((unsigned)howl_ceil_div(d.B, DWS_THREADS), c8);
  if (d.hin == 3 && d.win == 3 && d.stride == 1) mbn_dw_small_bwd_data_kernel<3, 3, 1><<<grid, DWS_THREADS, 0, st>>>(d);
  else if (d.hin == 3 && d.win == 3 && d.stride == 2) mbn_dw_small_bwd_data_kernel<3, 3, 2><<<grid, DWS_THREADS, 0, st>>>(d);
  else if (d.hin == 2 && d.win == 2 && d.stride == 1) mbn_dw_small_bwd_data_kernel<2, 2, 1><<<grid, DWS_THREADS, 0, st>>>(d);
  else return false;
  return true;
}
static bool mb_dw_small_bwd_weight(cudaStream_t st, const MbDwWgArgs& d, int c8, int sm_count) {
  // each CTA ends with 72 atomics per chunk: cap the grid so that a thread accumulates several utterances
  const int64_t gx = std::max<int64_t>(1, std::min<int64_t>(howl_ceil_div(d.B, DWS_THREADS), std::max<int64_t>(1, (int64_t)sm_count * 12 / c8)));
  const dim3 grid((unsigned)gx, c8);
  if (d.hin == 3 && d.win == 3 && d.stride == 1) mbn_dw_small_bwd_weight_kernel<3, 3, 1><<<grid, DWS_THREADS, 0, st>>>(d);
  else if (d.hin == 3 && d.win == 3 && d.stride == 2) mbn_dw_small_bwd_weight_kernel<3, 3, 2><<<grid, DWS_THREADS, 0, st>>>(d);
  else if (d.hin == 2 && d.win == 2 && d.stride == 1) mbn_dw_small_bwd_weight_kernel<2, 2, 1><<<grid, DWS_THREADS, 0, st>>>(d);
  else return false;
  return true;
}

// ---------------------------------------------------------------------------------------------
// BatchNorm (+ ReLU6) backward, two passes over (dY, raw):
//   dn = dY * [0 < n < 6] (act) with n = raw * scale + shift;  S1 = sum dn, S2 = sum dn * xhat
//   dRaw = gamma * rstd * (dn - S1 / count - xhat * S2 / count);  dgamma = S2, dbeta = S1
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) mbn_bn_bwd_stats_kernel(const uint4* __restrict__ dy, const uint4* __restrict__ raw,
                                                               const float* __restrict__ bn, int cp, int act, int64_t rows,
                                                               double* __restrict__ stats) {
  const int chunk = blockIdx.y, c8 = cp / 8;
  float sc[8], sh[8], mu[8], rs[8], part[16];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = chunk * 8 + j;
    sc[j] = bn[c]; sh[j] = bn[cp + c]; mu[j] = bn[2 * cp + c]; rs[j] = bn[3 * cp + c];
    part[j] = part[8 + j] = 0.f;
  }
  // two rows per thread and iteration: four 16-byte loads in flight before the first use (the kernel is register-limited to a few
  // warps per SM, so the bytes in flight per thread are what feeds HBM)
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += 2 * stride) {
    const int64_t r2 = r + stride;
    const bool two = r2 < rows;
    const size_t idx = mbn_vec(r, chunk, c8), idx2 = mbn_vec(two ? r2 : r, chunk, c8);
    const uint4 qg = dy[idx], qx = raw[idx], qg2 = dy[idx2], qx2 = raw[idx2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      if (h == 1 && !two) break;
      float g[8], x[8];
      mb_unpack(h ? qg2 : qg, g);
      mb_unpack(h ? qx2 : qx, x);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float dn = g[j];
        if (act) {
          const float n = fmaf(x[j], sc[j], sh[j]);
          if (!(n > 0.f && n < 6.f)) dn = 0.f;
        }
        part[j] += dn;
        part[8 + j] = fmaf(dn, (x[j] - mu[j]) * rs[j], part[8 + j]);
      }
    }
  }
  __shared__ float s_part[8][16];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const float s = warp_sum(part[i]);
    if (lane == 0) s_part[warp][i] = s;
  }
  __syncthreads();
  if (threadIdx.x < 16) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += (double)s_part[w][threadIdx.x];
    atomicAdd(stats + (size_t)(threadIdx.x >> 3) * cp + chunk * 8 + (threadIdx.x & 7), t);
  }
}

__global__ void __launch_bounds__(256) mbn_bn_bwd_apply_kernel(const uint4* __restrict__ dy, const uint4* __restrict__ raw,
                                                               const float* __restrict__ bn, int c, int cp, int act, int64_t rows,
                                                               int64_t rows_pad, const double* __restrict__ stats, double count,
                                                               uint4* __restrict__ draw, float* __restrict__ dgamma,
                                                               float* __restrict__ dbeta) {
  const int chunk = blockIdx.y, c8 = cp / 8;
  // dRaw = k0 (dn - m1 - (x - mu) rs m2) = k0 dn + ka x + kb
  float sc[8], sh[8], k0[8], ka[8], kb[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int ch = chunk * 8 + j;
    const float mu = bn[2 * cp + ch], rs = bn[3 * cp + ch];
    sc[j] = bn[ch]; sh[j] = bn[cp + ch];
    k0[j] = bn[4 * cp + ch] * rs;
    const float m1 = (float)(stats[ch] / count), m2 = (float)(stats[cp + ch] / count);
    ka[j] = -k0[j] * rs * m2;
    kb[j] = -k0[j] * m1 - ka[j] * mu;
    if (blockIdx.x == 0 && threadIdx.x == 0 && ch < c) {     // parameter gradients of the affine transform
      dgamma[ch] = (float)stats[cp + ch];
      dbeta[ch] = (float)stats[ch];
    }
  }
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < rows_pad; r += 2 * stride) {
    const int64_t r2 = r + stride;
    const bool two = r2 < rows_pad;
    const size_t idx = mbn_vec(r, chunk, c8), idx2 = mbn_vec(two ? r2 : r, chunk, c8);
    uint4 qg = make_uint4(0u, 0u, 0u, 0u), qx = qg, qg2 = qg, qx2 = qg;
    if (r < rows) { qg = dy[idx]; qx = raw[idx]; }
    if (two && r2 < rows) { qg2 = dy[idx2]; qx2 = raw[idx2]; }
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      if (h == 1 && !two) break;
      float g[8];
      if ((h ? r2 : r) < rows) {
        float x[8];
        mb_unpack(h ? qg2 : qg, g);
        mb_unpack(h ? qx2 : qx, x);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float dn = g[j];
          if (act) {
            const float n = fmaf(x[j], sc[j], sh[j]);
            if (!(n > 0.f && n < 6.f)) dn = 0.f;
          }
          g[j] = fmaf(k0[j], dn, fmaf(ka[j], x[j], kb[j]));
        }
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) g[j] = 0.f;
      }
      draw[h ? idx2 : idx] = mb_pack(g);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// stem: Conv2d(1, 3, 3, padding=(1, 3)) + bias -> BatchNorm(3) -> ReLU -> MaxPool(1, 2), feeding the 3x3 stride-2 entry convolution as
// an im2col matrix [B * ho * wo][27 -> 32] (TMO).  x is the frontend's [B, n_mels, frames] ("mels") layout.  One CTA per utterance;
// the stem's own activations never touch HBM (they are recomputed from the 13 KB feature map where needed).
// ---------------------------------------------------------------------------------------------
struct MbStemArgs {
  const float* x;        // [B, H, W]
  const float* w;        // [3][9]
  const float* bias;     // [3]
  const float* bn;       // [5][16] scale, shift, mean, rstd, gamma (after finalize)
  double* stats;         // [2][16]
  uint4* im2col;         // TMO [B * ho * wo][32]
  const uint4* dcol;     // backward: gradient of the im2col matrix
  double* bstats;        // backward: [2][16] S1, S2
  float* dw;             // backward: d conv weight [27]
  float* dbias;          // [3]
  int H, W, Wc, Wp, ho, wo;   // Wc = W + 4 conv columns, Wp = Wc / 2 pooled columns
  int64_t B;
};

// stage x with zero halo: s_x[(y + 1) * (W + 8) + (x + 3)]; returns the row pitch
__device__ __forceinline__ int mb_stem_stage(const MbStemArgs& a, int64_t b, float* s_x) {
  const int pitch = a.W + 8;
  for (int i = threadIdx.x; i < (a.H + 2) * pitch; i += blockDim.x) s_x[i] = 0.f;
  __syncthreads();
  const float* src = a.x + b * (int64_t)a.H * a.W;
  for (int i = threadIdx.x; i < a.H * a.W; i += blockDim.x) {
    const int y = i / a.W, x = i - y * a.W;
    s_x[(y + 1) * pitch + x + 3] = src[i];
  }
  __syncthreads();
  return pitch;
}
__device__ __forceinline__ float mb_stem_conv(const float* s_x, int pitch, const float* w, float bias, int y, int x) {
  // conv output (y, x) of the padded map: taps in rows y-1..y+1, columns x-3..x-1 -> staged at (y + ky) * pitch + (x + kx)
  float acc = bias;
#pragma unroll
  for (int ky = 0; ky < 3; ++ky)
#pragma unroll
    for (int kx = 0; kx < 3; ++kx) acc = fmaf(w[ky * 3 + kx], s_x[(y + ky) * pitch + x + kx], acc);
  return acc;
}

__global__ void __launch_bounds__(256) mbn_stem_stats_kernel(const MbStemArgs a) {
  extern __shared__ float smem[];
  const int pitch = mb_stem_stage(a, blockIdx.x, smem);
  float part[6] = {0, 0, 0, 0, 0, 0};
  for (int i = threadIdx.x; i < a.H * a.Wc; i += blockDim.x) {
    const int y = i / a.Wc, x = i - y * a.Wc;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float v = mb_stem_conv(smem, pitch, a.w + c * 9, a.bias[c], y, x);
      part[c] += v;
      part[3 + c] = fmaf(v, v, part[3 + c]);
    }
  }
  __shared__ float s_part[8][6];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    const float s = warp_sum(part[i]);
    if (lane == 0) s_part[warp][i] = s;
  }
  __syncthreads();
  if (threadIdx.x < 6) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += (double)s_part[w][threadIdx.x];
    atomicAdd(a.stats + (threadIdx.x / 3) * 16 + threadIdx.x % 3, t);
  }
}

// pooled stem activation s[c][y][xp] = max(relu(n[c][y][2 xp]), relu(n[c][y][2 xp + 1])), n = conv * scale + shift
__device__ __forceinline__ float mb_stem_act(const MbStemArgs& a, const float* s_x, int pitch, int c, int y, int xp) {
  const float sc = a.bn[c], sh = a.bn[16 + c];
  const float n0 = fmaf(mb_stem_conv(s_x, pitch, a.w + c * 9, a.bias[c], y, 2 * xp), sc, sh);
  const float n1 = fmaf(mb_stem_conv(s_x, pitch, a.w + c * 9, a.bias[c], y, 2 * xp + 1), sc, sh);
  return fmaxf(fmaxf(n0, 0.f), fmaxf(n1, 0.f));
}

__global__ void __launch_bounds__(256) mbn_stem_im2col_kernel(const MbStemArgs a) {
  extern __shared__ float smem[];
  const int64_t b = blockIdx.x;
  const int pitch = mb_stem_stage(a, b, smem);
  float* s_act = smem + (a.H + 2) * (a.W + 8);          // [3][H][Wp]
  for (int i = threadIdx.x; i < 3 * a.H * a.Wp; i += blockDim.x) {
    const int c = i / (a.H * a.Wp), rem = i - c * a.H * a.Wp, y = rem / a.Wp, xp = rem - y * a.Wp;
    s_act[i] = mb_stem_act(a, smem, pitch, c, y, xp);
  }
  __syncthreads();
  const int npix = a.ho * a.wo;
  for (int i = threadIdx.x; i < npix * 4; i += blockDim.x) {
    const int p = i >> 2, chunk = i & 3, yo = p / a.wo, xo = p - yo * a.wo;
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int col = chunk * 8 + j;
      float val = 0.f;
      if (col < 27) {
        const int c = col / 9, k = col - c * 9, y = 2 * yo - 1 + k / 3, x = 2 * xo - 1 + k % 3;
        if (y >= 0 && y < a.H && x >= 0 && x < a.Wp) val = s_act[(c * a.H + y) * a.Wp + x];
      }
      v[j] = val;
    }
    a.im2col[mbn_vec(b * npix + p, chunk, 4)] = mb_pack(v);
  }
}

// backward of the stem.  pass 0: S1 = sum dn, S2 = sum dn * xhat;  pass 1: dRaw -> conv weight / bias gradients
template <int PASS>
__global__ void __launch_bounds__(256) mbn_stem_bwd_kernel(const MbStemArgs a, double count) {
  extern __shared__ __align__(16) float smem[];
  const int64_t b = blockIdx.x;
  const int pitch = mb_stem_stage(a, b, smem);
  float* s_ds = smem + (a.H + 2) * (a.W + 8);           // [3][H][Wp] gradient w.r.t. the pooled activations (col2im, gathered)
  const int npix = a.ho * a.wo;
  // the utterance's rows of the im2col gradient, staged once with coalesced 128-bit loads (consecutive threads = consecutive rows of a
  // 128-row tile); the col2im gather below then reads single bf16 values out of shared memory instead of scattered 2-byte global loads
  uint4* s_dcol = reinterpret_cast<uint4*>(smem + (((a.H + 2) * (a.W + 8) + 3 * a.H * a.Wp + 3) & ~3));      // [npix][4 chunks], 16-byte aligned
  for (int i = threadIdx.x; i < npix * 4; i += blockDim.x) {
    const int chunk = i / npix, pix = i - chunk * npix;
    s_dcol[pix * 4 + chunk] = __ldg(a.dcol + mbn_vec(b * npix + pix, chunk, 4));
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 3 * a.H * a.Wp; i += blockDim.x) {
    const int c = i / (a.H * a.Wp), rem = i - c * a.H * a.Wp, y = rem / a.Wp, x = rem - y * a.Wp;
    float g = 0.f;
    for (int ky = 0; ky < 3; ++ky) {
      const int ty = y + 1 - ky;
      if (ty < 0 || (ty & 1) || ty / 2 >= a.ho) continue;
      for (int kx = 0; kx < 3; ++kx) {
        const int tx = x + 1 - kx;
        if (tx < 0 || (tx & 1) || tx / 2 >= a.wo) continue;
        const int col = c * 9 + ky * 3 + kx;
        const __nv_bfloat16* vec = reinterpret_cast<const __nv_bfloat16*>(s_dcol + ((ty / 2) * a.wo + tx / 2) * 4 + (col >> 3));
        g += __bfloat162float(vec[col & 7]);
      }
    }
    s_ds[i] = g;
  }
  __syncthreads();
  float part[30];
#pragma unroll
  for (int i = 0; i < 30; ++i) part[i] = 0.f;
  for (int i = threadIdx.x; i < a.H * a.Wc; i += blockDim.x) {
    const int y = i / a.Wc, x = i - y * a.Wc, xp = x >> 1;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float sc = a.bn[c], sh = a.bn[16 + c], mu = a.bn[32 + c], rs = a.bn[48 + c];
      const float raw = mb_stem_conv(smem, pitch, a.w + c * 9, a.bias[c], y, x);
      float dn = 0.f;
      if (xp < a.Wp) {       // the odd last conv column is dropped by the pooling
        const float n = fmaf(raw, sc, sh);
        const float other = fmaf(mb_stem_conv(smem, pitch, a.w + c * 9, a.bias[c], y, x ^ 1), sc, sh);
        const float r = fmaxf(n, 0.f), ro = fmaxf(other, 0.f);
        const bool win = (x & 1) ? (r > ro) : (r >= ro);      // torch's max pooling keeps the first maximum
        if (win && n > 0.f) dn = s_ds[(c * a.H + y) * a.Wp + xp];
      }
      const float xhat = (raw - mu) * rs;
      if (PASS == 0) {
        part[c] += dn;
        part[3 + c] = fmaf(dn, xhat, part[3 + c]);
      } else {
        const float m1 = (float)(a.bstats[c] / count), m2 = (float)(a.bstats[16 + c] / count);
        const float dr = a.bn[64 + c] * rs * (dn - m1 - xhat * m2);
        part[27 + c] += dr;
#pragma unroll
        for (int ky = 0; ky < 3; ++ky)
#pragma unroll
          for (int kx = 0; kx < 3; ++kx) part[c * 9 + ky * 3 + kx] = fmaf(dr, smem[(y + ky) * pitch + x + kx], part[c * 9 + ky * 3 + kx]);
      }
    }
  }
  __shared__ float s_part[8][30];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int n = PASS == 0 ? 6 : 30;
#pragma unroll
  for (int i = 0; i < 30; ++i) {
    if (i < n) {
      const float s = warp_sum(part[i]);
      if (lane == 0) s_part[warp][i] = s;
    }
  }
  __syncthreads();
  if (threadIdx.x < n) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += (double)s_part[w][threadIdx.x];
    if (PASS == 0) atomicAdd(a.bstats + (threadIdx.x / 3) * 16 + threadIdx.x % 3, t);
    else if (threadIdx.x < 27) atomicAdd(a.dw + threadIdx.x, (float)t);
    else atomicAdd(a.dbias + threadIdx.x - 27, (float)t);
  }
}

// ---------------------------------------------------------------------------------------------
// head: ReLU6(bn(last raw)) -> spatial mean -> Dropout(p) -> Linear(1280 -> L) -> CrossEntropy, and back
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ bool mb_keep(unsigned long long seed, int64_t b, int c, float p) {
  if (p <= 0.f) return true;
  unsigned long long z = seed + 0x9E3779B97F4A7C15ull * (unsigned long long)(b * MB_LAST + c + 1);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z ^= z >> 31;
  return (float)(z >> 40) * (1.f / 16777216.f) >= p;
}

// one CTA per utterance: pooled (fp32, with the dropout mask and 1 / (1 - p) applied) + logits
__global__ void __launch_bounds__(256) mbn_head_fwd_kernel(const uint4* __restrict__ raw, const float* __restrict__ bn, int hw,
                                                           const float* __restrict__ wc, const float* __restrict__ bc, int L, float drop_p,
                                                           unsigned long long seed, float* __restrict__ pooled, float* __restrict__ logits) {
  __shared__ float s_pool[MB_LAST];
  const int64_t b = blockIdx.x;
  const int cp = MB_LAST, c8 = cp / 8;
  for (int chunk = threadIdx.x; chunk < c8; chunk += blockDim.x) {
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int p = 0; p < hw; ++p) {
      float v[8];
      mb_unpack(raw[mbn_vec(b * hw + p, chunk, c8)], v);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += fminf(fmaxf(fmaf(v[j], bn[chunk * 8 + j], bn[cp + chunk * 8 + j]), 0.f), 6.f);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = chunk * 8 + j;
      float v = acc[j] / (float)hw;
      v = mb_keep(seed, b, c, drop_p) ? v / (1.f - drop_p) : 0.f;
      s_pool[c] = v;
      pooled[b * MB_LAST + c] = v;
    }
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int l = warp; l < L; l += 8) {
    float acc = 0.f;
    for (int c = lane; c < MB_LAST; c += 32) acc = fmaf(wc[(size_t)l * MB_LAST + c], s_pool[c], acc);
    acc = warp_sum(acc);
    if (lane == 0) logits[b * L + l] = acc + bc[l];
  }
}

// softmax cross-entropy: dlogits = (p - onehot) / batch (or the caller's dlogits), loss
__global__ void mbn_ce_kernel(const float* __restrict__ logits, const int64_t* __restrict__ labels, const float* __restrict__ dlogits_in,
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

// classifier gradients: dWc[l][c] = sum_b dlogits[b][l] pooled[b][c]; dbc[l] = sum_b dlogits[b][l].   grid = (L, ceil(1280 / 256))
__global__ void __launch_bounds__(256) mbn_cls_wgrad_kernel(const float* __restrict__ dlogits, const float* __restrict__ pooled, int64_t B, int L,
                                                            float* __restrict__ dwc, float* __restrict__ dbc, const double* __restrict__ loss_acc,
                                                            float* __restrict__ loss) {
  // grid = (L, 1280 / 256, batch slices): partial sums over a slice of the batch, accumulated into the (zeroed) gradient
  const int l = blockIdx.x, c = blockIdx.y * 256 + threadIdx.x;
  const int64_t per = (B + gridDim.z - 1) / gridDim.z;
  const int64_t lo = (int64_t)blockIdx.z * per, hi = lo + per < B ? lo + per : B;
  float acc = 0.f, accb = 0.f;
  for (int64_t b = lo; b < hi; ++b) {
    const float d = dlogits[b * L + l];
    acc = fmaf(d, pooled[b * MB_LAST + c], acc);
    accb += d;
  }
  atomicAdd(dwc + (size_t)l * MB_LAST + c, acc);
  if (c == 0) atomicAdd(dbc + l, accb);
  if (l == 0 && c == 0 && blockIdx.z == 0 && loss) *loss = (float)(*loss_acc);
}

// gradient at the last activation: dY[b, p, c] = keep * dPooled[b][c] / ((1 - p) hw), dPooled = dlogits . Wc      (TMO bf16)
__global__ void __launch_bounds__(256) mbn_head_bwd_kernel(const float* __restrict__ dlogits, const float* __restrict__ wc, int L, int hw,
                                                           float drop_p, unsigned long long seed, uint4* __restrict__ dy, int64_t B) {
  __shared__ float s_d[96];
  const int64_t b = blockIdx.x;
  if (threadIdx.x < L) s_d[threadIdx.x] = dlogits[b * L + threadIdx.x];
  __syncthreads();
  const int c8 = MB_LAST / 8;
  for (int chunk = threadIdx.x; chunk < c8; chunk += blockDim.x) {
    float g[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = chunk * 8 + j;
      float acc = 0.f;
      for (int l = 0; l < L; ++l) acc = fmaf(s_d[l], wc[(size_t)l * MB_LAST + c], acc);
      g[j] = mb_keep(seed, b, c, drop_p) ? acc / ((1.f - drop_p) * (float)hw) : 0.f;
    }
    const uint4 q = mb_pack(g);
    for (int p = 0; p < hw; ++p) dy[mbn_vec(b * hw + p, chunk, c8)] = q;
  }
}

__global__ void mbn_zero_pad_rows_kernel(uint4* __restrict__ t, int64_t rows, int64_t rows_pad, int c8) {
  const int64_t n = (rows_pad - rows) * c8;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = rows + i / c8;
    t[mbn_vec(r, (int)(i % c8), c8)] = make_uint4(0, 0, 0, 0);
  }
}

// =============================================================================================
// workspace
// =============================================================================================
struct MbWs {
  std::vector<__nv_bfloat16*> raw;      // raw output of conv i (TMO)
  std::vector<__nv_bfloat16*> act;      // materialised activation after conv i (projection outputs, depthwise outputs), or null
  std::vector<float*> bn;               // [5][cp] per conv
  std::vector<double*> stats;           // [2][cp] forward statistics per conv
  std::vector<__nv_bfloat16*> wop;      // forward weight operand per conv (GEMM convs)
  std::vector<__nv_bfloat16*> wopT;     // data-gradient weight operand
  __nv_bfloat16* im2col;
  __nv_bfloat16* g[3];                  // gradient ping-pong buffers (largest tensor)
  double* bstats;                       // [2][MB_MAXC] backward statistics (reused)
  double* loss_acc;
  float* pooled;
  float* dlogits;
  float* logits;
  size_t bytes;
};

static MbWs mb_carve(const MbNet& net, void* base, int64_t B, int L) {
  MbWs w;
  size_t off = 0;
  char* p = (char*)base;
  auto take = [&](size_t bytes) {
    void* r = p ? (void*)(p + off) : nullptr;
    off += howl_align_up(bytes, 256);
    return r;
  };
  size_t gmax = 0;
  const size_t n = net.convs.size();
  w.raw.resize(n); w.act.resize(n); w.bn.resize(n); w.stats.resize(n); w.wop.resize(n); w.wopT.resize(n);
  for (size_t i = 0; i < n; ++i) {
    const MbConv& c = net.convs[i];
    const int64_t rows = B * c.hout * c.wout, rows_in = B * c.hin * c.win;
    const int cp = mbn_pad16(c.cout);
    w.raw[i] = (c.kind == 0) ? nullptr : (__nv_bfloat16*)take(mbn_tmo_bytes(rows, c.cout));
    // materialised activations: depthwise outputs (GEMM operand of the projection) and block outputs (projection + BN [+ skip])
    const bool mat = (c.kind == 3) || (c.kind == 2 && c.act == 0);
    w.act[i] = mat ? (__nv_bfloat16*)take(mbn_tmo_bytes(rows, c.cout)) : nullptr;
    w.bn[i] = (float*)take(sizeof(float) * 5 * cp);
    w.stats[i] = (double*)take(sizeof(double) * 2 * cp);
    const bool gemm = c.kind == 1 || c.kind == 2;
    const int k = c.kind == 1 ? 27 : c.cin;
    w.wop[i] = gemm ? (__nv_bfloat16*)take(mbn_weight_operand_bytes(c.cout, k)) : nullptr;
    w.wopT[i] = gemm ? (__nv_bfloat16*)take(mbn_weight_operand_bytes(k, c.cout)) : nullptr;
    if (c.kind != 0) {
      gmax = std::max(gmax, mbn_tmo_bytes(rows, c.cout));
      gmax = std::max(gmax, mbn_tmo_bytes(rows_in, c.kind == 1 ? 32 : c.cin));
    }
  }
  const MbConv& e = net.convs[1];
  w.im2col = (__nv_bfloat16*)take(mbn_tmo_bytes(B * e.hout * e.wout, 32));
  for (int i = 0; i < 3; ++i) w.g[i] = (__nv_bfloat16*)take(gmax);
  w.bstats = (double*)take(sizeof(double) * 2 * MB_MAXC);
  w.loss_acc = (double*)take(sizeof(double) * 2);
  w.pooled = (float*)take(sizeof(float) * B * MB_LAST);
  w.dlogits = (float*)take(sizeof(float) * B * L);
  w.logits = (float*)take(sizeof(float) * B * L);
  w.bytes = off;
  return w;
}

// grid.x of the depthwise kernels: groups of `ni` utterances, capped at ~6 CTAs per SM over all chunks
static unsigned mb_dw_blocks(howl_ctx_t* ctx, int64_t B, int ni, int chunks) {
  const int64_t groups = howl_ceil_div(B, ni);
  const int64_t cap = std::max<int64_t>(1, (int64_t)ctx->sm_count * 6 / chunks);
  return (unsigned)std::max<int64_t>(1, std::min(groups, cap));
}

static unsigned mb_rowblocks(howl_ctx_t* ctx, int64_t rows, int chunks) {
  int64_t bx = howl_ceil_div(rows, 256);
  const int64_t cap = std::max<int64_t>(1, (int64_t)ctx->sm_count * 8 / chunks);
  return (unsigned)std::max<int64_t>(1, std::min(bx, cap));
}

// =============================================================================================
// C ABI
// =============================================================================================
extern "C" int64_t howl_b200_mobilenet_param_count(int32_t num_labels) {
  if (num_labels < 1) return -1;
  return (int64_t)mb_build(40, 81, num_labels).n_params;
}
extern "C" int64_t howl_b200_mobilenet_bn_channels(void) { return mb_build(40, 81, 1).n_bn_channels; }
extern "C" int64_t howl_b200_mobilenet_bn_layers(void) { return mb_build(40, 81, 1).n_bn; }

extern "C" int64_t howl_b200_mobilenet_workspace_bytes(int64_t B, int32_t frames, int32_t n_mels, int32_t num_labels) {
  if (B < 1 || frames < 8 || n_mels < 8 || num_labels < 1 || num_labels > 96) return -1;
  const MbNet net = mb_build(n_mels, frames, num_labels);
  return (int64_t)mb_carve(net, nullptr, B, num_labels).bytes;
}

static int mb_check(howl_ctx_t* ctx, int64_t B, int frames, int n_mels, int L, const void* ws, size_t ws_bytes, MbNet* net, MbWs* out) {
  HOWL_REQUIRE(ctx, B >= 1 && frames >= 8 && n_mels >= 8, HOWL_E_INVALID, "mobilenet: bad shape B=%lld frames=%d mels=%d", (long long)B, frames, n_mels);
  HOWL_REQUIRE(ctx, L >= 1 && L <= 96, HOWL_E_UNSUPPORTED, "mobilenet: num_labels=%d outside 1..96", L);
  HOWL_REQUIRE(ctx, ws != nullptr, HOWL_E_WORKSPACE, "mobilenet: null workspace");
  *net = mb_build(n_mels, frames, L);
  *out = mb_carve(*net, const_cast<void*>(ws), B, L);
  HOWL_REQUIRE(ctx, out->bytes <= ws_bytes, HOWL_E_WORKSPACE, "mobilenet: workspace %zu < required %zu", ws_bytes, out->bytes);
  HOWL_REQUIRE(ctx, B * (int64_t)n_mels * ((frames + 4) / 2) < ((int64_t)1 << 30), HOWL_E_UNSUPPORTED,
               "mobilenet: batch of %lld clips exceeds the 32-bit row indices of the stencil kernels", (long long)B);
  const size_t stem_smem = sizeof(float) * ((size_t)(n_mels + 2) * (frames + 8) + 3 * (size_t)n_mels * ((frames + 4) / 2));
  HOWL_REQUIRE(ctx, stem_smem <= 200 * 1024, HOWL_E_UNSUPPORTED, "mobilenet: clip of %d frames exceeds the stem's shared-memory tile", frames);
  return HOWL_OK;
}

static MbStemArgs mb_stem_args(const MbNet& net, const MbWs& ws, const float* feats, const float* params, int64_t B) {
  const MbConv& s = net.convs[0];
  const MbConv& e = net.convs[1];
  MbStemArgs a;
  memset(&a, 0, sizeof(a));
  a.x = feats; a.w = params + s.w_off; a.bias = params + s.bias_off; a.bn = ws.bn[0]; a.stats = ws.stats[0];
  a.im2col = reinterpret_cast<uint4*>(ws.im2col);
  a.H = s.hin; a.W = s.win; a.Wc = s.win + 4; a.Wp = s.wout; a.ho = e.hout; a.wo = e.wout; a.B = B;
  return a;
}

extern "C" int howl_b200_mobilenet_fwd(howl_ctx_t* ctx, void* stream, const float* feats, int64_t B, int32_t frames, int32_t n_mels,
                                       int32_t num_labels, const float* params, float* bn_running, int64_t* num_batches_tracked, int train,
                                       float dropout_p, uint64_t seed, float* logits, void* workspace, size_t workspace_bytes) {
  if (!ctx) return HOWL_E_INVALID;
  HOWL_REQUIRE(ctx, feats && params && bn_running && logits, HOWL_E_INVALID, "mobilenet_fwd: null pointer");
  HOWL_REQUIRE(ctx, dropout_p >= 0.f && dropout_p < 1.f, HOWL_E_INVALID, "mobilenet_fwd: dropout_p outside [0, 1)");
  MbNet net;
  MbWs ws;
  int rc = mb_check(ctx, B, frames, n_mels, num_labels, workspace, workspace_bytes, &net, &ws);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const int L = num_labels, NB = net.n_bn_channels;
  const size_t n = net.convs.size();
  float* run_mean = bn_running;
  float* run_var = bn_running + NB;
  auto finalize = [&](size_t i, double count) {
    const MbConv& c = net.convs[i];
    const int cp = mbn_pad16(c.cout);
    mbn_bn_finalize_kernel<<<(cp + 127) / 128, 128, 0, st>>>(ws.stats[i], count, c.cout, cp, params + c.g_off, params + c.b_off,
                                                             run_mean + c.bn_off, run_var + c.bn_off,
                                                             num_batches_tracked ? num_batches_tracked + c.bn_index : nullptr, train, ws.bn[i]);
    HOWL_LAUNCHED(ctx, "mbn_bn_finalize");
    return HOWL_OK;
  };
  if (train) {
    for (size_t i = 0; i < n; ++i) HOWL_CUDA(ctx, cudaMemsetAsync(ws.stats[i], 0, sizeof(double) * 2 * mbn_pad16(net.convs[i].cout), st));
  }
  // ---- weight operands (bf16) from the fp32 master weights
  {
    std::vector<MbnWopDesc> descs;
    for (size_t i = 1; i < n; ++i) {
      const MbConv& c = net.convs[i];
      if (c.kind == 1) descs.push_back(MbnWopDesc{params + c.w_off, ws.wop[i], c.cout, 27, 27, 0, 0});
      else if (c.kind == 2) descs.push_back(MbnWopDesc{params + c.w_off, ws.wop[i], c.cout, c.cin, c.cin, 0, 0});
    }
    if ((rc = mbn_weight_operand_batch(ctx, st, descs.data(), (int)descs.size()))) return rc;
  }
  // ---- stem (+ the entry convolution's im2col matrix)
  {
    MbStemArgs a = mb_stem_args(net, ws, feats, params, B);
    const size_t sm = sizeof(float) * ((size_t)(a.H + 2) * (a.W + 8) + 3 * (size_t)a.H * a.Wp);
    HOWL_CUDA(ctx, cudaFuncSetAttribute(mbn_stem_stats_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    HOWL_CUDA(ctx, cudaFuncSetAttribute(mbn_stem_im2col_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    if (train) {
      mbn_stem_stats_kernel<<<(unsigned)B, 256, sm, st>>>(a);
      HOWL_LAUNCHED(ctx, "mbn_stem_stats");
    }
    rc = finalize(0, (double)B * a.H * a.Wc);
    if (rc) return rc;
    mbn_stem_im2col_kernel<<<(unsigned)B, 256, sm, st>>>(a);
    HOWL_LAUNCHED(ctx, "mbn_stem_im2col");
    const int64_t rows = B * a.ho * a.wo, rows_pad = mbn_tiles(rows) * MBN_TILE;
    if (rows_pad > rows) {
      mbn_zero_pad_rows_kernel<<<8, 256, 0, st>>>(reinterpret_cast<uint4*>(ws.im2col), rows, rows_pad, 4);
      HOWL_LAUNCHED(ctx, "mbn_zero_pad");
    }
  }
  // ---- the convolution stack
  const __nv_bfloat16* x = ws.im2col;         // current GEMM input (materialised)
  const __nv_bfloat16* block_in = nullptr;    // input of the current block (skip connection)
  for (size_t i = 1; i < n; ++i) {
    const MbConv& c = net.convs[i];
    const int64_t rows = B * c.hout * c.wout, rows_pad = mbn_tiles(rows) * MBN_TILE;
    const int cp = mbn_pad16(c.cout), c8 = cp / 8;
    if (c.kind == 3) {
      // depthwise: reads the raw output of the previous conv with its BatchNorm + ReLU6 applied on the fly
      MbDwArgs d;
      d.in = reinterpret_cast<const uint4*>(ws.raw[i - 1]); d.bn_in = ws.bn[i - 1]; d.w = params + c.w_off;
      d.out = reinterpret_cast<uint4*>(ws.raw[i]); d.stats = train ? ws.stats[i] : nullptr;
      d.B = B; d.c = c.cout; d.cp = cp; d.hin = c.hin; d.win = c.win; d.hout = c.hout; d.wout = c.wout; d.stride = c.stride;
      HOWL_REQUIRE(ctx, mb_dw_fits(c.hin, c.win), HOWL_E_UNSUPPORTED, "mobilenet: a %d x %d depthwise input exceeds the shared-memory image tile", c.hin, c.win);
      d.g = mb_dw_geom(B, c.hin, c.win, c.hout, c.wout);
      if (!mb_dw_small_fwd(st, d, c8)) mbn_dw_fwd_kernel<<<dim3(mb_dw_blocks(ctx, B, d.g.ni, c8), c8), 256, 0, st>>>(d);
      HOWL_LAUNCHED(ctx, "mbn_dw_fwd");
    } else {
      if (c.kind == 2 && c.block >= 0 && net.convs[i - 1].block != c.block) block_in = x;   // first conv of a block: its input is the skip
      rc = mbn_gemm_nt(ctx, st, x, ws.wop[i], nullptr, ws.raw[i], rows, c.kind == 1 ? 27 : c.cin, c.cout);
      if (rc) return rc;
      if (train) {
        mbn_stats_kernel<<<dim3(mb_rowblocks(ctx, rows, c8), c8), 256, 0, st>>>(reinterpret_cast<const uint4*>(ws.raw[i]), rows, c8, ws.stats[i], cp);
        HOWL_LAUNCHED(ctx, "mbn_stats");
      }
    }
    rc = finalize(i, (double)rows);
    if (rc) return rc;
    if (ws.act[i]) {
      const __nv_bfloat16* res = c.residual ? block_in : nullptr;
      mbn_apply_kernel<<<dim3(mb_rowblocks(ctx, rows_pad, c8), c8), 256, 0, st>>>(reinterpret_cast<const uint4*>(ws.raw[i]), ws.bn[i], cp, c.act,
                                                                                reinterpret_cast<const uint4*>(res),
                                                                                reinterpret_cast<uint4*>(ws.act[i]), rows, rows_pad);
      HOWL_LAUNCHED(ctx, "mbn_apply");
      x = ws.act[i];
    }
  }
  const MbConv& last = net.convs[n - 1];
  mbn_head_fwd_kernel<<<(unsigned)B, 256, 0, st>>>(reinterpret_cast<const uint4*>(ws.raw[n - 1]), ws.bn[n - 1], last.hout * last.wout,
                                                  params + net.cls_w, params + net.cls_b, L, train ? dropout_p : 0.f, seed, ws.pooled, ws.logits);
  HOWL_LAUNCHED(ctx, "mbn_head_fwd");
  HOWL_CUDA(ctx, cudaMemcpyAsync(logits, ws.logits, sizeof(float) * B * L, cudaMemcpyDeviceToDevice, st));
  return HOWL_OK;
}

// stem BatchNorm parameter gradients from the backward statistics (3 channels)
__global__ void mbn_stem_bn_grads_kernel(const double* __restrict__ bstats, float* __restrict__ dgamma, float* __restrict__ dbeta) {
  if (threadIdx.x < 3) {
    dgamma[threadIdx.x] = (float)bstats[16 + threadIdx.x];
    dbeta[threadIdx.x] = (float)bstats[threadIdx.x];
  }
}

static int mb_bwd_impl(howl_ctx_t* ctx, void* stream, const float* feats, const int64_t* labels, const float* dlogits_in, int64_t B,
                       int32_t frames, int32_t n_mels, int32_t num_labels, int64_t loss_scale_batch, const float* params, float* grads,
                       float dropout_p, uint64_t seed, float* loss, void* workspace, size_t workspace_bytes) {
  MbNet net;
  MbWs ws;
  int rc = mb_check(ctx, B, frames, n_mels, num_labels, workspace, workspace_bytes, &net, &ws);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const int L = num_labels;
  const size_t n = net.convs.size();
  HOWL_CUDA(ctx, cudaMemsetAsync(grads, 0, sizeof(float) * net.n_params, st));
  HOWL_CUDA(ctx, cudaMemsetAsync(ws.loss_acc, 0, sizeof(double) * 2, st));
  // data-gradient weight operands (W^T) of the GEMM convolutions
  {
    std::vector<MbnWopDesc> descs;
    for (size_t i = 1; i < n; ++i) {
      const MbConv& c = net.convs[i];
      if (c.kind == 1) descs.push_back(MbnWopDesc{params + c.w_off, ws.wopT[i], 27, c.cout, 27, 1, 0});
      else if (c.kind == 2) descs.push_back(MbnWopDesc{params + c.w_off, ws.wopT[i], c.cin, c.cout, c.cin, 1, 0});
    }
    if ((rc = mbn_weight_operand_batch(ctx, st, descs.data(), (int)descs.size()))) return rc;
  }
  // ---- head
  mbn_ce_kernel<<<(unsigned)howl_ceil_div(B, 128), 128, 0, st>>>(ws.logits, labels, dlogits_in, ws.dlogits, ws.loss_acc, B, L,
                                                                 1.f / (float)loss_scale_batch);
  HOWL_LAUNCHED(ctx, "mbn_ce");
  mbn_cls_wgrad_kernel<<<dim3(L, MB_LAST / 256, (unsigned)std::max<int64_t>(1, std::min<int64_t>(64, B / 64))), 256, 0, st>>>(ws.dlogits, ws.pooled, B, L, grads + net.cls_w, grads + net.cls_b, ws.loss_acc, loss);
  HOWL_LAUNCHED(ctx, "mbn_cls_wgrad");
  const MbConv& last = net.convs[n - 1];
  // Three gradient buffers rotate: `gy` = dL/d(activation after conv i's BatchNorm), `gr` = dL/d(raw output of conv i), and the
  // gradient handed to conv i - 1.  `held` keeps the output gradient of a block with a skip connection until the block's first
  // convolution adds it to its data gradient.
  __nv_bfloat16* gy = ws.g[0];
  __nv_bfloat16* held = nullptr;
  auto pick = [&](const __nv_bfloat16* x0, const __nv_bfloat16* x1, const __nv_bfloat16* x2) -> __nv_bfloat16* {
    for (int k = 0; k < 3; ++k)
      if (ws.g[k] != x0 && ws.g[k] != x1 && ws.g[k] != x2) return ws.g[k];
    return nullptr;
  };
  mbn_head_bwd_kernel<<<(unsigned)B, 256, 0, st>>>(ws.dlogits, params + net.cls_w, L, last.hout * last.wout, dropout_p, seed,
                                                  reinterpret_cast<uint4*>(gy), B);
  HOWL_LAUNCHED(ctx, "mbn_head_bwd");
  bool stats_ready = false;       // the depthwise data gradient of conv i + 1 already left conv i's BatchNorm-backward statistics in ws.bstats
  for (size_t i = n - 1; i >= 1; --i) {
    const MbConv& c = net.convs[i];
    const int64_t rows = B * c.hout * c.wout, rows_pad = mbn_tiles(rows) * MBN_TILE;
    const int64_t rows_in = B * c.hin * c.win, rows_in_pad = mbn_tiles(rows_in) * MBN_TILE;
    const int cp = mbn_pad16(c.cout), c8 = cp / 8;
    const dim3 grid(mb_rowblocks(ctx, rows_pad, c8), c8);
    // ---- BatchNorm (+ ReLU6) backward of conv i: gy -> gr (+ dgamma, dbeta)
    __nv_bfloat16* gr = pick(gy, held, nullptr);
    HOWL_REQUIRE(ctx, gr != nullptr, HOWL_E_INVALID, "mobilenet_bwd: gradient buffer rotation");
    if (!stats_ready) {
      HOWL_CUDA(ctx, cudaMemsetAsync(ws.bstats, 0, sizeof(double) * 2 * MB_MAXC, st));
      mbn_bn_bwd_stats_kernel<<<grid, 256, 0, st>>>(reinterpret_cast<const uint4*>(gy), reinterpret_cast<const uint4*>(ws.raw[i]), ws.bn[i], cp,
                                                    c.act, rows, ws.bstats);
      HOWL_LAUNCHED(ctx, "mbn_bn_bwd_stats");
    }
    stats_ready = false;
    mbn_bn_bwd_apply_kernel<<<grid, 256, 0, st>>>(reinterpret_cast<const uint4*>(gy), reinterpret_cast<const uint4*>(ws.raw[i]), ws.bn[i], c.cout, cp,
                                                  c.act, rows, rows_pad, ws.bstats, (double)rows, reinterpret_cast<uint4*>(gr),
                                                  grads + c.g_off, grads + c.b_off);
    HOWL_LAUNCHED(ctx, "mbn_bn_bwd_apply");
    if (c.residual) held = gy;        // the block's output gradient also reaches the block input through the skip connection
    __nv_bfloat16* dst = (gy != held) ? gy : pick(gy, gr, held);
    HOWL_REQUIRE(ctx, dst != nullptr, HOWL_E_INVALID, "mobilenet_bwd: gradient buffer rotation");
    if (c.kind == 3) {
      // depthwise: weight gradient (the activated input is recomputed from the producer's raw output), then data gradient
      MbDwWgArgs wg;
      wg.dout = reinterpret_cast<const uint4*>(gr); wg.in = reinterpret_cast<const uint4*>(ws.raw[i - 1]); wg.bn_in = ws.bn[i - 1];
      wg.dw = grads + c.w_off; wg.B = B; wg.c = c.cout; wg.cp = cp; wg.hin = c.hin; wg.win = c.win; wg.hout = c.hout; wg.wout = c.wout;
      wg.stride = c.stride;
      wg.g = mb_dw_geom(B, c.hin, c.win, c.hout, c.wout);
      if (!mb_dw_small_bwd_weight(st, wg, c8, ctx->sm_count))
        mbn_dw_bwd_weight_kernel<<<dim3(mb_dw_blocks(ctx, B, wg.g.ni, c8), c8), 256, 0, st>>>(wg);
      HOWL_LAUNCHED(ctx, "mbn_dw_bwd_weight");
      MbDwBwdArgs d;
      d.dout = reinterpret_cast<const uint4*>(gr); d.w = params + c.w_off; d.din = reinterpret_cast<uint4*>(dst);
      d.B = B; d.c = c.cout; d.cp = cp; d.hin = c.hin; d.win = c.win; d.hout = c.hout; d.wout = c.wout; d.stride = c.stride;
      d.g = mb_dw_geom(B, c.hout, c.wout, c.hin, c.win);
      d.raw_prev = nullptr; d.bn_prev = nullptr; d.bstats = nullptr; d.act_prev = 0;
      if (!mb_dw_small_bwd_data(st, d, c8)) {
        if (i >= 2) {       // the producer is a GEMM convolution with its own BatchNorm: fuse its backward statistics (same channel count)
          HOWL_CUDA(ctx, cudaMemsetAsync(ws.bstats, 0, sizeof(double) * 2 * MB_MAXC, st));      // free again: conv i's apply pass has run
          d.raw_prev = reinterpret_cast<const uint4*>(ws.raw[i - 1]); d.bn_prev = ws.bn[i - 1]; d.bstats = ws.bstats;
          d.act_prev = net.convs[i - 1].act;
          stats_ready = true;
        }
        mbn_dw_bwd_data_kernel<<<dim3(mb_dw_blocks(ctx, B, d.g.ni, c8), c8), 256, 0, st>>>(d);
      }
      HOWL_LAUNCHED(ctx, "mbn_dw_bwd_data");
    } else {
      // GEMM convolution: dW[cout][k] += gr^T * input;  d(input) = gr * W (+ the skip gradient at a block's first convolution)
      const int k_in = c.kind == 1 ? 27 : c.cin;
      const __nv_bfloat16* input = c.kind == 1 ? ws.im2col : ws.act[i - 1];
      HOWL_REQUIRE(ctx, input != nullptr, HOWL_E_INVALID, "mobilenet_bwd: conv %d has no materialised input", (int)i);
      rc = mbn_gemm_wgrad(ctx, st, gr, input, grads + c.w_off, rows, c.cout, k_in, c.cout, k_in, k_in);
      if (rc) return rc;
      const bool first_of_block = c.block >= 0 && net.convs[i - 1].block != c.block;
      const __nv_bfloat16* add = (first_of_block && held) ? held : nullptr;
      rc = mbn_gemm_nt(ctx, st, gr, ws.wopT[i], add, dst, rows, c.cout, k_in);
      if (rc) return rc;
      if (add) held = nullptr;
    }
    gy = dst;
  }
  // ---- stem: gy = gradient of the im2col matrix
  {
    MbStemArgs a = mb_stem_args(net, ws, feats, params, B);
    a.dcol = reinterpret_cast<const uint4*>(gy);
    a.bstats = ws.bstats;
    a.dw = grads + net.convs[0].w_off;
    a.dbias = grads + net.convs[0].bias_off;
    const size_t sm = sizeof(float) * ((size_t)(a.H + 2) * (a.W + 8) + 3 * (size_t)a.H * a.Wp + 4) + 64 * (size_t)a.ho * a.wo;   // x tile, ds, dcol tile
    HOWL_REQUIRE(ctx, sm <= 200 * 1024, HOWL_E_UNSUPPORTED, "mobilenet_bwd: %d frames do not fit the stem's shared-memory tiles", a.W);
    HOWL_CUDA(ctx, cudaFuncSetAttribute(mbn_stem_bwd_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    HOWL_CUDA(ctx, cudaFuncSetAttribute(mbn_stem_bwd_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    HOWL_CUDA(ctx, cudaMemsetAsync(ws.bstats, 0, sizeof(double) * 2 * MB_MAXC, st));
    const double count = (double)B * a.H * a.Wc;
    mbn_stem_bwd_kernel<0><<<(unsigned)B, 256, sm, st>>>(a, count);
    HOWL_LAUNCHED(ctx, "mbn_stem_bwd_stats");
    mbn_stem_bwd_kernel<1><<<(unsigned)B, 256, sm, st>>>(a, count);
    HOWL_LAUNCHED(ctx, "mbn_stem_bwd_apply");
  }
  return HOWL_OK;
}

extern "C" int howl_b200_mobilenet_bwd(howl_ctx_t* ctx, void* stream, const float* feats, const int64_t* labels, int64_t B, int32_t frames,
                                       int32_t n_mels, int32_t num_labels, int64_t loss_scale_batch, const float* params, float* grads,
                                       float dropout_p, uint64_t seed, float* loss, void* workspace, size_t workspace_bytes) {
  if (!ctx) return HOWL_E_INVALID;
  HOWL_REQUIRE(ctx, feats && labels && params && grads && loss, HOWL_E_INVALID, "mobilenet_bwd: null pointer");
  HOWL_REQUIRE(ctx, loss_scale_batch >= 1, HOWL_E_INVALID, "mobilenet_bwd: loss_scale_batch must be >= 1");
  int rc = mb_bwd_impl(ctx, stream, feats, labels, nullptr, B, frames, n_mels, num_labels, loss_scale_batch, params, grads, dropout_p, seed,
                       loss, workspace, workspace_bytes);
  if (rc) return rc;
  MbNet net = mb_build(n_mels, frames, num_labels);
  MbWs ws = mb_carve(net, workspace, B, num_labels);
  mbn_stem_bn_grads_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(ws.bstats, grads + net.convs[0].g_off, grads + net.convs[0].b_off);
  HOWL_LAUNCHED(ctx, "mbn_stem_bn_grads");
  return HOWL_OK;
}

extern "C" int howl_b200_mobilenet_bwd_dlogits(howl_ctx_t* ctx, void* stream, const float* feats, const float* dlogits, int64_t B,
                                               int32_t frames, int32_t n_mels, int32_t num_labels, const float* params, float* grads,
                                               float dropout_p, uint64_t seed, void* workspace, size_t workspace_bytes) {
  if (!ctx) return HOWL_E_INVALID;
  HOWL_REQUIRE(ctx, feats && dlogits && params && grads, HOWL_E_INVALID, "mobilenet_bwd_dlogits: null pointer");
  int rc = mb_bwd_impl(ctx, stream, feats, nullptr, dlogits, B, frames, n_mels, num_labels, 1, params, grads, dropout_p, seed, nullptr,
                       workspace, workspace_bytes);
  if (rc) return rc;
  MbNet net = mb_build(n_mels, frames, num_labels);
  MbWs ws = mb_carve(net, workspace, B, num_labels);
  mbn_stem_bn_grads_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(ws.bstats, grads + net.convs[0].g_off, grads + net.convs[0].b_off);
  HOWL_LAUNCHED(ctx, "mbn_stem_bn_grads");
  return HOWL_OK;
}

extern "C" int howl_b200_mobilenet_train_step(howl_ctx_t* ctx, void* stream, const float* pcm, const int64_t* labels, int64_t B, int64_t T,
                                              const float* fb, float zmuv_mean, float zmuv_std, int32_t num_labels, float* params,
                                              float* bn_running, int64_t* num_batches_tracked, float* grads, float* exp_avg, float* exp_avg_sq,
                                              int64_t step, float lr, float weight_decay, float dropout_p, uint64_t seed, float* loss,
                                              float* logits, void* workspace, size_t workspace_bytes) {
  if (!ctx) return HOWL_E_INVALID;
  HOWL_REQUIRE(ctx, pcm && labels && fb && params && grads && exp_avg && exp_avg_sq && loss && logits, HOWL_E_INVALID,
               "mobilenet_train_step: null pointer");
  const int M = ctx->fe.n_mels;
  const int64_t F = howl_b200_num_frames(T, ctx->fe.hop);
  const size_t feat_bytes = howl_align_up((size_t)B * F * M * sizeof(float), 256);
  HOWL_REQUIRE(ctx, workspace && workspace_bytes > feat_bytes, HOWL_E_WORKSPACE, "mobilenet_train_step: workspace too small");
  float* feats = (float*)workspace;
  void* ws = (char*)workspace + feat_bytes;
  int rc = howl_b200_frontend_fwd(ctx, stream, pcm, B, T, fb, zmuv_mean, zmuv_std, nullptr, HOWL_FE_MELS_ONLY | HOWL_FE_ZMUV, feats);
  if (rc) return rc;
  rc = howl_b200_mobilenet_fwd(ctx, stream, feats, B, (int32_t)F, M, num_labels, params, bn_running, num_batches_tracked, 1, dropout_p, seed,
                               logits, ws, workspace_bytes - feat_bytes);
  if (rc) return rc;
  rc = howl_b200_mobilenet_bwd(ctx, stream, feats, labels, B, (int32_t)F, M, num_labels, B, params, grads, dropout_p, seed, loss, ws,
                               workspace_bytes - feat_bytes);
  if (rc) return rc;
  const int64_t np = howl_b200_mobilenet_param_count(num_labels);
  return howl_b200_adamw(ctx, stream, params, grads, exp_avg, exp_avg_sq, np, step, lr, 0.9f, 0.999f, 1e-8f, weight_decay);
}

// ---------------------------------------------------------------------------------------------
// test hook (include/howl_b200_debug.h): the activation decisions the backward takes, for the mask-forced gradient oracle
// ---------------------------------------------------------------------------------------------
__global__ void mbn_debug_mask_kernel(const uint4* __restrict__ raw, const float* __restrict__ bn, int c, int cp, int64_t rows, uint8_t* __restrict__ out) {
  const int64_t n = rows * c;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = i / c;
    const int ch = (int)(i - row * c);
    const __nv_bfloat16* v = reinterpret_cast<const __nv_bfloat16*>(raw + mbn_vec(row, ch >> 3, cp / 8));
    const float y = fmaf(__bfloat162float(v[ch & 7]), bn[ch], bn[cp + ch]);
    out[i] = (y > 0.f && y < 6.f) ? 1 : 0;
  }
}

__global__ void __launch_bounds__(256) mbn_debug_stem_mask_kernel(const MbStemArgs a, uint8_t* __restrict__ out) {
  extern __shared__ float smem[];
  const int64_t b = blockIdx.x;
  const int pitch = mb_stem_stage(a, b, smem);
  for (int i = threadIdx.x; i < 3 * a.H * a.Wc; i += blockDim.x) {
    const int c = i / (a.H * a.Wc), rem = i - c * a.H * a.Wc, y = rem / a.Wc, x = rem - y * a.Wc;
    uint8_t m = 0;
    if ((x >> 1) < a.Wp) {
      const float sc = a.bn[c], sh = a.bn[16 + c];
      const float n = fmaf(mb_stem_conv(smem, pitch, a.w + c * 9, a.bias[c], y, x), sc, sh);
      const float other = fmaf(mb_stem_conv(smem, pitch, a.w + c * 9, a.bias[c], y, x ^ 1), sc, sh);
      const float r = fmaxf(n, 0.f), ro = fmaxf(other, 0.f);
      const bool win = (x & 1) ? (r > ro) : (r >= ro);
      m = (win && n > 0.f) ? 1 : 0;
    }
    out[b * 3 * a.H * a.Wc + i] = m;
  }
}

extern "C" int64_t howl_b200_mobilenet_debug_mask_bytes(int64_t B, int32_t frames, int32_t n_mels) {
  const MbNet net = mb_build(n_mels, frames, 1);
  int64_t n = B * 3 * n_mels * (frames + 4);
  for (size_t i = 1; i < net.convs.size(); ++i)
    if (net.convs[i].act) n += B * net.convs[i].hout * net.convs[i].wout * net.convs[i].cout;
  return n;
}

extern "C" int howl_b200_mobilenet_debug_masks(howl_ctx_t* ctx, void* stream, const float* feats, const float* params, int64_t B, int32_t frames,
                                               int32_t n_mels, int32_t num_labels, const void* workspace, size_t workspace_bytes, uint8_t* out) {
  if (!ctx) return HOWL_E_INVALID;
  HOWL_REQUIRE(ctx, feats && params && out, HOWL_E_INVALID, "mobilenet_debug_masks: null pointer");
  MbNet net;
  MbWs ws;
  int rc = mb_check(ctx, B, frames, n_mels, num_labels, workspace, workspace_bytes, &net, &ws);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  MbStemArgs a = mb_stem_args(net, ws, feats, params, B);
  const size_t sm = sizeof(float) * ((size_t)(a.H + 2) * (a.W + 8) + 3 * (size_t)a.H * a.Wp);
  HOWL_CUDA(ctx, cudaFuncSetAttribute(mbn_debug_stem_mask_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
  mbn_debug_stem_mask_kernel<<<(unsigned)B, 256, sm, st>>>(a, out);
  HOWL_LAUNCHED(ctx, "mbn_debug_stem_mask");
  uint8_t* p = out + B * 3 * a.H * a.Wc;
  for (size_t i = 1; i < net.convs.size(); ++i) {
    const MbConv& c = net.convs[i];
    if (!c.act) continue;
    const int64_t rows = B * c.hout * c.wout;
    mbn_debug_mask_kernel<<<ctx->sm_count * 4, 256, 0, st>>>(reinterpret_cast<const uint4*>(ws.raw[i]), ws.bn[i], c.cout, mbn_pad16(c.cout), rows, p);
    HOWL_LAUNCHED(ctx, "mbn_debug_mask");
    p += rows * c.cout;
  }
  return HOWL_OK;
}
