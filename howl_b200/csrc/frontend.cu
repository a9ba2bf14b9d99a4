// K1 -- fused audio frontend for sm_100a.
//
// One CTA = one utterance x a tile of FE_TILE frames.  The PCM span the tile needs (frames overlap by
// 312 of 512 samples, so every sample is staged once) is brought from HBM into shared memory with one TMA
// bulk copy (cp.async.bulk + mbarrier); each warp then owns one frame at a time:
//   reflect-pad indexing -> periodic Hann window -> 512-point real FFT as a 256-point complex radix-4
//   Stockham FFT in the warp's private shared-memory scratch -> |X|^2 -> sparse mel contraction (only the
//   non-zero span of every filterbank column) -> log(x + 1e-7) -> ZMUV -> SpecAugment mask
// and the finished tile is written with 128-bit coalesced stores in the layout the consumer wants.
// The [B,257,F] spectrogram of the reference (torchaudio MelSpectrogram, howl/data/transform/transform.py:249-254)
// never exists in HBM.  Arithmetic follows SURVEY.md App. A.1.
#include <math.h>

#include "common.cuh"

#define FE_TILE 27          // frames per CTA (81 = 3 tiles for 1 s clips, 41 = 27 + 14 for 0.5 s)
#define FE_WARPS 8
#define FE_THREADS (FE_WARPS * 32)
#define FE_FBC_CAP 1024     // compact filterbank entries kept in shared memory (standard 40-mel bank: 493, VTLP worst ~750)
#define FE_MAXU 128         // work units of the mel contraction: a filter span, or half of a span longer than 16 bins
#define FE_LANE_UNITS 8     // units one lane may own
#define FE_LOG_EPS 1e-7f

struct FeParams {
  const float* pcm;
  const float* fb;        // dense [257, M] (global) -- only used beyond FE_FBC_CAP
  const float* fbc;       // compact non-zero spans
  const int* fb_lo;       // [M]
  const int* fb_hi;       // [M]
  const int* fb_off;      // [M + 1]
  const int* mel_plan;    // balanced work plan built by fb_compact_kernel (layout: FePlan)
  const float* window;
  const float2* tw256;
  const float2* tw512;
  const int32_t* rects;
  float* out;
  int64_t B, T;
  int F, M, hop;
  float zmean, zstd;
  uint32_t flags;
  int use_tma;
};

// Balanced plan of the sparse mel contraction.  A unit = bins [lo, hi) of one filter (spans longer than 16 bins are cut
// in two); units are dealt to the 32 lanes longest-first onto the least loaded lane, so a frame costs ~nnz/32 + a few
// iterations per lane instead of the longest span plus the tail round.
struct FePlan {
  int n_units;
  int u_lo[FE_MAXU], u_hi[FE_MAXU], u_off[FE_MAXU];     // bin range and offset of the unit's first weight in fbc
  int mel_unit[HOWL_MAX_MELS][2];                       // units of every filter (-1 = none)
  int lane_n[32];
  int lane_unit[32][FE_LANE_UNITS];
};

// ---------------------------------------------------------------------------------------------
// compact filterbank: [lo, hi) non-zero span per column + prefix offsets.  One block, M threads.
// ---------------------------------------------------------------------------------------------
__global__ void fb_compact_kernel(const float* __restrict__ fb, int M, int* __restrict__ lo, int* __restrict__ hi,
                                  int* __restrict__ off, float* __restrict__ fbc, FePlan* __restrict__ plan) {
  extern __shared__ float s_fb[];              // the whole [257][M] bank, staged once with coalesced loads
  __shared__ int s_len[HOWL_MAX_MELS];
  __shared__ int s_lo2[HOWL_MAX_MELS];
  for (int i = threadIdx.x; i < HOWL_NFREQ * M; i += blockDim.x) s_fb[i] = fb[i];
  __syncthreads();
  __shared__ int s_off[HOWL_MAX_MELS + 1];
  const int m = threadIdx.x;
  int l = 0, h = 0;
  if (m < M) {
    l = HOWL_NFREQ;
    for (int j = 0; j < HOWL_NFREQ; ++j) {
      if (s_fb[j * M + m] != 0.f) {
        if (j < l) l = j;
        h = j + 1;
      }
    }
    if (h == 0) l = 0;
    lo[m] = l;
    hi[m] = h;
    s_len[m] = h - l;
    s_lo2[m] = l;
  }
  __syncthreads();
  if (m == 0) {
    int acc = 0;
    for (int i = 0; i < M; ++i) {
      s_off[i] = acc;
      acc += s_len[i];
    }
    s_off[M] = acc;
  }
  __syncthreads();
  if (m < M) {
    off[m] = s_off[m];
    for (int j = l; j < h; ++j) fbc[s_off[m] + (j - l)] = s_fb[j * M + m];
  }
  if (m == 0) off[M] = s_off[M];
  __syncthreads();
  __shared__ FePlan s_plan;      // planned in shared memory (no global-latency chain), copied out by all
  __shared__ short s_order[FE_MAXU];
  __shared__ short s_cnt[HOWL_NFREQ + 2];
  if (m == 0) {   // tiny serial planner (<= 128 units, 32 lanes)
    FePlan* plan = &s_plan;
    int n = 0;
    for (int i = 0; i < M; ++i) {
      const int l = s_lo2[i], h = s_lo2[i] + s_len[i], len = s_len[i];
      plan->mel_unit[i][0] = plan->mel_unit[i][1] = -1;
      if (len <= 0) continue;
      const int cut = (len > 16 && n + 2 <= FE_MAXU) ? l + len / 2 : h;
      plan->u_lo[n] = l; plan->u_hi[n] = cut; plan->u_off[n] = s_off[i];
      plan->mel_unit[i][0] = n++;
      if (cut < h) {
        plan->u_lo[n] = cut; plan->u_hi[n] = h; plan->u_off[n] = s_off[i] + (cut - l);
        plan->mel_unit[i][1] = n++;
      }
    }
    plan->n_units = n;
    // longest-first order by a counting sort on the span length (<= 257)
    for (int i = 0; i < HOWL_NFREQ + 2; ++i) s_cnt[i] = 0;
    for (int i = 0; i < n; ++i) s_cnt[plan->u_hi[i] - plan->u_lo[i]]++;
    {
      int pos = 0;
      for (int len = HOWL_NFREQ + 1; len >= 0; --len) {
        const int c = s_cnt[len];
        s_cnt[len] = (short)pos;
        pos += c;
      }
    }
    for (int i = 0; i < n; ++i) s_order[s_cnt[plan->u_hi[i] - plan->u_lo[i]]++] = (short)i;
  }
  __syncthreads();
  if (threadIdx.x < 32) {
    // greedy onto the least loaded lane (ties: lowest lane; lanes holding FE_LANE_UNITS units are closed): warp 0, lane i
    // keeps the load of plan lane i in a register, one warp-min per unit
    const int lane = threadIdx.x, n = s_plan.n_units;
    int my_load = 0, my_n = 0;
    for (int it = 0; it < n; ++it) {
      const int best = s_order[it];
      const int blen = s_plan.u_hi[best] - s_plan.u_lo[best];
      const unsigned key = (my_n < FE_LANE_UNITS) ? (unsigned)my_load * 32u + (unsigned)lane : 0xFFFFFFFFu;
      const unsigned sel = __reduce_min_sync(0xffffffffu, key) & 31u;
      if ((unsigned)lane == sel) {
        s_plan.lane_unit[lane][my_n++] = best;
        my_load += blen + 2;
      }
    }
    s_plan.lane_n[lane] = my_n;
  }
  __syncthreads();
  {
    const int* src = reinterpret_cast<const int*>(&s_plan);
    int* dst = reinterpret_cast<int*>(plan);
    for (int i = threadIdx.x; i < (int)(sizeof(FePlan) / sizeof(int)); i += blockDim.x) dst[i] = src[i];
  }
}

// ---------------------------------------------------------------------------------------------
// small PTX helpers (TMA 1-D bulk copy + mbarrier)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done = 0;
  while (!done) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  }
}
__device__ __forceinline__ void tma_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
  return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}

// one radix-4 Stockham pass over 256 complex points held in the warp's scratch (2 butterflies per lane)
template <int NS>
__device__ __forceinline__ void fft256_pass(const float2* __restrict__ in, float2* __restrict__ out, int lane,
                                            const float2* __restrict__ tw) {
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int j = lane + 32 * h;
    const int k = j & (NS - 1);
    float2 v0 = in[j], v1 = in[j + 64], v2 = in[j + 128], v3 = in[j + 192];
    if (NS > 1) {
      const int m = k * (64 / NS);
      v1 = cmul(v1, tw[m]);
      v2 = cmul(v2, tw[2 * m]);
      v3 = cmul(v3, tw[3 * m]);
    }
    const float2 t0 = make_float2(v0.x + v2.x, v0.y + v2.y);
    const float2 t1 = make_float2(v0.x - v2.x, v0.y - v2.y);
    const float2 t2 = make_float2(v1.x + v3.x, v1.y + v3.y);
    const float2 t3 = make_float2(v1.y - v3.y, -(v1.x - v3.x));  // (v1 - v3) * (-i)
    const int j0 = ((j - k) << 2) + k;
    out[j0] = make_float2(t0.x + t2.x, t0.y + t2.y);
    out[j0 + NS] = make_float2(t1.x + t3.x, t1.y + t3.y);
    out[j0 + 2 * NS] = make_float2(t0.x - t2.x, t0.y - t2.y);
    out[j0 + 3 * NS] = make_float2(t1.x - t3.x, t1.y - t3.y);
  }
  __syncwarp();
}

__global__ void __launch_bounds__(FE_THREADS) frontend_kernel(const FeParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  // carve-up (all offsets multiples of 16 bytes)
  const int span_cap = (FE_TILE - 1) * p.hop + HOWL_NFFT + 8;
  float* s_pcm = reinterpret_cast<float*>(smem_raw);
  float* s_win = s_pcm + ((span_cap + 3) & ~3);
  float2* s_tw256 = reinterpret_cast<float2*>(s_win + HOWL_NFFT);
  float2* s_tw512 = s_tw256 + 256;
  float* s_fbc = reinterpret_cast<float*>(s_tw512 + 258);
  float2* s_scratch = reinterpret_cast<float2*>(s_fbc + FE_FBC_CAP);        // [FE_WARPS][2][256]
  float* s_res = reinterpret_cast<float*>(s_scratch + FE_WARPS * 512);      // [FE_TILE][M]
  int* s_lo = reinterpret_cast<int*>(s_res + FE_TILE * p.M);
  int* s_hi = s_lo + HOWL_MAX_MELS;
  int* s_off = s_hi + HOWL_MAX_MELS;
  float* s_upart = reinterpret_cast<float*>(s_off + HOWL_MAX_MELS + 4);     // [FE_WARPS][FE_MAXU] unit partial sums
  __shared__ __align__(8) uint64_t s_bar;
  const FePlan* plan = reinterpret_cast<const FePlan*>(p.mel_plan);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t b = blockIdx.y;
  const int f0 = blockIdx.x * FE_TILE;
  const int nfr = min(FE_TILE, p.F - f0);
  const int64_t T = p.T;
  // staged span [lo, hi) of the clip; lo is a multiple of 4 samples, 4 below the first frame's start so
  // that the reflected tail (2(T-1)-s) of the last frame is always inside the span
  int64_t lo = (int64_t)f0 * p.hop - (HOWL_NFFT / 2 + 4);
  if (lo < 0) lo = 0;
  // (+4 above as well: frame 0 reflects s = -256 onto sample 256)
  int64_t hi = (int64_t)(f0 + nfr - 1) * p.hop + HOWL_NFFT / 2 + 4;
  if (hi > T) hi = T;
  const int span = (int)(hi - lo);
  const float* src = p.pcm + b * T + lo;

  if (p.use_tma) {
    if (tid == 0) mbar_init(&s_bar, 1);
    __syncthreads();
    if (tid == 0) {
      mbar_expect_tx(&s_bar, (uint32_t)span * 4u);
      tma_bulk_g2s(s_pcm, src, (uint32_t)span * 4u, &s_bar);
    }
  } else {
    for (int i = tid; i < span; i += FE_THREADS) s_pcm[i] = __ldg(src + i);
  }
  // tables (overlaps the bulk copy)
  for (int i = tid; i < HOWL_NFFT; i += FE_THREADS) s_win[i] = __ldg(p.window + i);
  for (int i = tid; i < 256; i += FE_THREADS) s_tw256[i] = __ldg(p.tw256 + i);
  for (int i = tid; i < HOWL_NFREQ; i += FE_THREADS) s_tw512[i] = __ldg(p.tw512 + i);
  for (int i = tid; i < p.M; i += FE_THREADS) {
    s_lo[i] = __ldg(p.fb_lo + i);
    s_hi[i] = __ldg(p.fb_hi + i);
  }
  for (int i = tid; i <= p.M; i += FE_THREADS) s_off[i] = __ldg(p.fb_off + i);
  {
    const int total = min(__ldg(p.fb_off + p.M), FE_FBC_CAP);
    for (int i = tid; i < total; i += FE_THREADS) s_fbc[i] = __ldg(p.fbc + i);
  }
  __syncthreads();
  if (p.use_tma) mbar_wait(&s_bar, 0);

  float2* bufA = s_scratch + warp * 512;
  float2* bufB = bufA + 256;
  float* pw = reinterpret_cast<float*>(bufB);  // power spectrum [257] re-uses bufB after the FFT

  int rf0 = 0, rfl = 0, rt0 = 0, rtl = 0;
  const bool masked = (p.rects != nullptr) && !(p.flags & HOWL_FE_STACKED);
  if (masked) {
    rf0 = p.rects[b * 4 + 0];
    rfl = p.rects[b * 4 + 1];
    rt0 = p.rects[b * 4 + 2];
    rtl = p.rects[b * 4 + 3];
  }
  const bool do_zmuv = (p.flags & HOWL_FE_ZMUV) && !(p.flags & HOWL_FE_STACKED);

  for (int fi = warp; fi < nfr; fi += FE_WARPS) {
    const int f = f0 + fi;
    const int64_t start = (int64_t)f * p.hop - HOWL_NFFT / 2;  // first sample of the frame (may be < 0)
    const bool interior = (start >= 0) && (start + HOWL_NFFT <= T);   // no reflection: plain 64-bit vector loads
    const float2* fr2 = reinterpret_cast<const float2*>(s_pcm + (interior ? (int)(start - lo) : 0));
    const float2* win2 = reinterpret_cast<const float2*>(s_win);
    // ---- pass 0 (NS = 1, no twiddles) straight from the staged PCM: z[n] = x[2n] w[2n] + i x[2n+1] w[2n+1]
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int j = lane + 32 * h;
      float2 v[4];
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int n = j + 64 * r;
        if (interior) {
          const float2 x = fr2[n], w = win2[n];
          v[r] = make_float2(x.x * w.x, x.y * w.y);
        } else {
          int64_t sa = start + 2 * n, sb = sa + 1;
          if (sa < 0) sa = -sa;
          if (sb < 0) sb = -sb;
          if (sa >= T) sa = 2 * (T - 1) - sa;
          if (sb >= T) sb = 2 * (T - 1) - sb;
          v[r] = make_float2(s_pcm[sa - lo] * s_win[2 * n], s_pcm[sb - lo] * s_win[2 * n + 1]);
        }
      }
      const float2 t0 = make_float2(v[0].x + v[2].x, v[0].y + v[2].y);
      const float2 t1 = make_float2(v[0].x - v[2].x, v[0].y - v[2].y);
      const float2 t2 = make_float2(v[1].x + v[3].x, v[1].y + v[3].y);
      const float2 t3 = make_float2(v[1].y - v[3].y, -(v[1].x - v[3].x));
      const int j0 = j << 2;
      bufB[j0] = make_float2(t0.x + t2.x, t0.y + t2.y);
      bufB[j0 + 1] = make_float2(t1.x + t3.x, t1.y + t3.y);
      bufB[j0 + 2] = make_float2(t0.x - t2.x, t0.y - t2.y);
      bufB[j0 + 3] = make_float2(t1.x - t3.x, t1.y - t3.y);
    }
    __syncwarp();
    fft256_pass<4>(bufB, bufA, lane, s_tw256);
    fft256_pass<16>(bufA, bufB, lane, s_tw256);
    fft256_pass<64>(bufB, bufA, lane, s_tw256);  // Z in bufA, natural order
    // ---- real-FFT post-processing + power: X[k] = E[k] + W512^k O[k], k = 0..256
#pragma unroll
    for (int h = 0; h < 9; ++h) {
      const int k = lane + 32 * h;
      if (k <= 256) {
        const float2 zk = bufA[k & 255];
        float2 zc = bufA[(256 - k) & 255];
        zc.y = -zc.y;
        const float2 e = make_float2(0.5f * (zk.x + zc.x), 0.5f * (zk.y + zc.y));
        const float2 d = make_float2(zk.x - zc.x, zk.y - zc.y);
        const float2 o = make_float2(0.5f * d.y, -0.5f * d.x);  // d / (2i)
        const float2 wo = cmul(s_tw512[k], o);
        const float xr = e.x + wo.x, xi = e.y + wo.y;
        pw[k] = xr * xr + xi * xi;
      }
    }
    __syncwarp();
    // ---- sparse mel contraction (balanced units) + log + zmuv + mask
    float* up = s_upart + warp * FE_MAXU;
    {
      const int nu = plan->lane_n[lane];
      for (int q = 0; q < nu; ++q) {
        const int u = plan->lane_unit[lane][q];
        const int jl = plan->u_lo[u], jh = plan->u_hi[u], off = plan->u_off[u];
        float acc = 0.f;
        for (int j = jl; j < jh; ++j) {
          const int e = off + (j - jl);
          const float w = (e < FE_FBC_CAP) ? s_fbc[e] : __ldg(p.fbc + e);
          acc = fmaf(pw[j], w, acc);
        }
        up[u] = acc;
      }
    }
    __syncwarp();
    for (int m = lane; m < p.M; m += 32) {
      const int u0 = plan->mel_unit[m][0], u1 = plan->mel_unit[m][1];
      float acc = (u0 >= 0 ? up[u0] : 0.f) + (u1 >= 0 ? up[u1] : 0.f);
      float v = logf(acc + FE_LOG_EPS);
      if (do_zmuv) v = __fdiv_rn(v - p.zmean, p.zstd);
      if (masked && ((m >= rf0 && m < rf0 + rfl) || (f >= rt0 && f < rt0 + rtl))) v = 0.f;
      s_res[fi * p.M + m] = v;
    }
    __syncwarp();
  }
  __syncthreads();

  // ---- coalesced tile store
  const int M = p.M, F = p.F;
  if (p.flags & HOWL_FE_TIME_MAJOR) {
    float* dst = p.out + ((int64_t)b * F + f0) * M;
    const int n = nfr * M;
    if ((M & 3) == 0) {
      float4* d4 = reinterpret_cast<float4*>(dst);
      const float4* s4 = reinterpret_cast<const float4*>(s_res);
      for (int i = tid; i < n / 4; i += FE_THREADS) d4[i] = s4[i];
    } else {
      for (int i = tid; i < n; i += FE_THREADS) dst[i] = s_res[i];
    }
  } else {
    // [B, (3,) M, F]: rows of the tile are contiguous along f
    const int64_t chan_stride = (p.flags & HOWL_FE_STACKED) ? 3 : 1;
    float* dst = p.out + (int64_t)b * chan_stride * M * F + f0;
    for (int i = tid; i < nfr * M; i += FE_THREADS) {
      const int m = i / nfr, fi = i - m * nfr;
      dst[(int64_t)m * F + fi] = s_res[fi * M + m];
    }
  }
}

// ---------------------------------------------------------------------------------------------
// deltas + ZMUV + SpecAugment for the stacked [B,3,M,F] layout: one warp per (b, m) row.
// d[t] = (-2x[t-2] - x[t-1] + x[t+1] + 2x[t+2]) / 10 with replicate padding, applied twice
// (torchaudio ComputeDeltas); channel 0 holds raw log-mel on entry.
// ---------------------------------------------------------------------------------------------
__global__ void deltas_kernel(float* __restrict__ out, int64_t rows, int M, int F, float zmean, float zstd,
                              int do_zmuv, const int32_t* __restrict__ rects) {
  extern __shared__ float s_rows[];  // [warps][2][F]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  float* xs = s_rows + (size_t)warp * 2 * F;
  float* ds = xs + F;
  for (int64_t row = (int64_t)blockIdx.x * nw + warp; row < rows; row += (int64_t)gridDim.x * nw) {
    const int64_t b = row / M;
    const int m = (int)(row - b * M);
    float* c0 = out + ((b * 3 + 0) * M + m) * (int64_t)F;
    float* c1 = out + ((b * 3 + 1) * M + m) * (int64_t)F;
    float* c2 = out + ((b * 3 + 2) * M + m) * (int64_t)F;
    for (int t = lane; t < F; t += 32) xs[t] = c0[t];
    __syncwarp();
    for (int t = lane; t < F; t += 32) {
      const float a = xs[max(t - 2, 0)], bb = xs[max(t - 1, 0)], c = xs[min(t + 1, F - 1)], d = xs[min(t + 2, F - 1)];
      ds[t] = __fdiv_rn(-2.f * a - bb + c + 2.f * d, 10.f);
    }
    __syncwarp();
    int rf0 = 0, rfl = 0, rt0 = 0, rtl = 0;
    if (rects) {
      rf0 = rects[b * 4 + 0];
      rfl = rects[b * 4 + 1];
      rt0 = rects[b * 4 + 2];
      rtl = rects[b * 4 + 3];
    }
    const bool fm = (m >= rf0 && m < rf0 + rfl);
    for (int t = lane; t < F; t += 32) {
      const float a = ds[max(t - 2, 0)], bb = ds[max(t - 1, 0)], c = ds[min(t + 1, F - 1)], d = ds[min(t + 2, F - 1)];
      float v0 = xs[t], v1 = ds[t], v2 = __fdiv_rn(-2.f * a - bb + c + 2.f * d, 10.f);
      if (do_zmuv) {
        v0 = __fdiv_rn(v0 - zmean, zstd);
        v1 = __fdiv_rn(v1 - zmean, zstd);
        v2 = __fdiv_rn(v2 - zmean, zstd);
      }
      if (fm || (t >= rt0 && t < rt0 + rtl)) v0 = v1 = v2 = 0.f;
      c0[t] = v0;
      c1[t] = v1;
      c2[t] = v2;
    }
    __syncwarp();
  }
}

// ---------------------------------------------------------------------------------------------
// sum / sum of squares (ZmuvTransform.update)
// ---------------------------------------------------------------------------------------------
__global__ void sum_sumsq_kernel(const float* __restrict__ x, int64_t n, double* __restrict__ sums) {
  double s = 0.0, s2 = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const double v = x[i];
    s += v;
    s2 += v * v;
  }
  s = warp_sum(s);
  s2 = warp_sum(s2);
  __shared__ double sh[2][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) {
    sh[0][warp] = s;
    sh[1][warp] = s2;
  }
  __syncthreads();
  if (warp == 0) {
    const int nw = blockDim.x >> 5;
    s = lane < nw ? sh[0][lane] : 0.0;
    s2 = lane < nw ? sh[1][lane] : 0.0;
    s = warp_sum(s);
    s2 = warp_sum(s2);
    if (lane == 0) {
      atomicAdd(&sums[0], s);
      atomicAdd(&sums[1], s2);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
static int fe_scratch(howl_ctx_t* ctx) {
  if (!ctx->fb_lo) {
    HOWL_CUDA(ctx, cudaMalloc(&ctx->fb_lo, sizeof(int) * HOWL_MAX_MELS));
    HOWL_CUDA(ctx, cudaMalloc(&ctx->fb_hi, sizeof(int) * HOWL_MAX_MELS));
    HOWL_CUDA(ctx, cudaMalloc(&ctx->fb_off, sizeof(int) * (HOWL_MAX_MELS + 1)));
    HOWL_CUDA(ctx, cudaMalloc(&ctx->fbc, sizeof(float) * HOWL_NFREQ * HOWL_MAX_MELS));
    HOWL_CUDA(ctx, cudaMalloc(&ctx->mel_plan, sizeof(FePlan)));
  }
  return HOWL_OK;
}

size_t howl_fe_smem_bytes(int hop, int M) {
  const int span_cap = (FE_TILE - 1) * hop + HOWL_NFFT + 8;
  size_t b = 0;
  b += sizeof(float) * ((span_cap + 3) & ~3);
  b += sizeof(float) * HOWL_NFFT;
  b += sizeof(float2) * (256 + 258);
  b += sizeof(float) * FE_FBC_CAP;
  b += sizeof(float2) * FE_WARPS * 512;
  b += sizeof(float) * FE_TILE * M;
  b += sizeof(int) * (3 * HOWL_MAX_MELS + 4);
  b += sizeof(float) * FE_WARPS * FE_MAXU;
  return howl_align_up(b, 16);
}

extern "C" int howl_b200_frontend_fwd(howl_ctx_t* ctx, void* stream, const float* pcm, int64_t B, int64_t T,
                                      const float* fb, float zmuv_mean, float zmuv_std, const int32_t* rects,
                                      uint32_t flags, float* out) {
  if (!ctx) return HOWL_E_INVALID;
  HOWL_REQUIRE(ctx, pcm && fb && out, HOWL_E_INVALID, "frontend_fwd: null pointer");
  HOWL_REQUIRE(ctx, B >= 0 && B <= 0x7fffffffLL / 4, HOWL_E_INVALID, "frontend_fwd: bad batch %lld", (long long)B);
  HOWL_REQUIRE(ctx, T > HOWL_NFFT / 2, HOWL_E_INVALID,
               "frontend_fwd: T=%lld must exceed n_fft/2=%d (reflect padding)", (long long)T, HOWL_NFFT / 2);
  const int layouts = ((flags & HOWL_FE_TIME_MAJOR) != 0) + ((flags & HOWL_FE_MELS_ONLY) != 0) +
                      ((flags & HOWL_FE_STACKED) != 0);
  HOWL_REQUIRE(ctx, layouts == 1, HOWL_E_INVALID, "frontend_fwd: exactly one output layout flag required");
  if ((flags & HOWL_FE_ZMUV)) HOWL_REQUIRE(ctx, zmuv_std > 0.f, HOWL_E_INVALID, "frontend_fwd: zmuv_std must be > 0");
  if (B == 0) return HOWL_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const int M = ctx->fe.n_mels, hop = ctx->fe.hop;
  const int64_t F64 = 1 + T / hop;
  HOWL_REQUIRE(ctx, F64 <= 0x7fffffff / HOWL_MAX_MELS, HOWL_E_UNSUPPORTED, "frontend_fwd: clip too long");
  const int F = (int)F64;
  int rc = fe_scratch(ctx);
  if (rc) return rc;
  const size_t fbsm = sizeof(float) * HOWL_NFREQ * M;
  HOWL_CUDA(ctx, cudaFuncSetAttribute(fb_compact_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fbsm));
  fb_compact_kernel<<<1, 512, fbsm, st>>>(fb, M, ctx->fb_lo, ctx->fb_hi, ctx->fb_off, ctx->fbc,
                                                 reinterpret_cast<FePlan*>(ctx->mel_plan));
  HOWL_LAUNCHED(ctx, "fb_compact");

  FeParams p;
  p.pcm = pcm; p.fb = fb; p.fbc = ctx->fbc; p.fb_lo = ctx->fb_lo; p.fb_hi = ctx->fb_hi; p.fb_off = ctx->fb_off; p.mel_plan = ctx->mel_plan;
  p.window = ctx->d_window; p.tw256 = ctx->d_tw256; p.tw512 = ctx->d_tw512;
  p.rects = rects; p.out = out; p.B = B; p.T = T; p.F = F; p.M = M; p.hop = hop;
  p.zmean = zmuv_mean; p.zstd = zmuv_std; p.flags = flags;
  p.use_tma = ((T & 3) == 0) && ((reinterpret_cast<uintptr_t>(pcm) & 15) == 0) && ((hop & 3) == 0);
  const size_t smem = howl_fe_smem_bytes(hop, M);
  HOWL_CUDA(ctx, cudaFuncSetAttribute(frontend_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((unsigned)howl_ceil_div(F, FE_TILE), (unsigned)1, 1);
  // grid.y is limited to 65535: walk the batch in slabs
  for (int64_t b0 = 0; b0 < B; b0 += 65535) {
    const int64_t nb = (B - b0 < 65535) ? (B - b0) : 65535;
    FeParams q = p;
    q.pcm = pcm + b0 * T;
    q.rects = rects ? rects + b0 * 4 : nullptr;
    const int64_t per = (flags & HOWL_FE_STACKED) ? 3LL * M * F : (int64_t)M * F;
    q.out = out + b0 * per;
    q.use_tma = p.use_tma && ((reinterpret_cast<uintptr_t>(q.pcm) & 15) == 0);
    grid.y = (unsigned)nb;
    frontend_kernel<<<grid, FE_THREADS, smem, st>>>(q);
    HOWL_LAUNCHED(ctx, "frontend");
  }
  if (flags & HOWL_FE_STACKED) {
    const int64_t rows = B * M;
    const int warps = 8;
    const size_t sm = sizeof(float) * warps * 2 * F;
    HOWL_REQUIRE(ctx, sm <= 200 * 1024, HOWL_E_UNSUPPORTED, "frontend_fwd: clip too long for the delta kernel");
    HOWL_CUDA(ctx, cudaFuncSetAttribute(deltas_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    int64_t blocks = howl_ceil_div(rows, warps);
    const int64_t cap = (int64_t)ctx->sm_count * 16;
    if (blocks > cap) blocks = cap;
    deltas_kernel<<<(unsigned)blocks, warps * 32, sm, st>>>(out, rows, M, F, zmuv_mean, zmuv_std,
                                                            (flags & HOWL_FE_ZMUV) ? 1 : 0, rects);
    HOWL_LAUNCHED(ctx, "deltas");
  }
  return HOWL_OK;
}

extern "C" int howl_b200_sum_sumsq(howl_ctx_t* ctx, void* stream, const float* x, int64_t n, double* sums) {
  if (!ctx) return HOWL_E_INVALID;
  HOWL_REQUIRE(ctx, x && sums && n >= 0, HOWL_E_INVALID, "sum_sumsq: bad argument");
  if (n == 0) return HOWL_OK;
  int64_t blocks = howl_ceil_div(n, 256 * 8);
  const int64_t cap = (int64_t)ctx->sm_count * 8;
  if (blocks > cap) blocks = cap;
  sum_sumsq_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(x, n, sums);
  HOWL_LAUNCHED(ctx, "sum_sumsq");
  return HOWL_OK;
}
