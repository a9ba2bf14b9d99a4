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
#define FE_WARPS 9          // one frame per warp at a time: 27 = 3 x 9
#define FE_THREADS (FE_WARPS * 32)
#define FE_PART 272         // floats of per-warp scratch: FE_ENT_SMEM partial sums, or the 257 power bins in the dense fallback
#define FE_ENT_SMEM 256     // mel entries kept in shared memory (standard 40-mel bank: ~100, 80 mels: ~170); the rest is read from global
#define FE_LOG_EPS 1e-7f

// Sparse mel contraction, organised by BIN BLOCKS: the FFT below leaves lane L of a warp with the power of the 8 consecutive bins
// [8 * br5(L), 8 * br5(L) + 8) in registers, so the filterbank is cut into "entries" (block, filter m, the 8 weights fb[8 blk + i][m])
// for every (block, filter) pair with a non-zero weight.  A lane walks the entries of its block -- 8 FMAs and one shared-memory
// store of the partial sum per entry; afterwards lane m gathers the partial sums of filter m's entries (no shared-memory atomics: fp32
// atomicAdd on shared memory is a compare-and-swap spin loop) -- and bin 256 (Nyquist) has its own (filter, weight) list.  Built on the device from any dense [257, M] bank
// (VTLP-warped ones included), rebuilt only when the bank changes.
struct FeEntry {
  int m;
  float w[8];
};
struct FeBank {
  int ent_off[33];          // entries of block blk: [ent_off[blk], ent_off[blk + 1])
  int ny_count;             // filters with a non-zero Nyquist weight
  int ny_m[HOWL_MAX_MELS];
  float ny_w[HOWL_MAX_MELS];
  float ny_dense[HOWL_MAX_MELS];      // Nyquist weight of every filter (0 if none)
  int filt_off[HOWL_MAX_MELS + 1];    // entries of filter m (one per block it overlaps): filt_idx[filt_off[m] .. filt_off[m + 1])
  int filt_idx[32 * HOWL_MAX_MELS];
};
#define FE_ENT_WORDS 8          // the 8 weights of a (block of 8 bins, filter) pair: two aligned 128-bit shared-memory reads

struct FeParams {
  const float* pcm;
  const float* fb;        // dense [257, M] bank (only read by the fallback for banks with more than FE_ENT_SMEM entries)
  const FeBank* bank;
  const float* ent;       // [n_entries][8]: the weights of bins 8 blk .. 8 blk + 7 for one (block, filter) pair
  const float* window;
  const float2* tw_lane;  // [32][8]: W256^(L * m2), m2 = 0..7
  const float2* tw_stage; // [32][4]: the cross-lane stage twiddles of lane L (spans 16, 8, 4, 2)
  const float2* w512_lane;// [32][8]: W512^(m2 + 8 * br5(L))
  const int32_t* rects;
  float* out;
  int64_t B, T;
  int F, M, hop;
  float zmean, zstd;
  uint32_t flags;
  int use_tma;
  int pcm_i16;            // pcm points to int16 samples
};

// ---------------------------------------------------------------------------------------------
// filterbank -> block entries.  One CTA; thread t scans the (block, filter) pairs t, t + 1024, ...
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) fb_compact_kernel(const float* __restrict__ fb, int M, FeBank* __restrict__ bank,
                                                          float* __restrict__ ent) {
  __shared__ unsigned char s_nz[32][HOWL_MAX_MELS];     // (block, filter) has a non-zero weight
  __shared__ short s_ent[32][HOWL_MAX_MELS];            // its entry index
  __shared__ int s_off[33];
  __shared__ int s_fcnt[HOWL_MAX_MELS + 1];
  const int tid = threadIdx.x;
  for (int pidx = tid; pidx < 32 * M; pidx += blockDim.x) {
    const int blk = pidx / M, m = pidx - blk * M;
    bool nz = false;
    for (int i = 0; i < 8; ++i) nz |= fb[(8 * blk + i) * M + m] != 0.f;
    s_nz[blk][m] = nz ? 1 : 0;
  }
  __syncthreads();
  if (tid < 32) {            // rank of every non-empty pair inside its block (entries of a block are in filter order)
    int n = 0;
    for (int m = 0; m < M; ++m) {
      s_ent[tid][m] = (short)n;
      n += s_nz[tid][m];
    }
    s_off[tid + 1] = n;      // count for now
  }
  if (tid >= 64 && tid < 64 + M) {
    const int m = tid - 64;
    int n = 0;
    for (int blk = 0; blk < 32; ++blk) n += s_nz[blk][m];
    s_fcnt[m + 1] = n;
    bank->ny_dense[m] = fb[256 * M + m];
  }
  __syncthreads();
  if (tid == 0) {
    s_off[0] = 0;
    for (int blk = 0; blk < 32; ++blk) s_off[blk + 1] += s_off[blk];
    for (int blk = 0; blk <= 32; ++blk) bank->ent_off[blk] = s_off[blk];
    s_fcnt[0] = 0;
    for (int m = 0; m < M; ++m) s_fcnt[m + 1] += s_fcnt[m];
    for (int m = 0; m <= M; ++m) bank->filt_off[m] = s_fcnt[m];
    int n = 0;
    for (int m = 0; m < M; ++m) {
      const float w = fb[256 * M + m];
      if (w != 0.f) {
        bank->ny_m[n] = m;
        bank->ny_w[n] = w;
        ++n;
      }
    }
    bank->ny_count = n;
  }
  __syncthreads();
  // entries: thread per (block, filter) pair writes its 8 weights; per-filter lists: thread per filter
  for (int pidx = tid; pidx < 32 * M; pidx += blockDim.x) {
    const int blk = pidx / M, m = pidx - blk * M;
    if (!s_nz[blk][m]) continue;
    const int e = s_off[blk] + s_ent[blk][m];
    for (int i = 0; i < 8; ++i) ent[e * FE_ENT_WORDS + i] = fb[(8 * blk + i) * M + m];
  }
  if (tid < M) {
    int n = s_fcnt[tid];
    for (int blk = 0; blk < 32; ++blk)
      if (s_nz[blk][tid]) bank->filt_idx[n++] = s_off[blk] + s_ent[blk][tid];
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
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 cmul_mi(float2 a) { return make_float2(a.y, -a.x); }   // a * (-i)

// 256-point complex FFT of one frame held by a warp: lane L enters with z[L + 32 k] in v[k] and leaves with X[m2 + 8 * br5(L)]
// in v[m2] (br5 = 5-bit reversal).  N = 8 x 32:  X[m2 + 8 m1] = sum_n1 W256^(n1 m2) W32^(n1 m1) [ sum_n2 z[n1 + 32 n2] W8^(n2 m2) ]
//   (1) the bracket: an 8-point DFT over the lane's own registers;
//   (2) the twiddles W256^(L m2) (per-lane constants tw[m2]);
//   (3) for every register a 32-point DFT ACROSS the lanes: five radix-2 decimation-in-frequency stages whose butterflies exchange
//       partners with __shfl_xor (span 16, 8, 4, 2, 1); the upper lane of a pair keeps (x_lo - x_hi) * W, the lower x_lo + x_hi.
__device__ __forceinline__ void fft256_warp(float2* v, const float2* tw, const float2* st, int lane) {
  const float c8 = 0.70710678118654752440f;
  {
    // ---- (1) 8-point DIF in registers; outputs renamed to natural order
    float2 u0 = cadd(v[0], v[4]), u1 = cadd(v[1], v[5]), u2 = cadd(v[2], v[6]), u3 = cadd(v[3], v[7]);
    float2 d0 = csub(v[0], v[4]), d1 = csub(v[1], v[5]), d2 = csub(v[2], v[6]), d3 = csub(v[3], v[7]);
    d1 = make_float2(c8 * (d1.x + d1.y), c8 * (d1.y - d1.x));        // * W8^1 = (c, -c)
    d2 = cmul_mi(d2);                                                // * W8^2 = -i
    d3 = make_float2(c8 * (d3.y - d3.x), -c8 * (d3.x + d3.y));       // * W8^3 = (-c, -c)
    const float2 uu0 = cadd(u0, u2), uu1 = cadd(u1, u3), uv0 = csub(u0, u2), uv1 = cmul_mi(csub(u1, u3));
    const float2 vu0 = cadd(d0, d2), vu1 = cadd(d1, d3), vv0 = csub(d0, d2), vv1 = cmul_mi(csub(d1, d3));
    v[0] = cadd(uu0, uu1); v[4] = csub(uu0, uu1);
    v[2] = cadd(uv0, uv1); v[6] = csub(uv0, uv1);
    v[1] = cadd(vu0, vu1); v[5] = csub(vu0, vu1);
    v[3] = cadd(vv0, vv1); v[7] = csub(vv0, vv1);
  }
  // ---- (2)
#pragma unroll
  for (int m2 = 1; m2 < 8; ++m2) v[m2] = cmul(v[m2], tw[m2]);
  // ---- (3)
#pragma unroll
  for (int s = 0; s < 5; ++s) {
    const int h = 16 >> s;
    const bool upper = (lane & h) != 0;
    const float sgn = upper ? -1.f : 1.f;          // partner + sgn * own: one FFMA per component, the same rounding as the add / subtract
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      const float ox = __shfl_xor_sync(0xffffffffu, v[r].x, h), oy = __shfl_xor_sync(0xffffffffu, v[r].y, h);
      float2 t = make_float2(fmaf(sgn, v[r].x, ox), fmaf(sgn, v[r].y, oy));
      if (s < 3) t = cmul(t, st[s]);              // st[s] = 1 in the lower lane
      else if (s == 3 && upper && (lane & 1)) t = cmul_mi(t);      // W4^1 = -i
      v[r] = t;
    }
  }
}

__global__ void __launch_bounds__(FE_THREADS, 3) frontend_kernel(const FeParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  // carve-up (all offsets multiples of 16 bytes)
  const int span_cap = (FE_TILE - 1) * p.hop + HOWL_NFFT + 8;
  float* s_pcm = reinterpret_cast<float*>(smem_raw);
  float* s_win = s_pcm + ((span_cap + 3) & ~3);
  float2* s_w512 = reinterpret_cast<float2*>(s_win + HOWL_NFFT);            // [8][32 lanes]
  float* s_ent = reinterpret_cast<float*>(s_w512 + 256);                    // [FE_ENT_SMEM][8]
  float* s_part = s_ent + FE_ENT_SMEM * FE_ENT_WORDS;                       // [FE_WARPS][FE_ENT_SMEM] partial sums of the entries
  float* s_res = s_part + FE_WARPS * FE_PART;                               // [FE_TILE][M]
  int* s_fidx = reinterpret_cast<int*>(s_res + FE_TILE * p.M);              // [FE_ENT_SMEM] entry ids grouped by filter
  int* s_foff = s_fidx + FE_ENT_SMEM;                                       // [M + 1]
  float* s_nyw = reinterpret_cast<float*>(s_foff + HOWL_MAX_MELS + 1);      // [M] Nyquist weights
  __shared__ __align__(8) uint64_t s_bar;
  __shared__ int s_off[33];

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

  if (p.pcm_i16) {
    // int16 PCM: x / 32768 (exact in fp32) while staging; 4 samples per 64-bit load where the span start is 8-byte aligned
    const short* s16 = reinterpret_cast<const short*>(p.pcm) + b * T + lo;
    if ((reinterpret_cast<uintptr_t>(s16) & 7) == 0) {
      const int n4 = span >> 2;
      for (int i = tid; i < n4; i += FE_THREADS) {
        const short4 q = __ldg(reinterpret_cast<const short4*>(s16) + i);
        reinterpret_cast<float4*>(s_pcm)[i] = make_float4(q.x * (1.f / 32768.f), q.y * (1.f / 32768.f), q.z * (1.f / 32768.f), q.w * (1.f / 32768.f));
      }
      for (int i = (n4 << 2) + tid; i < span; i += FE_THREADS) s_pcm[i] = s16[i] * (1.f / 32768.f);
    } else {
      for (int i = tid; i < span; i += FE_THREADS) s_pcm[i] = s16[i] * (1.f / 32768.f);
    }
  } else if (p.use_tma) {
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
  for (int i = tid; i < 256; i += FE_THREADS) s_w512[(i & 7) * 32 + (i >> 3)] = __ldg(p.w512_lane + i);   // [m2][lane]: conflict-free reads
  if (tid < 33) s_off[tid] = p.bank->ent_off[tid];
  {
    const int total = min(p.bank->ent_off[32], FE_ENT_SMEM) * FE_ENT_WORDS;
    for (int i = tid; i < total; i += FE_THREADS) s_ent[i] = __ldg(p.ent + i);
  }
  const int n_ent = p.bank->ent_off[32];
  const bool small_bank = n_ent <= FE_ENT_SMEM;        // the usual case: every entry and list fits in shared memory
  for (int i = tid; i < min(n_ent, FE_ENT_SMEM); i += FE_THREADS) s_fidx[i] = p.bank->filt_idx[i];
  for (int i = tid; i <= p.M; i += FE_THREADS) s_foff[i] = p.bank->filt_off[i];
  for (int i = tid; i < p.M; i += FE_THREADS) s_nyw[i] = p.bank->ny_dense[i];
  // per-lane FFT constants
  float2 tw[8], st[3];
#pragma unroll
  for (int m2 = 0; m2 < 8; ++m2) tw[m2] = __ldg(p.tw_lane + lane * 8 + m2);
#pragma unroll
  for (int i = 0; i < 3; ++i) st[i] = __ldg(p.tw_stage + lane * 4 + i);
  const int blk = (int)(__brev((unsigned)lane) >> 27);               // this lane's bin block = br5(lane)
  const int src0 = (int)(__brev((unsigned)((32 - blk) & 31)) >> 27); // lane holding X[256 - 8 * blk] in register 0
  __syncthreads();
  if (p.use_tma) mbar_wait(&s_bar, 0);

  int rf0 = 0, rfl = 0, rt0 = 0, rtl = 0;
  const bool masked = (p.rects != nullptr) && !(p.flags & HOWL_FE_STACKED);
  if (masked) {
    rf0 = p.rects[b * 4 + 0];
    rfl = p.rects[b * 4 + 1];
    rt0 = p.rects[b * 4 + 2];
    rtl = p.rects[b * 4 + 3];
  }
  const bool do_zmuv = (p.flags & HOWL_FE_ZMUV) && !(p.flags & HOWL_FE_STACKED);
  float* part = s_part + warp * FE_PART;
  const int e_begin = s_off[blk], e_end = s_off[blk + 1];

  for (int fi = warp; fi < nfr; fi += FE_WARPS) {
    const int f = f0 + fi;
    const int64_t start = (int64_t)f * p.hop - HOWL_NFFT / 2;  // first sample of the frame (may be < 0)
    const bool interior = (start >= 0) && (start + HOWL_NFFT <= T);   // no reflection: plain 64-bit vector loads
    const float2* fr2 = reinterpret_cast<const float2*>(s_pcm + (interior ? (int)(start - lo) : 0));
    const float2* win2 = reinterpret_cast<const float2*>(s_win);
    // ---- z[n] = x[2n] w[2n] + i x[2n+1] w[2n+1],  n = lane + 32 k
    float2 v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int n = lane + 32 * k;
      const float2 w = win2[n];
      if (interior) {
        const float2 x = fr2[n];
        v[k] = make_float2(x.x * w.x, x.y * w.y);
      } else {
        int64_t sa = start + 2 * n, sb = sa + 1;
        if (sa < 0) sa = -sa;
        if (sb < 0) sb = -sb;
        if (sa >= T) sa = 2 * (T - 1) - sa;
        if (sb >= T) sb = 2 * (T - 1) - sb;
        v[k] = make_float2(s_pcm[sa - lo] * w.x, s_pcm[sb - lo] * w.y);
      }
    }
    fft256_warp(v, tw, st, lane);
    // ---- real-FFT post-processing + power: X[m] = E + W512^m O with E = (Z[m] + conj Z[256-m]) / 2, O = (Z[m] - conj Z[256-m]) / 2i;
    //      Z[256 - m] for m = m2 + 8 blk sits in register 8 - m2 of lane 31 - L (m2 > 0) or register 0 of lane src0 (m2 = 0)
    float pw[8];
    float nyq = 0.f;
#pragma unroll
    for (int m2 = 0; m2 < 8; ++m2) {
      const int r = (8 - m2) & 7;
      const int from = m2 == 0 ? src0 : (31 - lane);
      float2 zc = make_float2(__shfl_sync(0xffffffffu, v[r].x, from), -__shfl_sync(0xffffffffu, v[r].y, from));
      const float2 zk = v[m2];
      const float2 e = make_float2(0.5f * (zk.x + zc.x), 0.5f * (zk.y + zc.y));
      const float2 d = make_float2(zk.x - zc.x, zk.y - zc.y);
      const float2 o = make_float2(0.5f * d.y, -0.5f * d.x);  // d / (2i)
      const float2 wo = cmul(s_w512[m2 * 32 + lane], o);
      const float xr = e.x + wo.x, xi = e.y + wo.y;
      pw[m2] = xr * xr + xi * xi;
      if (m2 == 0) {                     // bin 256 from Z[0] (lane 0 only uses it): X[256] = Re Z[0] - Im Z[0]
        const float t = zk.x - zk.y;
        nyq = t * t;
      }
    }
    // ---- sparse mel contraction: this lane's block entries -> one partial sum each; then lane m gathers filter m's entries
    const float nyq0 = __shfl_sync(0xffffffffu, nyq, 0);
    if (small_bank) {
      for (int eidx = e_begin; eidx < e_end; ++eidx) {
        const float4 w0 = *reinterpret_cast<const float4*>(s_ent + eidx * FE_ENT_WORDS);
        const float4 w1 = *reinterpret_cast<const float4*>(s_ent + eidx * FE_ENT_WORDS + 4);
        float acc = pw[0] * w0.x;
        acc = fmaf(pw[1], w0.y, acc); acc = fmaf(pw[2], w0.z, acc); acc = fmaf(pw[3], w0.w, acc);
        acc = fmaf(pw[4], w1.x, acc); acc = fmaf(pw[5], w1.y, acc); acc = fmaf(pw[6], w1.z, acc); acc = fmaf(pw[7], w1.w, acc);
        part[eidx] = acc;
      }
    } else {
      // banks with more than FE_ENT_SMEM (block, filter) pairs (e.g. 128 mels): the power bins go through shared memory and every
      // filter walks its dense column
#pragma unroll
      for (int i = 0; i < 8; ++i) part[8 * blk + i] = pw[i];
      if (lane == 0) part[256] = nyq;
    }
    __syncwarp();
    for (int m = lane; m < p.M; m += 32) {
      float acc;
      if (small_bank) {
        acc = nyq0 * s_nyw[m];
        for (int k = s_foff[m]; k < s_foff[m + 1]; ++k) acc += part[s_fidx[k]];
      } else {
        acc = 0.f;
        for (int j = 0; j < HOWL_NFREQ; ++j) acc = fmaf(part[j], __ldg(p.fb + (size_t)j * p.M + m), acc);
      }
      float val = logf(acc + FE_LOG_EPS);
      if (do_zmuv) val = __fdiv_rn(val - p.zmean, p.zstd);
      if (masked && ((m >= rf0 && m < rf0 + rfl) || (f >= rt0 && f < rt0 + rtl))) val = 0.f;
      s_res[fi * p.M + m] = val;
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
__global__ void deltas_kernel(float* __restrict__ out, const float* __restrict__ src, int64_t rows, int M, int F, float zmean, float zstd,
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
    const float* x0 = src ? src + row * (int64_t)F : c0;     // deltas_only: the log-mels come from the caller's tensor
    for (int t = lane; t < F; t += 32) xs[t] = x0[t];
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
// scratch of the block-entry filterbank, allocated once by howl_b200_create (no allocation on the call path)
int howl_fe_alloc_scratch(howl_ctx_t* ctx) {
  if (cudaMalloc(&ctx->fe_bank, sizeof(FeBank)) != cudaSuccess) return HOWL_E_CUDA;
  if (cudaMalloc(&ctx->fe_ent, sizeof(float) * FE_ENT_WORDS * 32 * HOWL_MAX_MELS) != cudaSuccess) return HOWL_E_CUDA;
  return HOWL_OK;
}

size_t howl_fe_smem_bytes(int hop, int M) {
  const int span_cap = (FE_TILE - 1) * hop + HOWL_NFFT + 8;
  size_t b = 0;
  b += sizeof(float) * ((span_cap + 3) & ~3);
  b += sizeof(float) * HOWL_NFFT;
  b += sizeof(float2) * 256;
  b += sizeof(float) * FE_ENT_SMEM * FE_ENT_WORDS;
  b += sizeof(float) * FE_WARPS * FE_PART;
  b += sizeof(float) * FE_TILE * M;
  b += sizeof(int) * (FE_ENT_SMEM + HOWL_MAX_MELS + 1) + sizeof(float) * HOWL_MAX_MELS;
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
  // the compact bank + balanced plan are rebuilt unless the caller promised (one-shot option "fb_unchanged") that this call's
  // filterbank is the previous call's -- the standard bank is constant between VTLP draws
  if (!(ctx->fb_plan_valid && ctx->fb_same_next)) {
    fb_compact_kernel<<<1, 1024, 0, st>>>(fb, M, reinterpret_cast<FeBank*>(ctx->fe_bank), ctx->fe_ent);
    HOWL_LAUNCHED(ctx, "fb_compact");
    ctx->fb_plan_valid = 1;
  }
  ctx->fb_same_next = 0;

  FeParams p;
  p.pcm = pcm; p.fb = fb; p.bank = reinterpret_cast<const FeBank*>(ctx->fe_bank); p.ent = ctx->fe_ent;
  p.window = ctx->d_window; p.tw_lane = ctx->d_tw_lane; p.tw_stage = ctx->d_tw_stage; p.w512_lane = ctx->d_w512_lane;
  p.rects = rects; p.out = out; p.B = B; p.T = T; p.F = F; p.M = M; p.hop = hop;
  p.zmean = zmuv_mean; p.zstd = zmuv_std; p.flags = flags;
  p.pcm_i16 = ((flags & HOWL_FE_PCM_I16) || ctx->pcm_i16) ? 1 : 0;
  p.use_tma = !p.pcm_i16 && ((T & 3) == 0) && ((reinterpret_cast<uintptr_t>(pcm) & 15) == 0) && ((hop & 3) == 0);
  const size_t smem = howl_fe_smem_bytes(hop, M);
  HOWL_CUDA(ctx, cudaFuncSetAttribute(frontend_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((unsigned)howl_ceil_div(F, FE_TILE), (unsigned)1, 1);
  // grid.y is limited to 65535: walk the batch in slabs
  for (int64_t b0 = 0; b0 < B; b0 += 65535) {
    const int64_t nb = (B - b0 < 65535) ? (B - b0) : 65535;
    FeParams q = p;
    q.pcm = p.pcm_i16 ? reinterpret_cast<const float*>(reinterpret_cast<const short*>(pcm) + b0 * T) : pcm + b0 * T;
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
    deltas_kernel<<<(unsigned)blocks, warps * 32, sm, st>>>(out, nullptr, rows, M, F, zmuv_mean, zmuv_std,
                                                            (flags & HOWL_FE_ZMUV) ? 1 : 0, rects);
    HOWL_LAUNCHED(ctx, "deltas");
  }
  return HOWL_OK;
}

// StandardAudioTransform._execute_op(deltas_only=True) (transform.py:272-280): the caller already holds log-mels x [B, M, F];
// out [B, 3, M, F] = stack(x, deltas(x), deltas(deltas(x))).
extern "C" int howl_b200_deltas_fwd(howl_ctx_t* ctx, void* stream, const float* x, int64_t B, int32_t M, int32_t F, float* out) {
  if (!ctx) return HOWL_E_INVALID;
  HOWL_REQUIRE(ctx, x && out && B >= 0 && M > 0 && F > 0, HOWL_E_INVALID, "deltas_fwd: bad argument");
  HOWL_REQUIRE(ctx, x != out, HOWL_E_INVALID, "deltas_fwd: in place is not supported (the output is three times the input)");
  if (B == 0) return HOWL_OK;
  const int64_t rows = B * (int64_t)M;
  const int warps = 8;
  const size_t sm = sizeof(float) * warps * 2 * (size_t)F;
  HOWL_REQUIRE(ctx, sm <= 200 * 1024, HOWL_E_UNSUPPORTED, "deltas_fwd: %d frames do not fit the delta kernel's row buffers", F);
  HOWL_CUDA(ctx, cudaFuncSetAttribute(deltas_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
  int64_t blocks = howl_ceil_div(rows, warps);
  const int64_t cap = (int64_t)ctx->sm_count * 16;
  if (blocks > cap) blocks = cap;
  deltas_kernel<<<(unsigned)blocks, warps * 32, sm, (cudaStream_t)stream>>>(out, x, rows, M, F, 0.f, 1.f, 0, nullptr);
  HOWL_LAUNCHED(ctx, "deltas");
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
