// Shared declarations for libhowl_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/howl_b200.h"

#define HOWL_NFFT 512
#define HOWL_NFREQ 257
#define HOWL_MAX_MELS 128
#define HOWL_PROF_CAP 1024

struct howl_ctx {
  int device;
  int sm_count;
  howl_frontend_cfg fe;
  // device tables (built in double on the host, rounded once to f32)
  float* d_window;   // [512]  periodic Hann
  float2* d_tw_lane;    // [32][8]  W256^(L * m2): twiddles between the in-register and the cross-lane part of the warp FFT
  float2* d_tw_stage;   // [32][4]  cross-lane stage twiddles of lane L (spans 16, 8, 4, 2; 1 in the lower lane of a pair)
  float2* d_w512_lane;  // [32][8]  W512^(m2 + 8 * br5(L)): real-FFT post-processing twiddles of the bins lane L ends up with
  // block-entry filterbank (frontend.cu: FeBank + entries), rebuilt when the bank changes; stream ordered
  void* fe_bank;
  float* fe_ent;
  int fb_plan_valid; // the compact bank / plan on the device belong to the filterbank of the previous frontend call
  int pcm_i16;       // option "pcm_i16": every PCM pointer handed to this context is int16 (the on-disk format), scaled by 1 / 32768 in K1
  int fb_same_next;  // one-shot promise of the caller (set_option "fb_unchanged"): the next call's fb equals the previous call's
  int64_t launches;
  int conv_engine;
  int lstm_engine;   // option "lstm_engine": 0 = plain recurrences, 1 = software pipelined, 2 = pipelined with the 2 x 8 register tile (default); same results
  unsigned long long* tc_prof;   // tuning aid (howl_b200_debug_stream_profile): device buffer [sm_count][16] or null
  int tc_prof_kind;              // 1 = forward, 2 = data gradient
  // optional per-launch timing
  int prof_on;
  int prof_n;
  cudaStream_t prof_stream;
  cudaEvent_t prof_ev[HOWL_PROF_CAP + 1];
  const char* prof_name[HOWL_PROF_CAP];
  char err[512];
};

extern char g_howl_create_error[512];

#define HOWL_SET_ERR(ctx, ...)                                    \
  do {                                                            \
    if (ctx) snprintf((ctx)->err, sizeof((ctx)->err), __VA_ARGS__); \
  } while (0)

#define HOWL_REQUIRE(ctx, cond, code, ...) \
  do {                                     \
    if (!(cond)) {                         \
      HOWL_SET_ERR(ctx, __VA_ARGS__);      \
      return (code);                       \
    }                                      \
  } while (0)

#define HOWL_CUDA(ctx, expr)                                                                   \
  do {                                                                                         \
    cudaError_t _e = (expr);                                                                   \
    if (_e != cudaSuccess) {                                                                   \
      HOWL_SET_ERR(ctx, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return HOWL_E_CUDA;                                                                      \
    }                                                                                          \
  } while (0)

// after a kernel launch: count it and surface launch-configuration errors
#define HOWL_LAUNCHED(ctx, name)                                                       \
  do {                                                                                 \
    (ctx)->launches++;                                                                 \
    HOWL_CUDA(ctx, cudaGetLastError());                                                \
    if ((ctx)->prof_on && (ctx)->prof_n < HOWL_PROF_CAP) {                             \
      (ctx)->prof_name[(ctx)->prof_n] = (name);                                        \
      (ctx)->prof_n++;                                                                 \
      HOWL_CUDA(ctx, cudaEventRecord((ctx)->prof_ev[(ctx)->prof_n], (ctx)->prof_stream)); \
    }                                                                                  \
  } while (0)

static inline int64_t howl_ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
static inline size_t howl_align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
