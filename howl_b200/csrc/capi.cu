// Context management and host-side integer helpers of libhowl_b200.so.
#include <math.h>
#include <stdlib.h>

#include "common.cuh"
#include "../../include/howl_b200_debug.h"

char g_howl_create_error[512] = "";
int howl_fe_alloc_scratch(howl_ctx_t* ctx);   // frontend.cu

extern "C" int howl_b200_abi_version(void) { return HOWL_B200_ABI_VERSION; }

#define CREATE_FAIL(code, ...)                                              \
  do {                                                                      \
    snprintf(g_howl_create_error, sizeof(g_howl_create_error), __VA_ARGS__); \
    if (ctx) howl_b200_destroy(ctx);                                        \
    return (code);                                                          \
  } while (0)

extern "C" int howl_b200_create(int device, const howl_frontend_cfg* cfg, howl_ctx_t** out_ctx) {
  howl_ctx_t* ctx = nullptr;
  if (!cfg || !out_ctx) CREATE_FAIL(HOWL_E_INVALID, "create: null argument");
  if (cfg->n_fft != HOWL_NFFT) CREATE_FAIL(HOWL_E_UNSUPPORTED, "create: n_fft=%d (only 512 is built)", cfg->n_fft);
  if (cfg->hop <= 0 || cfg->hop > HOWL_NFFT) CREATE_FAIL(HOWL_E_INVALID, "create: hop=%d out of range", cfg->hop);
  if (cfg->hop & 1) CREATE_FAIL(HOWL_E_UNSUPPORTED, "create: odd hop=%d (the frontend reads frames with 64-bit vector loads)", cfg->hop);
  if (cfg->n_mels < 1 || cfg->n_mels > HOWL_MAX_MELS)
    CREATE_FAIL(HOWL_E_UNSUPPORTED, "create: n_mels=%d outside 1..%d", cfg->n_mels, HOWL_MAX_MELS);
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    CREATE_FAIL(HOWL_E_CUDA, "create: no CUDA device (%s); libhowl_b200 has no CPU fallback", cudaGetErrorString(e));
  if (device < 0 || device >= ndev) CREATE_FAIL(HOWL_E_INVALID, "create: device %d of %d", device, ndev);
  if ((e = cudaSetDevice(device)) != cudaSuccess) CREATE_FAIL(HOWL_E_CUDA, "cudaSetDevice: %s", cudaGetErrorString(e));
  cudaDeviceProp prop;
  if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess)
    CREATE_FAIL(HOWL_E_CUDA, "cudaGetDeviceProperties: %s", cudaGetErrorString(e));
  if (prop.major != 10)
    CREATE_FAIL(HOWL_E_UNSUPPORTED, "create: device is sm_%d%d; this library is built for sm_100a only", prop.major,
                prop.minor);
  ctx = (howl_ctx_t*)calloc(1, sizeof(howl_ctx_t));
  if (!ctx) CREATE_FAIL(HOWL_E_INVALID, "create: out of host memory");
  ctx->device = device;
  ctx->sm_count = prop.multiProcessorCount;
  ctx->fe = *cfg;
  ctx->conv_engine = 1;
  ctx->lstm_engine = 2;
  if (const char* e = getenv("HOWL_B200_LSTM_ENGINE")) {   // tuning aid: run unmodified callers on another recurrence engine (same results)
    if (e[0] >= '0' && e[0] <= '2' && e[1] == 0) ctx->lstm_engine = e[0] - '0';
  }
  ctx->tc_prof = nullptr;
  ctx->tc_prof_kind = 0;
  // tables in double, rounded once
  float win[HOWL_NFFT];
  static float2 tw_lane[32 * 8], tw_stage[32 * 4], w512_lane[32 * 8];
  const double two_pi = 6.283185307179586476925286766559;
  auto br5 = [](int x) { return ((x & 1) << 4) | ((x & 2) << 2) | (x & 4) | ((x & 8) >> 2) | ((x & 16) >> 4); };
  for (int n = 0; n < HOWL_NFFT; ++n) win[n] = (float)(0.5 - 0.5 * cos(two_pi * n / HOWL_NFFT));
  for (int L = 0; L < 32; ++L) {
    for (int m2 = 0; m2 < 8; ++m2) {
      tw_lane[L * 8 + m2] = make_float2((float)cos(two_pi * L * m2 / 256), (float)(-sin(two_pi * L * m2 / 256)));
      const int m = m2 + 8 * br5(L);
      w512_lane[L * 8 + m2] = make_float2((float)cos(two_pi * m / 512), (float)(-sin(two_pi * m / 512)));
    }
    for (int s = 0; s < 4; ++s) {
      const int h = 16 >> s;                       // radix-2 DIF stage of span h: the upper lane multiplies by W_(2h)^(L mod h)
      const bool upper = (L & h) != 0;
      const double ang = two_pi * (L % h) / (2 * h);
      tw_stage[L * 4 + s] = upper ? make_float2((float)cos(ang), (float)(-sin(ang))) : make_float2(1.f, 0.f);
    }
  }
#define CK(x)                                                                          \
  if ((e = (x)) != cudaSuccess) CREATE_FAIL(HOWL_E_CUDA, "%s: %s", #x, cudaGetErrorString(e))
  CK(cudaMalloc(&ctx->d_window, sizeof(win)));
  CK(cudaMalloc(&ctx->d_tw_lane, sizeof(tw_lane)));
  CK(cudaMalloc(&ctx->d_tw_stage, sizeof(tw_stage)));
  CK(cudaMalloc(&ctx->d_w512_lane, sizeof(w512_lane)));
  CK(cudaMemcpy(ctx->d_window, win, sizeof(win), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(ctx->d_tw_lane, tw_lane, sizeof(tw_lane), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(ctx->d_tw_stage, tw_stage, sizeof(tw_stage), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(ctx->d_w512_lane, w512_lane, sizeof(w512_lane), cudaMemcpyHostToDevice));
#undef CK
  if (howl_fe_alloc_scratch(ctx) != HOWL_OK) CREATE_FAIL(HOWL_E_CUDA, "create: cudaMalloc of the filterbank scratch failed");
  *out_ctx = ctx;
  return HOWL_OK;
}

extern "C" int howl_b200_set_option(howl_ctx_t* ctx, const char* name, int64_t value) {
  if (!ctx || !name) return HOWL_E_INVALID;
  if (strcmp(name, "conv_engine") == 0) {
    HOWL_REQUIRE(ctx, value >= 0 && value <= 2, HOWL_E_INVALID, "set_option: conv_engine must be 0 (fp32), 1 (tcgen05, split bf16) or 2 (tcgen05, single bf16)");
    ctx->conv_engine = (int)value;
    return HOWL_OK;
  }
  if (strcmp(name, "lstm_engine") == 0) {
    HOWL_REQUIRE(ctx, value >= 0 && value <= 2, HOWL_E_INVALID, "set_option: lstm_engine must be 0 (plain), 1 (software pipelined) or 2 (pipelined, 2 x 8 register tile)");
    ctx->lstm_engine = (int)value;
    return HOWL_OK;
  }
  if (strcmp(name, "pcm_i16") == 0) {
    ctx->pcm_i16 = value != 0;
    return HOWL_OK;
  }
  if (strcmp(name, "fb_unchanged") == 0) {   // one-shot: the next frontend call's filterbank equals the previous call's
    ctx->fb_same_next = value != 0;
    return HOWL_OK;
  }
  HOWL_SET_ERR(ctx, "set_option: unknown option '%s'", name);
  return HOWL_E_INVALID;
}

extern "C" int howl_b200_profile_begin(howl_ctx_t* ctx, void* stream) {
  if (!ctx) return HOWL_E_INVALID;
  if (!ctx->prof_ev[0]) {
    for (int i = 0; i <= HOWL_PROF_CAP; ++i) HOWL_CUDA(ctx, cudaEventCreate(&ctx->prof_ev[i]));
  }
  ctx->prof_stream = (cudaStream_t)stream;
  ctx->prof_n = 0;
  ctx->prof_on = 1;
  HOWL_CUDA(ctx, cudaEventRecord(ctx->prof_ev[0], ctx->prof_stream));
  return HOWL_OK;
}

extern "C" int howl_b200_profile_end(howl_ctx_t* ctx, char* names, size_t names_bytes, float* ms, int32_t cap) {
  if (!ctx || !ms || !names) return HOWL_E_INVALID;
  ctx->prof_on = 0;
  if (ctx->prof_n == 0) return 0;
  HOWL_CUDA(ctx, cudaEventSynchronize(ctx->prof_ev[ctx->prof_n]));
  size_t off = 0;
  int n = ctx->prof_n < cap ? ctx->prof_n : cap;
  for (int i = 0; i < n; ++i) {
    HOWL_CUDA(ctx, cudaEventElapsedTime(&ms[i], ctx->prof_ev[i], ctx->prof_ev[i + 1]));
    const size_t len = strlen(ctx->prof_name[i]) + 1;
    if (off + len > names_bytes) { n = i; break; }
    memcpy(names + off, ctx->prof_name[i], len);
    off += len;
  }
  return n;
}

extern "C" void howl_b200_destroy(howl_ctx_t* ctx) {
  if (!ctx) return;
  if (ctx->prof_ev[0])
    for (int i = 0; i <= HOWL_PROF_CAP; ++i) cudaEventDestroy(ctx->prof_ev[i]);
  cudaFree(ctx->d_window);
  cudaFree(ctx->d_tw_lane);
  cudaFree(ctx->d_tw_stage);
  cudaFree(ctx->d_w512_lane);
  cudaFree(ctx->fe_bank);
  cudaFree(ctx->fe_ent);
  free(ctx);
}

extern "C" const char* howl_b200_last_error(const howl_ctx_t* ctx) { return ctx ? ctx->err : g_howl_create_error; }
extern "C" int howl_b200_sm_count(const howl_ctx_t* ctx) { return ctx ? ctx->sm_count : 0; }
extern "C" int64_t howl_b200_launch_count(const howl_ctx_t* ctx) { return ctx ? ctx->launches : 0; }

extern "C" int64_t howl_b200_num_frames(int64_t num_samples, int32_t hop) {
  if (hop <= 0 || num_samples < 0) return -1;
  return 1 + num_samples / hop;
}

static inline int64_t floor_div(int64_t a, int64_t b) {
  int64_t q = a / b;
  if ((a % b != 0) && ((a < 0) != (b < 0))) --q;
  return q;
}

extern "C" int howl_b200_compute_lengths(const int64_t* lengths, int64_t n, int32_t win, int32_t hop, int64_t* out) {
  if (!lengths || !out || n < 0 || hop <= 0) return HOWL_E_INVALID;
  for (int64_t i = 0; i < n; ++i) out[i] = floor_div(lengths[i] - win, hop) + 1;
  return HOWL_OK;
}

// Tuning aid, not part of the drop-in surface: the stream convolution kernels of the given kind (1 forward, 2 data gradient)
// write their per-CTA pipeline-wait cycle counters to buf[sm_count][16] (device memory); null switches it off.
extern "C" int howl_b200_debug_stream_profile(howl_ctx_t* ctx, void* buf, int32_t kind) {
  if (!ctx) return HOWL_E_INVALID;
  ctx->tc_prof = (unsigned long long*)buf;
  ctx->tc_prof_kind = kind;
  return HOWL_OK;
}
