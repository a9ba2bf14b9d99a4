/*
 * howl_b200.h -- C ABI of libhowl_b200.so: the B200 (sm_100a) implementation of castorini/howl's
 * data-parallel hot path (audio frontend -> res8 forward/backward -> AdamW).
 *
 * The reference (howl @ 4ba5f42) is pure Python and has no FFI of its own; each entry
 * point below therefore cites the reference *call site* it replaces.  Conventions:
 *   - extern "C", plain pointers and sizes, no torch / C++ types;
 *   - every tensor argument is a caller-owned, contiguous DEVICE pointer unless marked "host";
 *   - calls are asynchronous on the cudaStream_t passed as `void* stream`; no hidden synchronisation;
 *   - return 0 (HOWL_OK) or a negative HOWL_E_* code; howl_b200_last_error() gives the message;
 *   - a context is bound to one device and is not thread-safe.
 * INTEGRATION.md shows the ctypes binding a howl maintainer would add.
 */
#ifndef HOWL_B200_H_
#define HOWL_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HOWL_B200_ABI_VERSION 1

enum {
  HOWL_OK = 0,
  HOWL_E_INVALID = -1,     /* bad argument (null pointer, size out of range, T <= n_fft/2, ...) */
  HOWL_E_CUDA = -2,        /* a CUDA runtime call or kernel launch failed */
  HOWL_E_UNSUPPORTED = -3, /* configuration outside what the kernels are built for */
  HOWL_E_WORKSPACE = -4    /* caller-supplied workspace too small */
};

/* frontend output layouts / options (flags argument of howl_b200_frontend_fwd) */
enum {
  HOWL_FE_TIME_MAJOR = 0x1, /* out = [B, F, M] log-mel only: the [B,1,F,M] tensor Res8.forward builds at cnn.py:128-129 */
  HOWL_FE_MELS_ONLY = 0x2,  /* out = [B, M, F]   (StandardAudioTransform(..., mels_only=True), transform.py:276-277) */
  HOWL_FE_STACKED = 0x4,    /* out = [B, 3, M, F] log-mel, delta, delta-delta (transform.py:278-280) */
  HOWL_FE_ZMUV = 0x10,      /* apply (x - mean) / std afterwards (ZmuvTransform.forward, operator.py:145-146) */
  HOWL_FE_PCM_I16 = 0x40    /* `pcm` points to int16 samples (the wav files' own format); K1 converts with x / 32768 exactly as the
                               reference's loader does (howl/utils/audio_utils.py / soundfile), halving the bytes shipped and read */
};

typedef struct howl_ctx howl_ctx_t;

typedef struct {
  int32_t sample_rate; /* 16000 (SETTINGS.audio_transform.sample_rate, settings.py:33) */
  int32_t n_fft;       /* 512 only (settings.py:31) */
  int32_t hop;         /* 200 (settings.py:34) */
  int32_t n_mels;      /* 1..128 (NUM_MELS; settings.py:32) */
} howl_frontend_cfg;

/* ---- context ------------------------------------------------------------------------------- */
int howl_b200_abi_version(void);
/* Replaces: StandardAudioTransform.__init__ (howl/data/transform/transform.py:237-264). */
int howl_b200_create(int device, const howl_frontend_cfg* cfg, howl_ctx_t** out_ctx);
void howl_b200_destroy(howl_ctx_t* ctx);
const char* howl_b200_last_error(const howl_ctx_t* ctx); /* ctx may be NULL: last create() error */
int howl_b200_sm_count(const howl_ctx_t* ctx);
/* number of kernels this context has launched so far (bench.py's gpu_launches) */
int64_t howl_b200_launch_count(const howl_ctx_t* ctx);

/* Options: "conv_engine" = 1 (default, PARITY mode) 45->45 convolutions on tcgen05 tensor cores with bf16x3-split operands
 * and fp32 accumulation (logits within 1e-5 of fp32); 0 = exact-fp32 FFMA kernels; 2 = FAST mode, the same kernels with the
 * low-order bf16 terms skipped (single bf16 x bf16 products, ~3e-3 relative: outside the 1e-4 parity bar, never the default). */
/* "pcm_i16" = 1 (persistent): every `pcm` argument of this context -- frontend_fwd and the fused train steps -- is int16 (see
 * HOWL_FE_PCM_I16).
 * "fb_unchanged" = 1: one-shot promise that the NEXT frontend / train-step call passes the same filterbank contents as the previous
 * one on this context, so the compact bank + work plan built from it are reused instead of rebuilt (cleared by that call).
 * "lstm_engine" = 2 (default) software-pipelined lstm / seq-lstm recurrences with a 2 x 8 register tile; 1 = pipelined, 1 x 16 tile;
 * 0 = the plain kernels.  All three keep the summation order per output: identical results, different speed. */
int howl_b200_set_option(howl_ctx_t* ctx, const char* name, int64_t value);
/* ---- per-launch device timing (CUDA events on the launching stream; used by bench.py's roofline) --------- */
/* After profile_begin every kernel launch of this context is bracketed by an event on `stream`. */
int howl_b200_profile_begin(howl_ctx_t* ctx, void* stream);
/* Synchronises, fills ms[i] with the device time of launch i and names (NUL-separated) with the kernel labels;
 * returns the number of launches recorded (or a negative error) and switches profiling off. */
int howl_b200_profile_end(howl_ctx_t* ctx, char* names, size_t names_bytes, float* ms, int32_t cap);

/* ---- integer frame arithmetic (host; bit exact) ------------------------------------------------ */
/* torch.stft(center=True): 1 + floor(T / hop).  Replaces the implicit frame count of transform.py:249-254. */
int64_t howl_b200_num_frames(int64_t num_samples, int32_t hop);
/* StandardAudioTransform.compute_lengths (transform.py:290-296): floor((len - win) / hop) + 1, host arrays. */
int howl_b200_compute_lengths(const int64_t* lengths, int64_t n, int32_t win, int32_t hop, int64_t* out);

/* ---- K1: fused audio frontend --------------------------------------------------------------- */
/*
 * Replaces: `zmuv_transform(audio_transform(batch.audio_data))` (training/run/train.py:289), i.e.
 * StandardAudioTransform._execute_op (transform.py:271-280) + ZmuvTransform.forward (operator.py:145-146)
 * + SpecAugmentTransform.fmask/tmask (transform.py:310-326).
 *   pcm    [B, T] f32
 *   fb     [n_fft/2+1, n_mels] f32: the mel filterbank (standard or the VTLP matrix drawn for this call)
 *   rects  [B, 4] i32 (f0, f_len, t0, t_len) SpecAugment rectangles drawn on the host, or NULL
 *   out    layout per `flags`
 * T must exceed n_fft/2 (reflect padding, as torch.stft requires).
 */
int howl_b200_frontend_fwd(howl_ctx_t* ctx, void* stream, const float* pcm, int64_t B, int64_t T, const float* fb,
                           float zmuv_mean, float zmuv_std, const int32_t* rects, uint32_t flags, float* out);

/* StandardAudioTransform._execute_op(..., deltas_only=True) (transform.py:272-280): x [B, M, F] already holds log-mels;
 * out [B, 3, M, F] = stack(x, ComputeDeltas(x), ComputeDeltas(ComputeDeltas(x))) (win_length 5, replicate padding).  x != out. */
int howl_b200_deltas_fwd(howl_ctx_t* ctx, void* stream, const float* x, int64_t B, int32_t M, int32_t F, float* out);

/* Sum and sum of squares of n floats into sums[2] (f64, ACCUMULATED onto the existing contents).
 * Replaces the two reductions of ZmuvTransform.update (operator.py:126-135). */
int howl_b200_sum_sumsq(howl_ctx_t* ctx, void* stream, const float* x, int64_t n, double* sums);

/* ZmuvTransform.forward on any tensor of n floats (operator.py:145-146): out = (x - mean) / std; in place allowed. */
int howl_b200_zmuv_fwd(howl_ctx_t* ctx, void* stream, const float* x, int64_t n, float mean, float std, float* out);
/* SpecAugmentTransform.fmask/tmask (transform.py:310-326) in place on x [B,C,M,F] with host-drawn rects [B,4] i32. */
int howl_b200_spec_mask(howl_ctx_t* ctx, void* stream, float* x, int64_t B, int32_t C, int32_t M, int32_t F,
                        const int32_t* rects);
/* x[:, :1].permute(0,1,3,2).contiguous() of Res8.forward (cnn.py:128-129): x [B,C,M,F] -> out [B,F,M]. */
int howl_b200_to_time_major(howl_ctx_t* ctx, void* stream, const float* x, int64_t B, int32_t C, int32_t M, int32_t F,
                            float* out);

/* Data movement of WakeWordFrameBatchifier / tensorize_audio_data (howl/data/transform/batchifier.py:56-118,
 * operator.py:89-109) for a plan drawn on the host: out[r] (max_length samples, zero filled) receives counts[r] samples of
 * `clips` starting at absolute offset starts[r], at column dst_off[r].  All index arrays [B] i64 on the device. */
int howl_b200_batch_gather(howl_ctx_t* ctx, void* stream, const float* clips, const int64_t* starts, const int64_t* counts,
                           const int64_t* dst_off, int64_t B, int64_t max_length, float* out);

/* The same gather with the reference's waveform augmentations (howl/data/transform/transform.py:120-231: TimeshiftTransform,
 * NoiseTransform, DatasetMixer) applied in the same pass, for draws replayed on the host: the time shift is folded into starts / counts;
 * row r is mixed with bg[bg_starts[r] + j] as x * (1 - alpha[r]) + bg * alpha[r] (alpha f64 [B]; bg_starts[r] < 0: not mixed), then white noise of
 * strength sigma[r] and salt-and-pepper noise of probability sp_prob[r] are added with the reference's clamps (0: off); the noise itself
 * comes from an in-kernel Philox4x32-10 keyed by (seed, row, sample).  bg / bg_starts / alpha / sigma / sp_prob may be NULL. */
int howl_b200_batch_gather_aug(howl_ctx_t* ctx, void* stream, const float* clips, const int64_t* starts, const int64_t* counts,
                               const int64_t* dst_off, int64_t B, int64_t max_length, const float* bg, const int64_t* bg_starts,
                               const double* alpha, const float* sigma, const float* sp_prob, uint64_t seed, float* out);

/* ---- res8 ----------------------------------------------------------------------------------- */
/* Flat parameter layout (state_dict order):
 *   conv0.weight[45,1,3,3] | conv1..6.weight[45,45,3,3] | output.weight[L,45] | output.bias[L]
 * BN running statistics: bn_running[6][2][45] f32 (mean, var per layer), num_batches_tracked[6] i64. */
int64_t howl_b200_res8_param_count(int32_t num_labels);
/* Bytes of workspace for a batch of B clips of `frames` x `n_mels` features (train != 0: keeps activations).  About 1.1 MB per
 * 1 s utterance: 10 planar fp32 activation / gradient tensors plus 8 operand-format (bf16 hi, lo) tensors of the tcgen05 engine. */
int64_t howl_b200_res8_workspace_bytes(int64_t B, int32_t frames, int32_t n_mels, int32_t num_labels, int train);

/*
 * Res8.forward (howl/model/cnn.py:127-145) on time-major features [B, F, M] (HOWL_FE_TIME_MAJOR output).
 * train != 0: batch statistics + running-stat update (nn.BatchNorm2d(affine=False) semantics) and the
 * activations are kept in `workspace` for howl_b200_res8_bwd.  Writes logits [B, L].
 */
int howl_b200_res8_fwd(howl_ctx_t* ctx, void* stream, const float* feats, int64_t B, int32_t frames, int32_t n_mels,
                       int32_t num_labels, const float* params, float* bn_running, int64_t* num_batches_tracked,
                       int train, float* logits, void* workspace, size_t workspace_bytes);

/*
 * CrossEntropyLoss(mean) + loss.backward() (training/run/train.py:293,299-301) for the forward kept in
 * `workspace`.  labels [B] i64.  `loss_scale_batch` is the batch size the mean is taken over (B, or the
 * global batch under data parallelism so that an allreduce(SUM) of `grads` yields the global-mean gradient).
 * grads (flat layout) is OVERWRITTEN; loss[1] f32 receives sum_b nll_b / loss_scale_batch.
 */
int howl_b200_res8_bwd(howl_ctx_t* ctx, void* stream, const float* feats, const int64_t* labels, int64_t B,
                       int32_t frames, int32_t n_mels, int32_t num_labels, int64_t loss_scale_batch,
                       const float* params, float* grads, float* loss, void* workspace, size_t workspace_bytes);

/* Same backward for an arbitrary upstream gradient dlogits [B, L] (autograd of Res8.forward when the loss is
 * computed by the caller, e.g. nn.CrossEntropyLoss / a custom criterion at training/run/train.py:250-253). */
int howl_b200_res8_bwd_dlogits(howl_ctx_t* ctx, void* stream, const float* feats, const float* dlogits, int64_t B,
                               int32_t frames, int32_t n_mels, int32_t num_labels, const float* params, float* grads,
                               void* workspace, size_t workspace_bytes);

/* ---- K5: lstm / seq-lstm ------------------------------------------------------------------------- */
/* Flat parameter layout (state_dict order): lstm.weight_ih_l0[512,M] | lstm.weight_hh_l0[512,128] |
 * lstm.bias_ih_l0[512] | lstm.bias_hh_l0[512] | dnn.0.weight[256,128] | dnn.0.bias[256] | dnn.2.weight[L,256] | dnn.2.bias[L]. */
int64_t howl_b200_lstm_param_count(int32_t num_labels, int32_t n_mels);
int64_t howl_b200_lstm_workspace_bytes(int64_t B, int32_t max_steps, int32_t n_mels, int32_t num_labels, int train,
                                       int sequential);
/*
 * SimpleLstm.forward (howl/model/rnn.py:85-91; sequential == 0, out [B, L] = dnn(h_n)) and SequentialLstm.forward
 * (rnn.py:60-71; sequential != 0, out [max_steps, B, L], zero rows beyond each length) on time-major features [B, F, M].
 * lengths [B] i64 (device) = frames each sequence advances (pack_padded_sequence semantics, any order);
 * max_steps = max(lengths) (host).  state_in / state_out: [2][B][128] (h, c) streaming state or NULL (rnn.py:62-68).
 * train != 0 keeps the activations in `workspace` for howl_b200_lstm_bwd (frame objective only).
 */
int howl_b200_lstm_fwd(howl_ctx_t* ctx, void* stream, const float* feats, const int64_t* lengths, int64_t B,
                       int32_t frames, int32_t n_mels, int32_t num_labels, int32_t max_steps, const float* params,
                       const float* state_in, float* state_out, int sequential, int train, float* out, void* workspace,
                       size_t workspace_bytes);
/* CrossEntropyLoss(mean) + backward through the MLP head and the recurrence (BPTT) for the forward kept in `workspace`
 * (training/run/train.py:293,299-301 with --model lstm).  grads (flat layout) is OVERWRITTEN. */
int howl_b200_lstm_bwd(howl_ctx_t* ctx, void* stream, const int64_t* lengths, const int64_t* labels, int64_t B,
                       int32_t frames, int32_t n_mels, int32_t num_labels, int32_t max_steps, int64_t loss_scale_batch,
                       const float* params, float* grads, float* loss, void* workspace, size_t workspace_bytes);
/* Backward for an arbitrary upstream gradient: dlogits [B, L] (sequential == 0) or [max_steps, B, L] (sequential != 0,
 * e.g. from torch's F.log_softmax + nn.CTCLoss at training/run/train.py:296-298). */
int howl_b200_lstm_bwd_dlogits(howl_ctx_t* ctx, void* stream, const int64_t* lengths, const float* dlogits, int sequential,
                               int64_t B, int32_t frames, int32_t n_mels, int32_t num_labels, int32_t max_steps,
                               const float* params, float* grads, void* workspace, size_t workspace_bytes);
/* F.log_softmax + nn.CTCLoss(blank) (reduction 'mean') + backward for a sequential forward kept in `workspace`
 * (training/run/train.py:253,294-301 with --model seq-lstm).  targets [B, max_target_len] i64 (padded), target_lengths [B]
 * i64, lengths [B] i64 = frames per sequence (all device).  loss[1] = mean_b(-ln p_b / max(U_b, 1)). */
int howl_b200_lstm_ctc_bwd(howl_ctx_t* ctx, void* stream, const int64_t* lengths, const int64_t* targets,
                           const int64_t* target_lengths, int32_t max_target_len, int32_t blank, int64_t B, int32_t frames,
                           int32_t n_mels, int32_t num_labels, int32_t max_steps, int64_t loss_scale_batch,
                           const float* params, float* grads, float* loss, void* workspace, size_t workspace_bytes);
/* frontend -> seq-lstm (streaming state [2][B][128] read and replaced, or NULL) -> CTC -> BPTT -> AdamW (single device). */
int howl_b200_seq_lstm_ctc_train_step(howl_ctx_t* ctx, void* stream, const float* pcm, const int64_t* targets,
                                      const int64_t* target_lengths, int32_t max_target_len, int32_t blank,
                                      const int64_t* lengths, int64_t B, int64_t T, const float* fb, float zmuv_mean,
                                      float zmuv_std, int32_t num_labels, int32_t max_steps, float* params, float* state,
                                      float* grads, float* exp_avg, float* exp_avg_sq, int64_t step, float lr,
                                      float weight_decay, float* loss, float* scores, void* workspace,
                                      size_t workspace_bytes);
/* frontend -> lstm -> CE -> backward -> AdamW in one call (single device). */
int howl_b200_lstm_train_step(howl_ctx_t* ctx, void* stream, const float* pcm, const int64_t* labels,
                              const int64_t* lengths, int64_t B, int64_t T, const float* fb, float zmuv_mean,
                              float zmuv_std, int32_t num_labels, int32_t max_steps, float* params, float* grads,
                              float* exp_avg, float* exp_avg_sq, int64_t step, float lr, float weight_decay, float* loss,
                              float* logits, void* workspace, size_t workspace_bytes);

/* ---- K6: MobileNetClassifier (MobileNetV2) ------------------------------------------------------- */
/* Replaces MobileNetClassifier.forward (howl/model/cnn.py:15-29: Conv2d(1,3,3,pad=(1,3)) + BatchNorm2d(3) + ReLU + MaxPool2d((1,2)), then
 * torchvision's MobileNetV2 features + Dropout(0.2) + Linear) and its autograd backward.  bf16 activations / gradients, fp32 master
 * weights, fp32 BatchNorm statistics, fp32 accumulation (BASELINE.json configs[2]).
 * Flat parameter layout = the trainable tensors in state_dict order (conv weight [, conv bias], BatchNorm weight, BatchNorm bias per
 * convolution; classifier weight [L,1280], bias [L]); 2,262,338 floats at 30 labels.  BatchNorm running statistics:
 * bn_running [2][bn_channels] (all means, then all variances, layers concatenated in state_dict order), num_batches_tracked [bn_layers]. */
int64_t howl_b200_mobilenet_param_count(int32_t num_labels);
int64_t howl_b200_mobilenet_bn_channels(void);
int64_t howl_b200_mobilenet_bn_layers(void);
int64_t howl_b200_mobilenet_workspace_bytes(int64_t B, int32_t frames, int32_t n_mels, int32_t num_labels);
/* feats: [B, n_mels, frames] f32 log-mel (HOWL_FE_MELS_ONLY layout = x[:, :1] of the stacked features, cnn.py:27).  train != 0: batch
 * statistics + running-stat update, activations kept in `workspace` for the backward; dropout_p / seed drive the classifier's dropout
 * mask (a counter-based hash of (seed, utterance, channel); eval: identity).  Writes logits [B, L]. */
int howl_b200_mobilenet_fwd(howl_ctx_t* ctx, void* stream, const float* feats, int64_t B, int32_t frames, int32_t n_mels,
                            int32_t num_labels, const float* params, float* bn_running, int64_t* num_batches_tracked, int train,
                            float dropout_p, uint64_t seed, float* logits, void* workspace, size_t workspace_bytes);
/* CrossEntropyLoss(mean) + backward (training/run/train.py:293,299-301 with --model mobilenet) for the forward kept in `workspace`;
 * dropout_p / seed as given to the forward.  grads (flat layout) is OVERWRITTEN. */
int howl_b200_mobilenet_bwd(howl_ctx_t* ctx, void* stream, const float* feats, const int64_t* labels, int64_t B, int32_t frames,
                            int32_t n_mels, int32_t num_labels, int64_t loss_scale_batch, const float* params, float* grads,
                            float dropout_p, uint64_t seed, float* loss, void* workspace, size_t workspace_bytes);
int howl_b200_mobilenet_bwd_dlogits(howl_ctx_t* ctx, void* stream, const float* feats, const float* dlogits, int64_t B, int32_t frames,
                                    int32_t n_mels, int32_t num_labels, const float* params, float* grads, float dropout_p,
                                    uint64_t seed, void* workspace, size_t workspace_bytes);
/* frontend -> MobileNetV2 -> CE -> backward -> AdamW in one call (single device). */
int howl_b200_mobilenet_train_step(howl_ctx_t* ctx, void* stream, const float* pcm, const int64_t* labels, int64_t B, int64_t T,
                                   const float* fb, float zmuv_mean, float zmuv_std, int32_t num_labels, float* params,
                                   float* bn_running, int64_t* num_batches_tracked, float* grads, float* exp_avg, float* exp_avg_sq,
                                   int64_t step, float lr, float weight_decay, float dropout_p, uint64_t seed, float* loss,
                                   float* logits, void* workspace, size_t workspace_bytes);

/* ---- K7: LASClassifier -------------------------------------------------------------------------------- */
/* LASClassifier.forward (howl/model/rnn.py:206-215) = LASEncoder (two Conv2d(.., 8, 3, padding=2) + BatchNorm2d + ReLU + MaxPool2d((1,2)) over
 * the three stacked feature channels, bidirectional LSTM(8 * (n_mels + 4) -> 96) over each clip's own length) + FixedAttentionModule (4 heads)
 * + Linear(192, 256) + ReLU + Dropout + Linear(256, L), and its autograd backward; exact fp32.
 * Flat parameter layout = `parameters()` order of the reference module (477,862 floats at 30 labels / 40 mels); bn_running [2 layers][2][8]
 * (mean, var), num_batches_tracked [2].  enc_lengths [B] i64 (device) = howl_b200_las_lengths of the clips' frame counts.
 * train != 0 in workspace_bytes adds the activations the backward needs. */
int64_t howl_b200_las_param_count(int32_t num_labels, int32_t n_mels);
int64_t howl_b200_las_workspace_bytes(int64_t B, int32_t frames, int32_t n_mels, int32_t num_labels, int train);
/* LASEncoder.forward's length arithmetic (rnn.py:163-168; float floors at every step as the reference), host arrays. */
int howl_b200_las_lengths(const int64_t* lengths, int64_t n, int64_t* out);
/* feats [B, 3, n_mels, frames] f32 (HOWL_FE_STACKED layout, normalised).  train != 0: batch statistics + running-stat update, activations
 * kept in `workspace` (sized with train = 1) for the backward; dropout_p / seed drive the fc dropout mask (a counter-based hash of
 * (seed, utterance, unit); eval: identity). */
int howl_b200_las_fwd(howl_ctx_t* ctx, void* stream, const float* feats, const int64_t* enc_lengths, int64_t B, int32_t frames,
                      int32_t n_mels, int32_t num_labels, const float* params, float* bn_running, int64_t* num_batches_tracked, int train,
                      float dropout_p, uint64_t seed, float* logits, void* workspace, size_t workspace_bytes);
/* CrossEntropyLoss(mean) + backward (training/run/train.py:293,299-301 with --model las) of the train-mode forward kept in `workspace`;
 * dropout_p as given to that forward.  grads (flat layout) is OVERWRITTEN; loss may be NULL. */
int howl_b200_las_bwd(howl_ctx_t* ctx, void* stream, const float* feats, const int64_t* enc_lengths, const int64_t* labels, int64_t B,
                      int32_t frames, int32_t n_mels, int32_t num_labels, int64_t loss_scale_batch, const float* params, float* grads,
                      float dropout_p, float* loss, void* workspace, size_t workspace_bytes);
int howl_b200_las_bwd_dlogits(howl_ctx_t* ctx, void* stream, const float* feats, const int64_t* enc_lengths, const float* dlogits,
                              int64_t B, int32_t frames, int32_t n_mels, int32_t num_labels, const float* params, float* grads,
                              float dropout_p, void* workspace, size_t workspace_bytes);

/* ---- K4: fused AdamW over a flat buffer ------------------------------------------------------- */
/* torch.optim.AdamW.step (training/run/train.py:256,302): decoupled weight decay, bias correction,
 * eps outside the sqrt; `step` is 1-based. */
int howl_b200_adamw(howl_ctx_t* ctx, void* stream, float* params, const float* grads, float* exp_avg,
                    float* exp_avg_sq, int64_t n, int64_t step, float lr, float beta1, float beta2, float eps,
                    float weight_decay);

/* ---- whole train step (single device, no collective) ---------------------------------------------- */
/*
 * One iteration of the loop body training/run/train.py:287-302 for res8 / frame objective:
 * frontend -> Res8 -> CE -> backward -> AdamW, PCM in, updated parameters out.
 * Under data parallelism call frontend_fwd + res8_fwd + res8_bwd, allreduce `grads`, then adamw.
 */
int howl_b200_res8_train_step(howl_ctx_t* ctx, void* stream, const float* pcm, const int64_t* labels, int64_t B,
                              int64_t T, const float* fb, float zmuv_mean, float zmuv_std, const int32_t* rects,
                              int32_t num_labels, float* params, float* bn_running, int64_t* num_batches_tracked,
                              float* grads, float* exp_avg, float* exp_avg_sq, int64_t step, float lr,
                              float weight_decay, float* loss, float* logits, void* workspace,
                              size_t workspace_bytes);

#ifdef __cplusplus
}
#endif
#endif /* HOWL_B200_H_ */
