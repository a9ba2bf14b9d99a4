"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the howl hot path.

A restatement (not a copy) of the arithmetic of castorini/howl's per-batch hot path.
Citations are ``file:line`` under ``/root/reference`` (howl @ 4ba5f42); the arithmetic that
lives in third-party code (torchaudio 0.10 pinned / 2.11 here, torch 1.10.1 pinned / 2.11 here)
is restated from its published formulae.

Parity pinning: the reference's own tests hold NO golden vectors for this path,
so this oracle is pinned against outputs of the reference itself, generated in the build
container by ``oracle/make_golden.py`` (imports /root/reference read-only) and committed under
``tests/golden/``.  ``tests/test_oracle_golden.py`` checks every function here against them.

Two flavours of the frontend are provided:
  * ``*_f32``  -- float32 torch-CPU ops in the reference's operation order (the parity oracle);
  * ``*_f64``  -- an independent float64 numpy restatement (tolerance calibration).

Nothing in ``howl_b200`` may import this module.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

# ----------------------------------------------------------------------------------------------
# constants of the reference configuration (howl/settings.py:27-35)
# ----------------------------------------------------------------------------------------------
SAMPLE_RATE = 16000
N_FFT = 512
HOP = 200
N_FREQS = N_FFT // 2 + 1
LOG_EPS = 1e-7  # howl/data/transform/transform.py:275
BN_EPS = 1e-5
BN_MOMENTUM = 0.1
RES8_MAPS = 45  # howl/model/cnn.py:110
RES8_LAYERS = 6
RES8_POOL = (3, 4)  # howl/model/cnn.py:109


# ----------------------------------------------------------------------------------------------
# frame-index arithmetic (integer; must be bit exact)
# ----------------------------------------------------------------------------------------------
def num_frames(num_samples: int, hop: int = HOP) -> int:
    """Frames produced by ``torch.stft(center=True)``: 1 + floor(T / hop)."""
    return 1 + num_samples // hop


def compute_lengths(lengths: np.ndarray, win: int = N_FFT, hop: int = HOP) -> np.ndarray:
    """``StandardAudioTransform.compute_lengths`` -- howl/data/transform/transform.py:290-296.

    floor_div(len - win, hop) + 1 as int64 (python floor semantics for negatives).
    """
    lengths = np.asarray(lengths, dtype=np.int64)
    return np.floor_divide(lengths - win, hop) + 1


# ----------------------------------------------------------------------------------------------
# filterbanks
# ----------------------------------------------------------------------------------------------
def _triangles(all_freqs: torch.Tensor, f_pts: torch.Tensor) -> torch.Tensor:
    f_diff = f_pts[1:] - f_pts[:-1]
    slopes = f_pts.unsqueeze(0) - all_freqs.unsqueeze(1)
    down = (-1.0 * slopes[:, :-2]) / f_diff[:-1]
    up = slopes[:, 2:] / f_diff[1:]
    return torch.clamp(torch.minimum(down, up), min=0.0)


def mel_filterbank(n_mels: int, sample_rate: int = SAMPLE_RATE, n_freqs: int = N_FREQS) -> torch.Tensor:
    """HTK mel triangles ``fb[n_freqs, n_mels]`` of ``torchaudio.transforms.MelSpectrogram`` defaults.

    Reference call site howl/data/transform/transform.py:249-254 (f_min=0, f_max=sr/2, norm=None,
    mel_scale='htk'); formula SURVEY App. A.1 item 4.  float32 torch ops in the published order.
    """
    f_max = float(sample_rate // 2)
    all_freqs = torch.linspace(0, sample_rate // 2, n_freqs)
    m_min = 2595.0 * math.log10(1.0 + 0.0 / 700.0)
    m_max = 2595.0 * math.log10(1.0 + f_max / 700.0)
    m_pts = torch.linspace(m_min, m_max, n_mels + 2)
    f_pts = 700.0 * (10 ** (m_pts / 2595.0) - 1.0)
    return _triangles(all_freqs, f_pts)


def vtlp_filterbank(
    alpha: float, n_mels: int, sample_rate: int = SAMPLE_RATE, n_freqs: int = N_FREQS, f_hi: float = 4800
) -> torch.Tensor:
    """VTLP-warped filterbank -- howl/data/transform/transform.py:373-410 with training=True.

    The reference warps ``f_pts`` in place and evaluates the second mask on the already-scaled
    tensor (``:397-401``); that sequencing is reproduced literally.
    """
    s = sample_rate
    all_freqs = torch.linspace(0, sample_rate // 2, n_freqs)
    m_max = 2595.0 * math.log10(1.0 + (float(sample_rate // 2) / 700.0))
    m_pts = torch.linspace(0.0, m_max, n_mels + 2)
    f_pts = 700.0 * (10 ** (m_pts / 2595.0) - 1.0)
    thr = f_hi * min(alpha, 1) / alpha
    lo = f_pts <= thr
    f_pts = torch.where(lo, f_pts * alpha, f_pts)  # step 1 (scales the low band)
    hi = f_pts > thr  # step 2 mask is taken AFTER step 1
    warped = s / 2 - ((s / 2 - f_hi * min(alpha, 1)) / (s / 2 - f_hi * min(alpha, 1) / alpha)) * (s / 2 - f_pts)
    f_pts = torch.where(hi, warped, f_pts)
    return _triangles(all_freqs, f_pts)


def filterbank_ranges(fb: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """First / one-past-last non-zero frequency row of every filterbank column (host helper)."""
    fb = np.asarray(fb)
    lo = np.zeros(fb.shape[1], dtype=np.int32)
    hi = np.zeros(fb.shape[1], dtype=np.int32)
    for m in range(fb.shape[1]):
        nz = np.nonzero(fb[:, m])[0]
        if nz.size:
            lo[m], hi[m] = nz[0], nz[-1] + 1
    return lo, hi


# ----------------------------------------------------------------------------------------------
# frontend, float32 (parity oracle)
# ----------------------------------------------------------------------------------------------
def power_spectrogram_f32(pcm: torch.Tensor) -> torch.Tensor:
    """|STFT|^2 ``[B, 257, F]`` -- torchaudio ``spectrogram``.

    center=True reflect pad 256, periodic Hann(512), hop 200, onesided, power 2, no normalisation.
    """
    window = torch.hann_window(N_FFT, periodic=True, dtype=torch.float32, device=pcm.device)
    spec = torch.stft(
        pcm.float(), N_FFT, hop_length=HOP, win_length=N_FFT, window=window, center=True,
        pad_mode="reflect", normalized=False, onesided=True, return_complex=True,
    )
    return spec.abs().pow(2.0)


def log_mel_f32(pcm: torch.Tensor, fb: torch.Tensor) -> torch.Tensor:
    """``log(mel + 1e-7)`` ``[B, M, F]`` -- transform.py:275 + torchaudio MelScale (App. A.1 item 5-6)."""
    power = power_spectrogram_f32(pcm)
    mel = torch.matmul(power.transpose(-1, -2), fb).transpose(-1, -2)
    return mel.add(LOG_EPS).log().contiguous()


def deltas_f32(x: torch.Tensor) -> torch.Tensor:
    """``torchaudio.functional.compute_deltas`` (win_length 5, replicate pad) along the last axis.

    d[t] = (-2 x[t-2] - x[t-1] + x[t+1] + 2 x[t+2]) / 10   (SURVEY App. A.1 item 7).
    """
    shape = x.shape
    flat = x.reshape(1, -1, shape[-1])
    padded = F.pad(flat, (2, 2), mode="replicate")
    kernel = torch.arange(-2, 3, dtype=x.dtype, device=x.device).repeat(flat.shape[1], 1, 1)
    out = F.conv1d(padded, kernel, groups=flat.shape[1]) / 10.0
    return out.reshape(shape)


def standard_audio_transform_f32(pcm: torch.Tensor, fb: Optional[torch.Tensor] = None, n_mels: int = 40) -> torch.Tensor:
    """``StandardAudioTransform.forward`` in eval mode (or train mode given the drawn ``fb``).

    howl/data/transform/transform.py:271-280: stack(log-mel, delta, delta-delta) ``[B, 3, M, F]``.
    """
    if fb is None:
        fb = mel_filterbank(n_mels)
    lm = log_mel_f32(pcm, fb)
    d = deltas_f32(lm)
    dd = deltas_f32(d)
    return torch.stack((lm, d, dd), 1)


def zmuv_std(mean: torch.Tensor, mean2: torch.Tensor) -> torch.Tensor:
    """howl/data/transform/operator.py:141-143."""
    return (mean2 - mean ** 2).sqrt()


def zmuv_forward(x: torch.Tensor, mean: torch.Tensor, mean2: torch.Tensor) -> torch.Tensor:
    """howl/data/transform/operator.py:145-146."""
    return (x - mean) / zmuv_std(mean, mean2)


def zmuv_update(total, mean, mean2, data: torch.Tensor):
    """howl/data/transform/operator.py:126-135 (mask=None branch). Returns new (total, mean, mean2)."""
    n = data.numel()
    new_mean = (data.sum() + mean * total) / (total + n)
    new_mean2 = ((data ** 2).sum() + mean2 * total) / (total + n)
    return total + n, new_mean, new_mean2


def spec_augment_apply(x: torch.Tensor, rects: Sequence[Tuple[int, int, int, int]]) -> torch.Tensor:
    """Zero the rectangles drawn by ``SpecAugmentTransform`` (transform.py:310-326).

    ``rects[b] = (f0, f_len, t0, t_len)`` -- the host draws (global ``random``) stay on the host;
    a zero length means "no mask".  In place on ``x[B, C, M, F]``.
    """
    for b, (f0, fl, t0, tl) in enumerate(rects):
        if fl > 0:
            x[b, :, f0:f0 + fl] = 0
        if tl > 0:
            x[b, :, :, t0:t0 + tl] = 0
    return x


# ----------------------------------------------------------------------------------------------
# frontend, float64 numpy (independent restatement)
# ----------------------------------------------------------------------------------------------
def log_mel_f64(pcm: np.ndarray, fb: np.ndarray) -> np.ndarray:
    """float64 restatement of log-mel: explicit reflect pad, framing, Hann, rFFT, |.|^2, fb, log."""
    pcm = np.asarray(pcm, dtype=np.float64)
    fb = np.asarray(fb, dtype=np.float64)
    b, t = pcm.shape
    f = num_frames(t)
    n = np.arange(N_FFT)
    window = 0.5 - 0.5 * np.cos(2.0 * np.pi * n / N_FFT)
    idx = (np.arange(f)[:, None] * HOP - N_FFT // 2) + n[None, :]  # index into the unpadded clip
    idx = np.where(idx < 0, -idx, idx)
    idx = np.where(idx >= t, 2 * (t - 1) - idx, idx)
    frames = pcm[:, idx] * window  # [B, F, 512]
    power = np.abs(np.fft.rfft(frames, axis=-1)) ** 2  # [B, F, 257]
    mel = power @ fb  # [B, F, M]
    return np.log(mel + LOG_EPS).transpose(0, 2, 1)


def deltas_f64(x: np.ndarray) -> np.ndarray:
    xp = np.pad(x, [(0, 0)] * (x.ndim - 1) + [(2, 2)], mode="edge")
    t = x.shape[-1]
    return (-2 * xp[..., 0:t] - xp[..., 1:t + 1] + xp[..., 3:t + 3] + 2 * xp[..., 4:t + 4]) / 10.0


def standard_audio_transform_f64(pcm: np.ndarray, fb: np.ndarray) -> np.ndarray:
    lm = log_mel_f64(pcm, fb)
    d = deltas_f64(lm)
    return np.stack((lm, d, deltas_f64(d)), 1)


# ----------------------------------------------------------------------------------------------
# res8 (howl/model/cnn.py:107-145), functional so that no nn.Module from the reference is needed
# ----------------------------------------------------------------------------------------------
def res8_param_shapes(num_labels: int) -> List[Tuple[str, Tuple[int, ...]]]:
    """state_dict order and shapes of the trainable tensors."""
    shapes = [("conv0.weight", (RES8_MAPS, 1, 3, 3))]
    for i in range(1, RES8_LAYERS + 1):
        shapes.append((f"conv{i}.weight", (RES8_MAPS, RES8_MAPS, 3, 3)))
    shapes.append(("output.weight", (num_labels, RES8_MAPS)))
    shapes.append(("output.bias", (num_labels,)))
    return shapes


def res8_init(num_labels: int, seed: int = 0) -> Dict[str, torch.Tensor]:
    """PyTorch-default init (kaiming_uniform a=sqrt(5) == U(-1/sqrt(fan_in), 1/sqrt(fan_in)))."""
    g = torch.Generator().manual_seed(seed)
    params: Dict[str, torch.Tensor] = {}
    for name, shape in res8_param_shapes(num_labels):
        fan_in = int(np.prod(shape[1:])) if len(shape) > 1 else RES8_MAPS
        bound = 1.0 / math.sqrt(fan_in)
        params[name] = (torch.rand(shape, generator=g) * 2 - 1) * bound
    return params


def res8_bn_init() -> Dict[str, torch.Tensor]:
    stats: Dict[str, torch.Tensor] = {}
    for i in range(1, RES8_LAYERS + 1):
        stats[f"bn{i}.running_mean"] = torch.zeros(RES8_MAPS)
        stats[f"bn{i}.running_var"] = torch.ones(RES8_MAPS)
        stats[f"bn{i}.num_batches_tracked"] = torch.zeros((), dtype=torch.int64)
    return stats


def res8_forward(
    x: torch.Tensor,
    params: Dict[str, torch.Tensor],
    bn: Dict[str, torch.Tensor],
    training: bool,
    taps: Optional[Dict[str, torch.Tensor]] = None,
) -> torch.Tensor:
    """``Res8.forward`` -- howl/model/cnn.py:127-145.  ``x`` is ``[B, C>=1, M, F]`` (frontend layout).

    ``bn`` running stats are updated in place when ``training`` (momentum 0.1, unbiased var),
    exactly as ``nn.BatchNorm2d(affine=False)`` does.  ``taps`` (optional) receives intermediates.
    """
    x = x[:, :1].permute(0, 1, 3, 2).contiguous()  # (time, freq) -- cnn.py:128-129
    old_x = None
    for i in range(RES8_LAYERS + 1):
        y = F.relu(F.conv2d(x, params[f"conv{i}.weight"], None, padding=1))
        if i == 0:
            y = F.avg_pool2d(y, RES8_POOL)
            old_x = y
        if i > 0 and i % 2 == 0:
            x = y + old_x
            old_x = x
        else:
            x = y
        if taps is not None:
            taps[f"u{i}"] = x
        if i > 0:
            if training:
                bn[f"bn{i}.num_batches_tracked"] += 1
            x = F.batch_norm(
                x, bn[f"bn{i}.running_mean"], bn[f"bn{i}.running_var"], None, None,
                training, BN_MOMENTUM, BN_EPS,
            )
    x = x.view(x.size(0), x.size(1), -1).mean(2)
    if taps is not None:
        taps["pooled"] = x
    return F.linear(x, params["output.weight"], params["output.bias"])


def res8_forward_masked(
    x: torch.Tensor,
    params: Dict[str, torch.Tensor],
    mask0: torch.Tensor,
    masks: torch.Tensor,
) -> torch.Tensor:
    """``Res8.forward`` (howl/model/cnn.py:127-145, train-mode BatchNorm) with every ReLU replaced by a GIVEN 0/1 mask:
    ``relu(z) -> z * mask``.  With the masks the implementation under test actually took, d(loss)/d(weights) of this graph is
    the gradient that implementation must produce -- free of the "flip noise" of pre-activations that sit within rounding of
    zero (parity notes), so it can be held to a tight tolerance.

    ``mask0`` [B,45,3H,M]: conv0 pre-activation > 0 on the rows the (3,4) pooling keeps; ``masks`` [6,B,45,H,10]: layers 1..6.
    Any float dtype (use float64).  Running statistics are not touched.
    """
    x = x[:, :1].permute(0, 1, 3, 2).contiguous()
    h3 = mask0.shape[2]
    y = F.conv2d(x, params["conv0.weight"], None, padding=1)[:, :, :h3] * mask0.to(x.dtype)
    x = old_x = F.avg_pool2d(y, RES8_POOL)
    for i in range(1, RES8_LAYERS + 1):
        y = F.conv2d(x, params[f"conv{i}.weight"], None, padding=1) * masks[i - 1].to(x.dtype)
        if i % 2 == 0:
            x = y + old_x
            old_x = x
        else:
            x = y
        x = F.batch_norm(x, None, None, None, None, True, BN_MOMENTUM, BN_EPS)
    x = x.view(x.size(0), x.size(1), -1).mean(2)
    return F.linear(x, params["output.weight"], params["output.bias"])


def adamw_step(
    params: Dict[str, torch.Tensor], grads: Dict[str, torch.Tensor], m: Dict[str, torch.Tensor],
    v: Dict[str, torch.Tensor], step: int, lr: float, weight_decay: float,
    betas: Tuple[float, float] = (0.9, 0.999), eps: float = 1e-8,
) -> None:
    """``torch.optim.AdamW`` single-tensor update restated; in place; ``step`` is 1-based."""
    b1, b2 = betas
    bc1 = 1.0 - b1 ** step
    bc2 = 1.0 - b2 ** step
    for k, p in params.items():
        g = grads[k]
        p.mul_(1.0 - lr * weight_decay)
        m[k].mul_(b1).add_(g, alpha=1.0 - b1)
        v[k].mul_(b2).addcmul_(g, g, value=1.0 - b2)
        denom = (v[k].sqrt() / math.sqrt(bc2)).add_(eps)
        p.addcdiv_(m[k], denom, value=-(lr / bc1))


def res8_train_step(
    feats: torch.Tensor, labels: torch.Tensor, params: Dict[str, torch.Tensor], bn: Dict[str, torch.Tensor],
    m: Dict[str, torch.Tensor], v: Dict[str, torch.Tensor], step: int, lr: float, weight_decay: float,
) -> Tuple[torch.Tensor, torch.Tensor, Dict[str, torch.Tensor]]:
    """One iteration of training/run/train.py:292-302 (frame objective): fwd, CE(mean), bwd, AdamW.

    Returns (loss, logits, grads); ``params``/``bn``/``m``/``v`` are updated in place.
    """
    leaves = {k: p.detach().clone().requires_grad_(True) for k, p in params.items()}
    logits = res8_forward(feats, leaves, bn, training=True)
    loss = F.cross_entropy(logits, labels)
    loss.backward()
    grads = {k: leaves[k].grad.detach() for k in leaves}
    with torch.no_grad():
        adamw_step(params, grads, m, v, step, lr, weight_decay)
    return loss.detach(), logits.detach(), grads


def flatten(tensors: Dict[str, torch.Tensor], num_labels: int) -> torch.Tensor:
    """Concatenate trainable tensors in state_dict order -> the flat layout of include/howl_b200.h."""
    return torch.cat([tensors[name].reshape(-1) for name, _ in res8_param_shapes(num_labels)])


def unflatten(flat: torch.Tensor, num_labels: int) -> Dict[str, torch.Tensor]:
    out, off = {}, 0
    for name, shape in res8_param_shapes(num_labels):
        n = int(np.prod(shape))
        out[name] = flat[off:off + n].reshape(shape).clone()
        off += n
    return out


# ----------------------------------------------------------------------------------------------
# full hot path as the reference runs it (used for the CPU baseline and end-to-end parity)
# ----------------------------------------------------------------------------------------------
def hot_path_features(pcm: torch.Tensor, fb: torch.Tensor, zmean: torch.Tensor, zmean2: torch.Tensor) -> torch.Tensor:
    """``zmuv_transform(audio_transform(batch.audio_data))`` -- training/run/train.py:289."""
    return zmuv_forward(standard_audio_transform_f32(pcm, fb), zmean, zmean2)


def synthetic_batch(batch: int, samples: int, num_labels: int, seed: int = 0):
    """Synthetic inputs of SURVEY §8(d): speech-like RMS noise clamped to [-1, 1], uniform labels."""
    g = torch.Generator().manual_seed(seed)
    pcm = (torch.randn(batch, samples, generator=g) * 0.1).clamp_(-1, 1)
    labels = torch.randint(0, num_labels, (batch,), generator=g)
    return pcm, labels


# ----------------------------------------------------------------------------------------------
# lstm / seq-lstm (howl/model/rnn.py:41-91): nn.LSTM(40 -> 128) over the first `length` frames, MLP(128 -> 256 -> L)
# ----------------------------------------------------------------------------------------------
LSTM_HIDDEN = 128
LSTM_MLP = 256


def lstm_param_shapes(num_labels: int, n_mels: int = 40) -> List[Tuple[str, Tuple[int, ...]]]:
    """state_dict order."""
    h = LSTM_HIDDEN
    return [("lstm.weight_ih_l0", (4 * h, n_mels)), ("lstm.weight_hh_l0", (4 * h, h)), ("lstm.bias_ih_l0", (4 * h,)),
            ("lstm.bias_hh_l0", (4 * h,)), ("dnn.0.weight", (LSTM_MLP, h)), ("dnn.0.bias", (LSTM_MLP,)),
            ("dnn.2.weight", (num_labels, LSTM_MLP)), ("dnn.2.bias", (num_labels,))]


def lstm_init(num_labels: int, seed: int = 0) -> Dict[str, torch.Tensor]:
    """PyTorch defaults: nn.LSTM U(+-1/sqrt(hidden)); nn.Linear U(+-1/sqrt(fan_in))."""
    g = torch.Generator().manual_seed(seed)
    out = {}
    for name, shape in lstm_param_shapes(num_labels):
        if name.startswith("lstm."):
            bound = 1.0 / math.sqrt(LSTM_HIDDEN)
        else:
            bound = 1.0 / math.sqrt(shape[1] if len(shape) > 1 else {"dnn.0.bias": LSTM_HIDDEN, "dnn.2.bias": LSTM_MLP}[name])
        out[name] = (torch.rand(shape, generator=g) * 2 - 1) * bound
    return out


def lstm_recurrence(feats: torch.Tensor, params: Dict[str, torch.Tensor], lengths: torch.Tensor,
                    state: Optional[Tuple[torch.Tensor, torch.Tensor]] = None):
    """The cell of torch.nn.LSTM restated: gates = W_ih x + b_ih + W_hh h + b_hh, rows ordered (i, f, g, o);
    c' = f c + i g; h' = o tanh(c').  `feats` is the frontend layout [B, C>=1, M, F]; as in
    pack_padded_sequence(x.permute(2,0,1), lengths) (rnn.py:65,88) sequence b only advances for t < lengths[b].
    Returns (h_seq [Tmax, B, H] zero beyond each length -- pad_packed_sequence semantics --, (h_n, c_n))."""
    x = feats[:, 0].permute(2, 0, 1)  # [F, B, M]
    b = x.shape[1]
    hdim = LSTM_HIDDEN
    h = state[0].reshape(b, hdim) if state is not None else x.new_zeros(b, hdim)
    c = state[1].reshape(b, hdim) if state is not None else x.new_zeros(b, hdim)
    tmax = int(lengths.max())
    outs = []
    for t in range(tmax):
        gates = F.linear(x[t], params["lstm.weight_ih_l0"], params["lstm.bias_ih_l0"]) + \
            F.linear(h, params["lstm.weight_hh_l0"], params["lstm.bias_hh_l0"])
        i, f, g, o = gates.chunk(4, 1)
        c_new = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(g)
        h_new = torch.sigmoid(o) * torch.tanh(c_new)
        live = (lengths > t).to(x.dtype).unsqueeze(1)
        c = live * c_new + (1 - live) * c
        h = live * h_new + (1 - live) * h
        outs.append(live * h_new)
    return torch.stack(outs), (h, c)


def lstm_forward(feats, params, lengths, sequential: bool = False, state=None):
    """SimpleLstm.forward (rnn.py:85-91): dnn(h_n) -> [B, L];  SequentialLstm.forward (rnn.py:60-71): dnn(h_seq) ->
    [Tmax, B, L] plus the carried state."""
    h_seq, (h, c) = lstm_recurrence(feats, params, lengths, state)
    top = h_seq if sequential else h
    z = F.relu(F.linear(top, params["dnn.0.weight"], params["dnn.0.bias"]))
    out = F.linear(z, params["dnn.2.weight"], params["dnn.2.bias"])
    return (out, (h, c)) if sequential else out


def lstm_flatten(tensors: Dict[str, torch.Tensor], num_labels: int) -> torch.Tensor:
    return torch.cat([tensors[name].reshape(-1) for name, _ in lstm_param_shapes(num_labels)])


def lstm_unflatten(flat: torch.Tensor, num_labels: int) -> Dict[str, torch.Tensor]:
    out, off = {}, 0
    for name, shape in lstm_param_shapes(num_labels):
        n = int(np.prod(shape))
        out[name] = flat[off:off + n].reshape(shape).clone()
        off += n
    return out


def lstm_train_step(feats, labels, lengths, params, m, v, step, lr, weight_decay):
    """Frame-objective iteration with the `lstm` model (training/run/train.py:292-302, pretrain_gsc.py:126-133)."""
    leaves = {k: p.detach().clone().requires_grad_(True) for k, p in params.items()}
    logits = lstm_forward(feats, leaves, lengths)
    loss = F.cross_entropy(logits, labels)
    loss.backward()
    grads = {k: leaves[k].grad.detach() for k in leaves}
    with torch.no_grad():
        adamw_step(params, grads, m, v, step, lr, weight_decay)
    return loss.detach(), logits.detach(), grads


def seq_lstm_ctc_step(feats, targets, target_lengths, lengths, params, state, blank, m, v, step, lr, weight_decay):
    """One iteration of the CTC branch of training/run/train.py:294-302 with the streaming `seq-lstm`:
    scores = model(x, lengths); log_softmax; nn.CTCLoss(blank) (reduction 'mean'); backward; AdamW.
    `state` is the detached (h, c) carried from the previous call (rnn.py:62-68) or None.
    Returns (loss, scores, grads, new_state)."""
    leaves = {k: p.detach().clone().requires_grad_(True) for k, p in params.items()}
    scores, new_state = lstm_forward(feats, leaves, lengths, sequential=True, state=state)
    logp = F.log_softmax(scores, -1)
    loss = F.ctc_loss(logp, targets, lengths, target_lengths, blank=blank, reduction="mean")
    loss.backward()
    grads = {k: leaves[k].grad.detach() for k in leaves}
    with torch.no_grad():
        adamw_step(params, grads, m, v, step, lr, weight_decay)
    return loss.detach(), scores.detach(), grads, (new_state[0].detach(), new_state[1].detach())


# =====================================================================================================
# MobileNetClassifier (howl/model/cnn.py:15-29 on torchvision's MobileNetV2, width 1.0) -- forward restatement for the next
# §8 row (a10).  Functional over a state dict with the reference's key names: `downsample.{0,1}.*` (Conv2d(1,3,3,pad=(1,3)) +
# BatchNorm2d(3) + ReLU + MaxPool2d((1,2))), `model.features.N...` and `model.classifier.1.{weight,bias}`.
# =====================================================================================================
MOBILENET_V2_SETTING = ((1, 16, 1, 1), (6, 24, 2, 2), (6, 32, 3, 2), (6, 64, 4, 2), (6, 96, 3, 1), (6, 160, 3, 2), (6, 320, 1, 1))


def _bn(x, sd, key, train, eps=1e-5):
    """BatchNorm2d with affine parameters; train: biased batch statistics (running stats are not updated here)."""
    if train:
        mean = x.mean((0, 2, 3))
        var = x.var((0, 2, 3), unbiased=False)
    else:
        mean, var = sd[key + ".running_mean"], sd[key + ".running_var"]
    y = (x - mean[None, :, None, None]) / torch.sqrt(var[None, :, None, None] + eps)
    return y * sd[key + ".weight"][None, :, None, None] + sd[key + ".bias"][None, :, None, None]


def mobilenet_plan():
    """[(features index, inp, oup, stride, expand)] of the 17 inverted-residual blocks (torchvision mobilenetv2.py)."""
    plan, inp, idx = [], 32, 1
    for t, c, n, s in MOBILENET_V2_SETTING:
        for i in range(n):
            plan.append((idx, inp, c, s if i == 0 else 1, t))
            inp, idx = c, idx + 1
    return plan


def _ste_bf16(t: torch.Tensor) -> torch.Tensor:
    """Round to bf16 (nearest even) and back, straight-through for autograd."""
    return t + (t.detach().to(torch.bfloat16).to(t.dtype) - t.detach())


def mobilenet_forward(x: torch.Tensor, sd: Dict[str, torch.Tensor], train: bool = False, bf16: bool = False, masks=None) -> torch.Tensor:
    """x: [B, C>=1, 40, F] stacked features (only the log-mel channel is used, cnn.py:27); returns logits [B, L].
    Dropout of the classifier is the identity (eval) -- the training-mode restatement is deterministic up to dropout.

    ``bf16=True`` restates the SAME graph with the storage precision BASELINE.json configs[2] asks for ("bf16 activations / weights with
    fp32 master weights and fp32 BatchNorm statistics"): every tensor that is stored between layers -- the stem's pooled activations,
    every convolution's raw output, the depthwise outputs after BatchNorm + ReLU6, the block outputs -- and the GEMM convolutions'
    weights are rounded to bf16; all arithmetic (accumulation, BatchNorm statistics from the rounded outputs, normalisation) stays fp32.
    With batch-statistics BatchNorm a randomly initialised MobileNetV2 amplifies that rounding to ~20 % of the logits (it is 0.7-1.3 %
    with the shipped GSC checkpoint; tests/test_oracle_golden.py), so the GPU path is held against THIS restatement."""
    import torch.nn.functional as F

    r = _ste_bf16 if bf16 else (lambda t: t)
    relu6 = lambda t: torch.clamp(t, 0.0, 6.0)
    if masks is not None:
        # mask-forced activations (as res8_forward_masked): the forward VALUES are the exact ones, the DERIVATIVES are the given 0/1
        # decisions -- ``masks`` = [stem route [B,3,H,W+4], then one [B,C,H,W] mask per ReLU6 in network order].  With the decisions the
        # implementation under test took, this graph's gradient is the one it must produce, free of activation-boundary flips.
        it = iter(masks[1:])

        def relu6(t):   # noqa: F811
            m = next(it).to(t.dtype)
            return t * m + (torch.clamp(t, 0.0, 6.0) - t * m).detach()
    x = x[:, :1]
    x = F.conv2d(x, sd["downsample.0.weight"], sd["downsample.0.bias"], padding=(1, 3))
    n0 = _bn(x, sd, "downsample.1", train)
    pooled = F.max_pool2d(torch.relu(n0), (1, 2))
    if masks is not None:
        routed = n0 * masks[0].to(n0.dtype)
        wp = routed.shape[-1] // 2
        routed = routed[..., :2 * wp].reshape(*routed.shape[:-1], wp, 2).sum(-1)
        pooled = routed + (pooled - routed).detach()
    x = r(pooled)
    f = "model.features."
    x = relu6(_bn(r(F.conv2d(x, r(sd[f + "0.0.weight"]), None, stride=2, padding=1)), sd, f + "0.1", train))
    for idx, inp, oup, stride, t in mobilenet_plan():
        p, h, j = f"{f}{idx}.conv.", x, 0
        if t != 1:      # pointwise expansion
            h = relu6(_bn(r(F.conv2d(h, r(sd[f"{p}0.0.weight"]))), sd, f"{p}0.1", train))
            j = 1
        hidden = h.shape[1]
        h = r(relu6(_bn(r(F.conv2d(h, sd[f"{p}{j}.0.weight"], None, stride=stride, padding=1, groups=hidden)), sd, f"{p}{j}.1", train)))   # depthwise
        h = _bn(r(F.conv2d(h, r(sd[f"{p}{j + 1}.weight"]))), sd, f"{p}{j + 2}", train)                                                  # linear projection
        x = r(x + h if (stride == 1 and inp == oup) else h)
    x = relu6(_bn(r(F.conv2d(x, r(sd[f + "18.0.weight"]))), sd, f + "18.1", train))
    x = x.mean((2, 3))                                                  # adaptive_avg_pool2d(1) + flatten
    return x @ sd["model.classifier.1.weight"].t() + sd["model.classifier.1.bias"]


def mobilenet_param_names(sd: Dict[str, torch.Tensor]) -> List[str]:
    """Trainable tensors of a MobileNetClassifier state dict, in state_dict order (= the flat layout of include/howl_b200.h)."""
    return [k for k in sd if not (k.endswith("running_mean") or k.endswith("running_var") or k.endswith("num_batches_tracked"))]


def mobilenet_grads(x: torch.Tensor, labels: torch.Tensor, sd: Dict[str, torch.Tensor], dtype=torch.float32, bf16: bool = False, masks=None):
    """CrossEntropyLoss(mean) + autograd through ``mobilenet_forward`` in train mode (batch statistics, dropout off):
    -> (loss, logits, {name: grad}).  ``bf16``: the bf16-storage restatement of the forward (gradients themselves stay fp32)."""
    trainable = set(mobilenet_param_names(sd))
    leaves = {k: (v.detach().clone().to(dtype).requires_grad_(True) if k in trainable else v) for k, v in sd.items()}
    logits = mobilenet_forward(x.to(dtype), leaves, train=True, bf16=bf16, masks=masks)
    loss = F.cross_entropy(logits, labels)
    loss.backward()
    return loss.detach(), logits.detach(), {k: leaves[k].grad.detach().clone() for k in mobilenet_param_names(sd)}


# =====================================================================================================
# LASClassifier (howl/model/rnn.py:133-215) -- forward restatement, groundwork for §8 row a12.  State-dict keys as the reference's:
# encoder.conv1/conv2, encoder.conv_encoder.{1,5} (BatchNorm2d), encoder.lstm_encoder.* (bidirectional, 1 layer), attn.*, fc.{0,3}.
# =====================================================================================================
def _lstm_direction(x: torch.Tensor, w_ih, w_hh, b_ih, b_hh, lengths: torch.Tensor, reverse: bool):
    """One direction of torch.nn.LSTM over a packed batch.  x [T, B, I]; sequence b has lengths[b] valid steps; the reverse
    direction starts at each sequence's own last step.  Returns (outputs [T, B, H] zero beyond the lengths, h_n [B, H])."""
    tmax, b = int(lengths.max()), x.shape[1]
    hdim = w_hh.shape[1]
    h, c = x.new_zeros(b, hdim), x.new_zeros(b, hdim)
    outs = [None] * tmax
    for t in (range(tmax - 1, -1, -1) if reverse else range(tmax)):
        gates = F.linear(x[t], w_ih, b_ih) + F.linear(h, w_hh, b_hh)
        i, f, g, o = gates.chunk(4, 1)
        c_new = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(g)
        h_new = torch.sigmoid(o) * torch.tanh(c_new)
        live = (lengths > t).to(x.dtype).unsqueeze(1)
        c = live * c_new + (1 - live) * c
        h = live * h_new + (1 - live) * h
        outs[t] = live * h_new
    return torch.stack(outs), h


def las_lengths(lengths: torch.Tensor, use_maxpool: bool = True) -> torch.Tensor:
    """LASEncoder.forward's length arithmetic (rnn.py:163-168): two 3-wide convolutions with padding 2 (+2 frames each), each
    followed by MaxPool2d((1, 2)); float floor at every step as the reference does."""
    l = ((lengths.float() - 3 + 4) / 1 + 1).floor()
    if use_maxpool:
        l = (l / 2).floor()
    l = ((l.float() - 3 + 4) / 1 + 1).floor()
    if use_maxpool:
        l = (l / 2).floor()
    return l.long()


def las_forward(x: torch.Tensor, sd: Dict[str, torch.Tensor], lengths: Optional[torch.Tensor] = None, train: bool = False,
                num_heads: int = 4, hid_mask: Optional[torch.Tensor] = None) -> torch.Tensor:
    """x: [B, 3, 40, F] stacked features, lengths: frames per clip, sorted descending (pack_padded_sequence, rnn.py:169).
    Dropout (rnn.py:202) is the identity unless `hid_mask` [B, 256] (keep / (1 - p) factors) is given.  Returns logits [B, L]."""
    if lengths is None:
        lengths = torch.full((x.shape[0],), x.shape[-1], dtype=torch.long)
    e = "encoder."
    h = F.conv2d(x, sd[e + "conv1.weight"], sd[e + "conv1.bias"], padding=2)
    h = F.max_pool2d(torch.relu(_bn(h, sd, e + "conv_encoder.1", train)), (1, 2))
    h = F.conv2d(h, sd[e + "conv2.weight"], sd[e + "conv2.bias"], padding=2)
    h = F.max_pool2d(torch.relu(_bn(h, sd, e + "conv_encoder.5", train)), (1, 2))
    h = h.permute(3, 0, 1, 2).contiguous()
    h = h.view(-1, h.size(1), h.size(2) * h.size(3))                       # [F', B, 8 * 44]
    ll = las_lengths(lengths).to(x.device)
    p = e + "lstm_encoder."
    fwd, _ = _lstm_direction(h, sd[p + "weight_ih_l0"], sd[p + "weight_hh_l0"], sd[p + "bias_ih_l0"], sd[p + "bias_hh_l0"], ll, False)
    bwd, _ = _lstm_direction(h, sd[p + "weight_ih_l0_reverse"], sd[p + "weight_hh_l0_reverse"], sd[p + "bias_ih_l0_reverse"],
                             sd[p + "bias_hh_l0_reverse"], ll, True)
    rnn_seq = torch.cat([fwd, bwd], 2)                                      # [Tmax, B, 192] = pad_packed_sequence output
    mask = (torch.arange(rnn_seq.shape[0], device=rnn_seq.device)[:, None] < ll.to(rnn_seq.device)[None, :]).to(rnn_seq.dtype)
    # FixedAttentionModule.forward (rnn.py:181-191)
    values = F.linear(rnn_seq, sd["attn.v_proj.weight"], sd["attn.v_proj.bias"])
    keys = F.linear(rnn_seq, sd["attn.k_proj.weight"], sd["attn.k_proj.bias"])
    t, b, d = values.shape
    v4 = values.view(t, b, num_heads, d // num_heads)
    k4 = keys.view(t, b, num_heads, d // num_heads)
    cvec = sd["attn.context_vec"].view(-1, num_heads).unsqueeze(-1).expand(-1, -1, t)
    logits = torch.einsum("ijkl,lki->ijk", v4, cvec) + ((1 - mask) * -100).unsqueeze(-1)
    scores = torch.softmax(logits, 0)
    context = torch.einsum("ijk,ijkl->jkl", scores, k4).reshape(b, -1)
    hid = torch.relu(F.linear(context, sd["fc.0.weight"], sd["fc.0.bias"]))
    if hid_mask is not None:
        hid = hid * hid_mask
    return F.linear(hid, sd["fc.3.weight"], sd["fc.3.bias"])


def las_grads(x: torch.Tensor, labels: torch.Tensor, sd: Dict[str, torch.Tensor], lengths: Optional[torch.Tensor] = None,
              dtype=torch.float64, hid_mask: Optional[torch.Tensor] = None):
    """Train-mode (batch statistics) forward + CrossEntropyLoss(mean) + autograd backward (train.py:293,299-301) of `las_forward`.
    Returns (loss, logits, {name: grad}) for the trainable tensors of `sd`."""
    names = [k for k in sd if "running_" not in k and "num_batches" not in k]
    leaves = {k: (v.detach().clone().to(dtype).requires_grad_(True) if k in names else v.detach().clone().to(dtype)) for k, v in sd.items()
              if v.is_floating_point()}
    logits = las_forward(x.to(dtype), leaves, lengths, train=True, hid_mask=None if hid_mask is None else hid_mask.to(dtype))
    loss = F.cross_entropy(logits, labels)
    loss.backward()
    return float(loss.detach()), logits.detach(), {k: leaves[k].grad for k in names}


# =====================================================================================================
# SimpleGru (howl/model/rnn.py:94-130) -- forward restatement.  Keys: conv_encoder.{0,1,4,6}, lstm_encoder.* (a GRU), dnn.{0,3}.
# =====================================================================================================
def gru_forward(x: torch.Tensor, sd: Dict[str, torch.Tensor], lengths: Optional[torch.Tensor] = None, train: bool = False) -> torch.Tensor:
    """x: [B, C>=1, 40, F]; lengths: frames per clip, sorted descending (the reference adds 4 to them IN PLACE -- conv1 pads the
    time axis by 3 on each side -- and halves them for the max-pool, rnn.py:121-124).  torch.nn.GRU cell: r, z, n gate rows;
    n = tanh(W_in x + b_in + r * (W_hn h + b_hn)); h' = (1 - z) * n + z * h.  Returns logits [B, L] from the last valid state."""
    if lengths is None:
        lengths = torch.full((x.shape[0],), x.shape[-1], dtype=torch.long)
    c = "conv_encoder."
    h = F.conv2d(x[:, :1], sd[c + "0.weight"], sd[c + "0.bias"], padding=(1, 3))
    h = F.max_pool2d(torch.relu(_bn(h, sd, c + "1", train)), (1, 2))
    h = torch.relu(F.conv2d(h, sd[c + "4.weight"], sd[c + "4.bias"], padding=1))
    h = _bn(h, sd, c + "6", train).squeeze(1)                             # [B, 40, F']
    ll = ((lengths + 4).float() / 2).floor().long()
    seq = h.permute(2, 0, 1).contiguous()                                   # [F', B, 40]
    p = "lstm_encoder."
    w_ih, w_hh, b_ih, b_hh = sd[p + "weight_ih_l0"], sd[p + "weight_hh_l0"], sd[p + "bias_ih_l0"], sd[p + "bias_hh_l0"]
    state = seq.new_zeros(seq.shape[1], w_hh.shape[1])
    for t in range(int(ll.max())):
        gi, gh = F.linear(seq[t], w_ih, b_ih), F.linear(state, w_hh, b_hh)
        i_r, i_z, i_n = gi.chunk(3, 1)
        h_r, h_z, h_n = gh.chunk(3, 1)
        r, z = torch.sigmoid(i_r + h_r), torch.sigmoid(i_z + h_z)
        n = torch.tanh(i_n + r * h_n)
        new = (1 - z) * n + z * state
        live = (ll > t).to(seq.dtype).unsqueeze(1)
        state = live * new + (1 - live) * state
    hid = torch.relu(F.linear(state, sd["dnn.0.weight"], sd["dnn.0.bias"]))
    return F.linear(hid, sd["dnn.3.weight"], sd["dnn.3.bias"])
