"""Host-side mirror of howl's transform operators on top of libhowl_b200.so.

Same class names, constructor arguments, call signatures, train/eval behaviour, global-`random` draw order and
state_dict keys as the reference:
  StandardAudioTransform  howl/data/transform/transform.py:234-296
  SpecAugmentTransform    howl/data/transform/transform.py:299-339
  ZmuvTransform           howl/data/transform/operator.py:119-146
The arithmetic runs in K1 (csrc/frontend.cu); this file only draws the host randomness, builds the [257, M]
filterbank with the reference's torch op sequence and passes pointers.
"""
from __future__ import annotations

import math
import random
from dataclasses import dataclass
from typing import Dict, Iterable, Optional, Sequence, Tuple

import torch
import torch.nn as nn

from .runtime import Context
from .settings import SETTINGS
from .trainer import mel_filterbank

_CONTEXTS: Dict[Tuple[int, int], Context] = {}


def get_context(device, n_mels: int) -> Context:
    device = torch.device(device)
    index = device.index if device.index is not None else torch.cuda.current_device()
    key = (index, n_mels)
    if key not in _CONTEXTS:
        s = SETTINGS.audio_transform
        _CONTEXTS[key] = Context(torch.device("cuda", index), n_mels=n_mels, sample_rate=s.sample_rate, n_fft=s.num_fft,
                                 hop=s.hop_length)
    return _CONTEXTS[key]


@dataclass
class AugmentationParameter:
    domain: Sequence[float]
    name: str
    current_value_idx: int = None
    prob: float = 0.75
    enabled: bool = True

    @property
    def magnitude(self):
        return self.domain[self.current_value_idx]


class AugmentModule(nn.Module):
    """Coin-flip dispatcher with the reference's draw order (transform.py:90-97): one `rand.random()` per enabled
    parameter per call, evaluated BEFORE the `self.training` test, from the global `random` unless seeded."""

    def __init__(self, seed: int = None):
        super().__init__()
        self.augment_params = self.default_params
        self.rand = random if seed is None else random.Random(seed)
        self.seed = seed

    @property
    def default_params(self):
        raise NotImplementedError

    def augment(self, param, examples, **kwargs):
        raise NotImplementedError

    def passthrough(self, examples, **kwargs):
        return examples

    def forward(self, x, **kwargs):
        for param in self.augment_params:
            if param.enabled and self.rand.random() < param.prob and self.training:
                x = self.augment(param, x, **kwargs)
            else:
                x = self.passthrough(x, **kwargs)
        return x


def vtlp_filterbank(alpha: float, n_mels: int, sample_rate: int = 16000, n_freqs: int = 257, f_hi: float = 4800) -> torch.Tensor:
    """The warped filterbank of create_vtlp_fb_matrix(training=True) (transform.py:373-410), including its
    in-place sequencing: the `>` mask is evaluated on the already alpha-scaled points."""
    s = sample_rate
    all_freqs = torch.linspace(0, sample_rate // 2, n_freqs)
    m_max = 2595.0 * math.log10(1.0 + (float(sample_rate // 2) / 700.0))
    f_pts = 700.0 * (10 ** (torch.linspace(0.0, m_max, n_mels + 2) / 2595.0) - 1.0)
    thr = f_hi * min(alpha, 1) / alpha
    f_pts = torch.where(f_pts <= thr, f_pts * alpha, f_pts)
    hi = f_pts > thr
    warped = s / 2 - ((s / 2 - f_hi * min(alpha, 1)) / (s / 2 - f_hi * min(alpha, 1) / alpha)) * (s / 2 - f_pts)
    f_pts = torch.where(hi, warped, f_pts)
    f_diff = f_pts[1:] - f_pts[:-1]
    slopes = f_pts.unsqueeze(0) - all_freqs.unsqueeze(1)
    down = (-1.0 * slopes[:, :-2]) / f_diff[:-1]
    up = slopes[:, 2:] / f_diff[1:]
    return torch.clamp(torch.minimum(down, up), min=0.0)


class _SpecShape:
    """Stands in for `spec_transform.win_length / hop_length` that callers of the reference read (transform.py:293-296)."""

    def __init__(self, n_fft: int, hop: int):
        self.win_length, self.hop_length, self.n_fft = n_fft, hop, n_fft


class StandardAudioTransform(AugmentModule):
    def __init__(self):
        super().__init__()
        s = SETTINGS.audio_transform
        if s.use_meyda_spectrogram:
            raise NotImplementedError("USE_MEYDA_SPECTROGRAM is outside the B200 hot path (SURVEY §2 row 23)")
        self.num_mels, self.sample_rate, self.num_fft, self.hop_length = s.num_mels, s.sample_rate, s.num_fft, s.hop_length
        self.spec_transform = _SpecShape(s.num_fft, s.hop_length)
        self.register_buffer("fb", mel_filterbank(s.num_mels, s.sample_rate, s.num_fft // 2 + 1), persistent=False)

    @property
    def default_params(self):
        return (AugmentationParameter([0], "vtlp", 0),)

    @torch.no_grad()
    def _execute(self, fb: torch.Tensor, audio: torch.Tensor, mels_only=False, deltas_only=False):
        if audio.device.type != "cuda":
            raise RuntimeError("howl_b200.StandardAudioTransform needs CUDA tensors (no CPU fallback)")
        ctx = get_context(audio.device, self.num_mels)
        if deltas_only:     # transform.py:275: `audio` already holds log-mels [B, M, F]; only the delta / delta-delta stack is computed
            if audio.dim() != 3:
                raise ValueError("deltas_only expects log-mels of shape [B, M, F]")
            log_mels = audio.contiguous().float()
            return log_mels if mels_only else ctx.deltas(log_mels)
        audio = audio.contiguous() if audio.dtype == torch.int16 else audio.contiguous().float()   # int16 PCM goes to the kernel as is
        if audio.dim() == 1:
            audio = audio.unsqueeze(0)
        return ctx.frontend(audio, fb.to(audio.device), "mels" if mels_only else "stacked")

    def augment(self, param, examples: torch.Tensor, **kwargs):
        if kwargs.get("deltas_only"):         # the VTLP op is never invoked (transform.py:275): no alpha draw
            return self._execute(self.fb, examples, **kwargs)
        alpha = random.random() * 0.2 + 0.9  # VtlpMelScale.forward draws from the GLOBAL random (transform.py:441)
        fb = vtlp_filterbank(alpha, self.num_mels, self.sample_rate, self.num_fft // 2 + 1)
        return self._execute(fb, examples, **kwargs)

    def passthrough(self, examples: torch.Tensor, **kwargs):
        return self._execute(self.fb, examples, **kwargs)

    @torch.no_grad()
    def compute_lengths(self, length: torch.Tensor):
        return (torch.div(length - self.spec_transform.win_length, self.spec_transform.hop_length, rounding_mode="floor") + 1).long()


class SpecAugmentTransform(AugmentModule):
    @property
    def default_params(self):
        return (AugmentationParameter([2, 5, 10, 20, 25], "sa_freq", 2), AugmentationParameter([10, 50, 75, 125, 150], "sa_time", 2))

    def draw_rects(self, batch: int, n_mels: int, n_frames: int) -> torch.Tensor:
        """Replays one forward()'s host draws and returns the rectangles [B,4] (f0, f_len, t0, t_len) without
        touching data -- used by the fused train step, which applies the mask inside K1."""
        rects = torch.zeros(batch, 4, dtype=torch.int32)
        for param in self.augment_params:
            if param.enabled and self.rand.random() < param.prob and self.training:
                self._draw(param, rects, n_mels, n_frames)
        return rects

    def _draw(self, param, rects, n_mels, n_frames):
        for idx in range(rects.size(0)):
            if param.name == "sa_freq":
                f = self.rand.randrange(0, param.magnitude)
                f0 = self.rand.randrange(0, n_mels - f)
                rects[idx, 0], rects[idx, 1] = f0, f
            elif param.name == "sa_time":
                t = self.rand.randrange(0, param.magnitude)
                if n_frames - t <= 0:   # the reference's `except ValueError: continue`
                    continue
                t0 = self.rand.randrange(0, n_frames - t)
                rects[idx, 2], rects[idx, 3] = t0, t
            else:
                raise RuntimeError(f"Invalid parameter name for SpecAugmentTransform: {param.name}")

    @torch.no_grad()
    def augment(self, param, examples: torch.Tensor, **kwargs):
        rects = torch.zeros(examples.size(0), 4, dtype=torch.int32)
        self._draw(param, rects, examples.size(2), examples.size(3))
        ctx = get_context(examples.device, SETTINGS.audio_transform.num_mels)
        if not examples.is_contiguous():
            raise RuntimeError("SpecAugmentTransform works in place and needs a contiguous [B,C,M,F] tensor")
        return ctx.spec_mask(examples, rects.to(examples.device))


class ZmuvTransform(nn.Module):
    def __init__(self):
        super().__init__()
        self.register_buffer("total", torch.zeros(1))
        self.register_buffer("mean", torch.zeros(1))
        self.register_buffer("mean2", torch.zeros(1))

    def update(self, data: torch.Tensor, mask=None):
        """Running sums (operator.py:126-135).  The two reductions run in `sum_sumsq_kernel`; the scalar bookkeeping stays on the
        device (no `.item()`): fitting ZMUV over 2001 clips (train.py:235-237) queues work without a host round trip per clip."""
        with torch.no_grad():
            if data.device.type != "cuda":
                raise RuntimeError("howl_b200.ZmuvTransform needs CUDA tensors (no CPU fallback)")
            dev = data.device
            if mask is not None:
                data = data * mask
                n = mask.sum().double()
            else:
                n = float(data.numel())
            ctx = get_context(dev, SETTINGS.audio_transform.num_mels)
            sums = torch.zeros(2, dtype=torch.float64, device=dev)
            ctx.sum_sumsq(data.contiguous().float(), sums)
            home = self.mean.device
            total = self.total.to(dev).double()
            self.mean = ((sums[0] + self.mean.to(dev).double() * total) / (total + n)).float().to(home)
            self.mean2 = ((sums[1] + self.mean2.to(dev).double() * total) / (total + n)).float().to(home)
            self.total = (total + n).float().to(home)

    def initialize(self, iterable: Iterable[torch.Tensor]):
        for ex in iterable:
            self.update(ex)

    @property
    def std(self):
        return (self.mean2 - self.mean ** 2).sqrt()

    def constants(self) -> Tuple[float, float]:
        """(mean, std) as host floats for the kernel arguments, read back ONCE per change of the buffers (update / load_state_dict /
        .to()), not per forward call."""
        key = (id(self.mean), self.mean._version, id(self.mean2), self.mean2._version)
        cached = getattr(self, "_const_cache", None)
        if cached is None or cached[0] != key:
            pair = torch.stack([self.mean.reshape(()), self.std.reshape(())]).cpu()
            cached = (key, (float(pair[0]), float(pair[1])))
            self._const_cache = cached
        return cached[1]

    def forward(self, x: torch.Tensor):
        if x.device.type != "cuda":
            raise RuntimeError("howl_b200.ZmuvTransform needs CUDA tensors (no CPU fallback)")
        mean, std = self.constants()
        return get_context(x.device, SETTINGS.audio_transform.num_mels).zmuv(x.contiguous().float(), mean, std)
