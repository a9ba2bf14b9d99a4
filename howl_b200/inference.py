"""Wake-word inference on top of the CUDA hot path: the API surface of ``howl.model.inference`` (constructor arguments,
attributes and method names of ``InferenceEngine`` / ``FrameInferenceEngine``, howl/model/inference.py:19-267; used by
``howl_client.py:94-105`` and ``train.py:42-94``), built B200-first rather than mirrored:

* device work is BATCHED -- ``FrameInferenceEngine.infer`` cuts every evaluation window of the clip with one strided view and
  pushes them through the fused frontend kernel and the model as a single batch, instead of one B=1 forward and
  one D2H synchronisation per 63 ms hop; ``ingest_frame`` remains for live streaming (one window at a time);
* the host side is a small ``SequenceDetector`` (posterior smoothing + the label-sequence automaton) that consumes the probability
  rows in time order and reproduces the reference's decisions and ``label_history`` (300 reference traces + the known-answer wavs
  of SURVEY App. B.3 pin it: tests/test_host_logic.py, tests/test_gpu_api.py).

Inside a reference checkout ``plugin.install()`` leaves the reference's own engines in place (they run unmodified on the substituted
transforms / models) and only adds ``infer_batched``.

``context`` is duck-typed: anything with ``num_labels``, ``negative_label``, ``blank_label`` and ``coloring`` (the reference's
``howl.context.InferenceContext`` qualifies; ``SimpleContext`` is a stand-in for standalone use).
"""
from __future__ import annotations

import time
from dataclasses import dataclass
from typing import Any, List, Optional, Sequence, Tuple

import numpy as np
import torch

from .settings import SETTINGS
from .transform import StandardAudioTransform, ZmuvTransform


@dataclass
class SimpleContext:
    """num_labels = len(vocab) + 1 (+1 with a CTC blank), negative label last (howl/context.py:86-97)."""
    num_labels: int
    negative_label: int
    blank_label: int = -1
    coloring: Any = None

    @classmethod
    def for_vocab(cls, vocab, use_blank: bool = False):
        n = len(vocab) + 1 + (1 if use_blank else 0)
        return cls(n, len(vocab), len(vocab) + 1 if use_blank else -1)


def window_starts(num_samples: int, window_ms: float, stride_ms: float, sample_rate: int, drop_incomplete: bool = True) -> Tuple[int, int, List[int]]:
    """Integer window arithmetic of ``howl.utils.audio_utils.stride`` (audio_utils.py:26-49): -> (window samples, hop samples, starts).
    ``int(ms / 1000 * sr)`` truncation is kept bit exact; incomplete trailing windows are dropped (or kept, shorter)."""
    size = int(window_ms / 1000 * sample_rate)
    hop = int(stride_ms / 1000 * sample_rate)
    starts, pos = [], 0
    while pos < num_samples:
        if pos + size > num_samples and drop_incomplete:
            break
        starts.append(pos)
        pos += hop
    return size, hop, starts


def stride(audio_data: torch.Tensor, window_ms: int, stride_ms: int, sample_rate: int, drop_incomplete: bool = True):
    """Generator form (same windows as the reference helper), for callers that iterate."""
    size, _, starts = window_starts(audio_data.size(-1), window_ms, stride_ms, sample_rate, drop_incomplete)
    for s in starts:
        yield audio_data[..., s:s + size]


def _expire(history: list, now: float, horizon: float) -> list:
    """Drop the leading entries older than ``horizon`` (only a prefix is dropped -- entries are time ordered)."""
    k = 0
    while k < len(history) and now - history[k][0] > horizon:
        k += 1
    return history[k:] if k else history


class SequenceDetector:
    """Host half of the engines: smoothed label decisions and the target-sequence automaton over ``label_history``.

    Automaton (howl/model/inference.py:84-131): walk the labels younger than ``inference_window_ms``; advance on the next expected
    label (detect when the last one is reached), hold while the label just matched repeats, and fall back to the start once the gap
    since the last useful label exceeds ``tolerance_window_ms``."""

    def _smoothed_label(self, now: float) -> int:
        self.pred_history = _expire(self.pred_history, now, self.smoothing_window_ms)
        envelope = np.maximum.reduce([p for _, p in self.pred_history])
        label = int(envelope.argmax())
        confident = envelope[label] >= self.threshold
        if self.coloring:
            label = self.coloring.color_map.get(label, self.negative_label)
        if not confident:
            label = self.negative_label
        self.label_history.append((now, label))
        return label

    def append_label(self, label: int, curr_time: float = None):
        self.label_history.append((self._now(curr_time), label))

    def _now(self, curr_time):
        return self.time_provider() * 1000 if curr_time is None else curr_time

    def sequence_present(self, curr_time: float = None) -> bool:
        if not self.sequence:
            return False
        now = self._now(curr_time)
        self.label_history = _expire(self.label_history, now, self.inference_window_ms)
        want, matched, fresh = 0, None, 0
        for stamp, label in self.label_history:
            if label == self.sequence[want]:
                want += 1
                if want == len(self.sequence):
                    return True
                matched, fresh = label, stamp
            elif label == matched:
                fresh = stamp
            elif stamp - fresh > self.tolerance_window_ms:
                want, matched, fresh = 0, None, 0
        return False

    def _append_probability_frame(self, prediction: np.ndarray, curr_time: float = None) -> int:
        now = self._now(curr_time)
        self.pred_history.append((now, prediction))
        return self._smoothed_label(now)

    _get_prediction = _smoothed_label   # reference name


class InferenceEngine(SequenceDetector):
    def __init__(self, model, zmuv_transform: ZmuvTransform, context, time_provider=time.time):
        cfg = SETTINGS.inference_engine
        self.model, self.zmuv, self.context, self.settings, self.time_provider = model, zmuv_transform, context, cfg, time_provider
        self.std = StandardAudioTransform().eval()
        self.coloring = context.coloring
        self.negative_label = self.coloring.color_map[context.negative_label] if self.coloring else context.negative_label
        self.blank_idx = context.blank_label
        self.sample_rate = SETTINGS.audio.sample_rate
        self.sequence, self.threshold = cfg.inference_sequence, cfg.inference_threshold
        self.inference_window_ms, self.smoothing_window_ms, self.tolerance_window_ms = (
            cfg.inference_window_ms, cfg.smoothing_window_ms, cfg.tolerance_window_ms)
        self.inference_weights = 1
        if cfg.inference_weights:   # per-label prior, missing labels weigh 1
            w = np.ones(context.num_labels)
            w[:len(cfg.inference_weights)] = cfg.inference_weights
            self.inference_weights = w
        self.reset()

    def to(self, device):
        self.model, self.zmuv = self.model.to(device), self.zmuv.to(device)
        return self

    def reset(self):
        self.model.streaming_state = None
        self.curr_time, self.pred_history, self.label_history = 0, [], []

    # ---- device half: log-mel + ZMUV + model for a BATCH of windows -> weighted, renormalised posteriors on the host
    def _posteriors(self, batch: torch.Tensor, lengths: Optional[torch.Tensor]) -> np.ndarray:
        self.std = self.std.to(batch.device)
        logits = self.model(self.zmuv(self.std(batch)), lengths)
        z = logits.double().cpu().numpy()
        p = np.exp(z - z.max(-1, keepdims=True))
        p = (p / p.sum(-1, keepdims=True)).astype(np.float32) * self.inference_weights
        return p / p.sum(-1, keepdims=True)

    @torch.no_grad()
    def infer(self, audio_data: torch.Tensor) -> bool:
        """Sequential models: the whole clip is one forward, one posterior row per output frame (inference.py:178-211);
        blank-argmax frames are skipped, time advances by clip_ms / frames per row."""
        clip_ms = int(audio_data.size(-1) / self.sample_rate * 1000)
        rows = self._posteriors(audio_data.unsqueeze(0), None).squeeze(1)
        tick = clip_ms / len(rows)
        for row in rows:
            self.curr_time += tick
            if int(np.argmax(row)) == self.blank_idx:
                continue
            self._append_probability_frame(row, curr_time=self.curr_time)
            if self.sequence_present(self.curr_time):
                return True
        return False


class FrameInferenceEngine(InferenceEngine):
    def __init__(self, max_window_size_ms: int, eval_stride_size_ms: int, *args):
        super().__init__(*args)
        self.max_window_size_ms, self.eval_stride_size_ms = max_window_size_ms, eval_stride_size_ms

    def _windows(self, audio_data: torch.Tensor) -> Optional[torch.Tensor]:
        """All evaluation windows of the clip as one [n, window] batch (a strided view made contiguous once)."""
        flat = audio_data.reshape(-1)
        size, hop, starts = window_starts(flat.numel(), self.max_window_size_ms, self.eval_stride_size_ms, self.sample_rate)
        if not starts or size < 1000:   # the reference stops at windows shorter than 1000 samples (inference.py:232-233)
            return None
        return flat.unfold(0, size, hop)[:len(starts)].contiguous()

    @torch.no_grad()
    def infer(self, audio_data: torch.Tensor) -> bool:
        """Decision of inference.py:222-244 with the device work batched: every window -> one frontend launch + one model forward,
        then the host automaton consumes the posterior rows in order and stops at the first detection."""
        batch = self._windows(audio_data)
        if batch is None:
            return False
        self.std = self.std.to(batch.device)
        lengths = self.std.compute_lengths(torch.full((batch.size(0),), batch.size(-1), device=batch.device))
        for row in self._posteriors(batch, lengths):
            self._append_probability_frame(row, curr_time=self.curr_time)
            self.curr_time += self.eval_stride_size_ms
            if self.sequence_present(self.curr_time):
                return True
        return False

    infer_batched = infer

    @torch.no_grad()
    def ingest_frame(self, frame: torch.Tensor, curr_time: float = None) -> int:
        """Live streaming: one window in, smoothed label out (inference.py:246-267)."""
        self.std = self.std.to(frame.device)
        lengths = self.std.compute_lengths(torch.tensor([frame.size(-1)], device=frame.device))
        return self._append_probability_frame(self._posteriors(frame.unsqueeze(0), lengths)[0], curr_time=curr_time)


def infer_batched(engine, audio_data: torch.Tensor) -> bool:
    """Add-on for the REFERENCE's ``FrameInferenceEngine`` (after ``plugin.install()``): same decision as ``engine.infer`` with all
    windows of the clip evaluated as one batch on the device."""
    flat = audio_data.reshape(-1)
    size, hop, starts = window_starts(flat.numel(), engine.max_window_size_ms, engine.eval_stride_size_ms, engine.sample_rate)
    if not starts or size < 1000:
        return False
    batch = flat.unfold(0, size, hop)[:len(starts)].contiguous()
    engine.std = engine.std.to(batch.device)
    lengths = engine.std.compute_lengths(torch.full((batch.size(0),), batch.size(-1), device=batch.device))
    with torch.no_grad():
        probs = torch.softmax(engine.model(engine.zmuv(engine.std(batch)), lengths), -1).cpu().numpy()
    for row in probs:
        row = row * engine.inference_weights
        engine._append_probability_frame(row / row.sum(), curr_time=engine.curr_time)
        engine.curr_time += engine.eval_stride_size_ms
        if engine.sequence_present(engine.curr_time):
            return True
    return False
