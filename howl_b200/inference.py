"""InferenceEngine / FrameInferenceEngine with the reference's constructor, attributes and decision logic
(howl/model/inference.py:19-267); the per-window device work (frontend + ZMUV + model + softmax inputs) is the
CUDA hot path, the label-sequence finite-state machine stays on the host as in the reference.

`context` is duck-typed: anything with `num_labels`, `negative_label`, `blank_label` and `coloring` (the reference's
howl.context.InferenceContext qualifies; `SimpleContext` below is a minimal stand-in).
"""
from __future__ import annotations

import itertools
import time
from dataclasses import dataclass
from typing import Any, Optional

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


def stride(audio_data: torch.Tensor, window_ms: int, stride_ms: int, sample_rate: int, drop_incomplete: bool = True):
    """Sliding windows over the last axis (howl/utils/audio_utils.py:26-49); integer sample arithmetic is bit exact."""
    chunk = int(window_ms / 1000 * sample_rate)
    step = int(stride_ms / 1000 * sample_rate)
    idx = 0
    while idx < audio_data.size(-1):
        sliced = audio_data[..., idx:idx + chunk]
        if sliced.size(-1) != chunk and drop_incomplete:
            return
        yield sliced
        idx += step


class InferenceEngine:
    def __init__(self, model, zmuv_transform: ZmuvTransform, context, time_provider=time.time):
        self.model = model
        self.zmuv = zmuv_transform
        self.std = StandardAudioTransform().eval()
        self.settings = SETTINGS.inference_engine
        self.context = context
        self.inference_weights = 1
        if self.settings.inference_weights:
            pad = context.num_labels - len(self.settings.inference_weights)
            self.inference_weights = np.pad(self.settings.inference_weights, (0, pad), "constant", constant_values=1)
        self.coloring = context.coloring
        self.negative_label = context.negative_label
        if self.coloring:
            self.negative_label = self.coloring.color_map[self.negative_label]
        self.sample_rate = SETTINGS.audio.sample_rate
        self.threshold = self.settings.inference_threshold
        self.inference_window_ms = self.settings.inference_window_ms
        self.smoothing_window_ms = self.settings.smoothing_window_ms
        self.tolerance_window_ms = self.settings.tolerance_window_ms
        self.sequence = self.settings.inference_sequence
        self.blank_idx = context.blank_label
        self.time_provider = time_provider
        self.curr_time = 0
        self.pred_history = []
        self.label_history = []
        self.reset()

    def to(self, device):
        self.model = self.model.to(device)
        self.zmuv = self.zmuv.to(device)
        return self

    def reset(self):
        self.model.streaming_state = None
        self.curr_time = 0
        self.pred_history = []
        self.label_history = []

    def append_label(self, label: int, curr_time: float = None):
        if curr_time is None:
            curr_time = self.time_provider() * 1000
        self.label_history.append((curr_time, label))

    def sequence_present(self, curr_time: float = None) -> bool:
        """Walk label_history (entries younger than inference_window_ms) through the target-sequence automaton:
        advance on the expected label, stay while the current label repeats, restart once the gap since the last
        useful label exceeds tolerance_window_ms."""
        if not self.sequence:
            return False
        if curr_time is None:
            curr_time = self.time_provider() * 1000
        self.label_history = list(itertools.dropwhile(lambda x: curr_time - x[0] > self.inference_window_ms, self.label_history))
        state, holding, last_ok = 0, None, 0
        for stamp, label in self.label_history:
            if label == self.sequence[state]:
                state += 1
                if state == len(self.sequence):
                    return True
                holding, last_ok = self.sequence[state - 1], stamp
            elif label == holding:
                last_ok = stamp
            elif last_ok + self.tolerance_window_ms < stamp:
                state, holding, last_ok = 0, None, 0
        return False

    def _get_prediction(self, curr_time: float) -> int:
        self.pred_history = list(itertools.dropwhile(lambda x: curr_time - x[0] > self.smoothing_window_ms, self.pred_history))
        lattice_max = np.max(np.vstack([t for _, t in self.pred_history]), 0)
        max_label = lattice_max.argmax()
        max_prob = lattice_max[max_label]
        if self.coloring:
            max_label = self.coloring.color_map.get(max_label, self.negative_label)
        if max_prob < self.threshold:
            max_label = self.negative_label
        self.label_history.append((curr_time, max_label))
        return max_label

    def _append_probability_frame(self, prediction: np.ndarray, curr_time: float = None) -> int:
        if curr_time is None:
            curr_time = self.time_provider() * 1000
        self.pred_history.append((curr_time, prediction))
        return self._get_prediction(curr_time)

    def _softmax_host(self, logits: torch.Tensor) -> np.ndarray:
        z = logits.double().cpu().numpy()
        z = z - z.max(-1, keepdims=True)
        e = np.exp(z)
        return (e / e.sum(-1, keepdims=True)).astype(np.float32)

    @torch.no_grad()
    def infer(self, audio_data: torch.Tensor) -> bool:
        """Whole clip as one batch through a sequential model: predictions [frames, 1, labels] (inference.py:178-211)."""
        delta_ms = int(audio_data.size(-1) / self.sample_rate * 1000)
        self.std = self.std.to(audio_data.device)
        transformed = self.zmuv(self.std(audio_data.unsqueeze(0)))
        predictions = self._softmax_host(self.model(transformed, lengths=None)).squeeze(1)
        sequence_present = False
        delta_ms /= len(predictions)
        for prediction in predictions:
            prediction = prediction * self.inference_weights
            prediction = prediction / prediction.sum()
            self.curr_time += delta_ms
            if np.argmax(prediction) == self.blank_idx:
                continue
            self._append_probability_frame(prediction, curr_time=self.curr_time)
            if self.sequence_present(self.curr_time):
                sequence_present = True
                break
        return sequence_present


class FrameInferenceEngine(InferenceEngine):
    def __init__(self, max_window_size_ms: int, eval_stride_size_ms: int, *args):
        super().__init__(*args)
        self.max_window_size_ms, self.eval_stride_size_ms = max_window_size_ms, eval_stride_size_ms

    @torch.no_grad()
    def infer(self, audio_data: torch.Tensor) -> bool:
        sequence_present = False
        for window in stride(audio_data, self.max_window_size_ms, self.eval_stride_size_ms, self.sample_rate):
            if window.size(-1) < 1000:
                break
            self.ingest_frame(window.squeeze(0), self.curr_time)
            self.curr_time += self.eval_stride_size_ms
            if self.sequence_present(self.curr_time):
                sequence_present = True
                break
        return sequence_present

    @torch.no_grad()
    def infer_batched(self, audio_data: torch.Tensor) -> bool:
        """Same decision as infer(), but every 63 ms-hop window of the clip goes through the frontend and the model
        as ONE batch; the host automaton then consumes the probabilities in order."""
        windows = [w.squeeze(0) for w in stride(audio_data, self.max_window_size_ms, self.eval_stride_size_ms, self.sample_rate)
                   if w.size(-1) >= 1000]
        if not windows:
            return False
        batch = torch.stack(windows)
        self.std = self.std.to(batch.device)
        lengths = self.std.compute_lengths(torch.full((batch.size(0),), batch.size(-1), device=batch.device))
        probs = self._softmax_host(self.model(self.zmuv(self.std(batch)), lengths))
        for prediction in probs:
            prediction = prediction * self.inference_weights
            prediction = prediction / prediction.sum()
            self._append_probability_frame(prediction, curr_time=self.curr_time)
            self.curr_time += self.eval_stride_size_ms
            if self.sequence_present(self.curr_time):
                return True
        return False

    @torch.no_grad()
    def ingest_frame(self, frame: torch.Tensor, curr_time: float = None) -> int:
        self.std = self.std.to(frame.device)
        lengths = torch.tensor([frame.size(-1)]).to(frame.device)
        transformed_lengths = self.std.compute_lengths(lengths)
        transformed = self.zmuv(self.std(frame.unsqueeze(0)))
        prediction = self._softmax_host(self.model(transformed, transformed_lengths))[0]
        prediction = prediction * self.inference_weights
        prediction = prediction / prediction.sum()
        return self._append_probability_frame(prediction, curr_time=curr_time)
