"""Device-side batchifier: the clips of a dataset stay resident in HBM; per step the host only draws
the plan (which window of which clip, which side is padded, the length-sorted order) with the reference's integer
arithmetic and global-`random` draw order, and one gather kernel builds the padded batch -- no H2D of audio.

Mirrors WakeWordFrameBatchifier.__call__ (howl/data/transform/batchifier.py:56-118), random_slice (operator.py:60-70)
and tensorize_audio_data (operator.py:89-109), including their quirks, which are kept on purpose and flagged:
  * the negative-sample branch builds its intervals around the LABEL VALUES (`timestamp_label_map.values()`), not the
    timestamps (batchifier.py:89-92), and
  * slices the audio with MILLISECOND values used as sample indices (batchifier.py:100-105).
"""
from __future__ import annotations

import random
from dataclasses import dataclass
from typing import Dict, List, Sequence, Tuple

import numpy as np
import torch

from .runtime import Context


@dataclass
class ClipRef:
    """One resident clip: [offset, offset + length) of the concatenated device buffer + its end-timestamp -> label map."""
    offset: int
    length: int
    timestamp_label_map: Dict[float, int]


def _clamped(a: int, b: int, n: int) -> Tuple[int, int]:
    """python slice [a:b] of a length-n sequence with non-negative a, b (as every caller guarantees)."""
    a, b = min(a, n), min(b, n)
    return a, max(b, a)


class DeviceFrameBatchifier:
    def __init__(self, negative_label: int, positive_sample_prob: float = 0.5, window_size_ms: int = 500,
                 sample_rate: int = 16000, positive_delta_ms: int = 150, eps_ms: int = 20, pad_to_window: bool = True):
        self.negative_label, self.positive_sample_prob = negative_label, positive_sample_prob
        self.window_size_ms, self.sample_rate = window_size_ms, sample_rate
        self.positive_delta_ms, self.eps_ms, self.pad_to_window = positive_delta_ms, eps_ms, pad_to_window

    # ------------------------------------------------------------------ host: the plan (integer math must be bit exact)
    def plan(self, clips: Sequence[ClipRef]):
        picked: List[Tuple[int, ClipRef, int, int]] = []   # (label, clip, a, b) in the order the reference appends them
        win = int(self.sample_rate * self.window_size_ms / 1000)
        for ex in clips:
            n = ex.length
            if not ex.timestamp_label_map:
                if n < win:
                    picked.append((self.negative_label, ex, 0, n))
                else:
                    a = random.randint(0, n - win)
                    picked.append((self.negative_label, ex, a, a + win))
                continue
            select_negative = random.random() > self.positive_sample_prob
            if not select_negative:
                end_ms, label = random.choice(list(ex.timestamp_label_map.items()))
                end_ms_rand = end_ms + (random.random() * self.eps_ms)
                b = int((end_ms_rand / 1000) * self.sample_rate)
                a = max(b - int((self.window_size_ms / 1000) * self.sample_rate), 0)
                random.random()                       # the reference's `if random.random() < 0:` consumes a draw
                if b - a < 0:
                    select_negative = True
                else:
                    picked.append((label, ex, *_clamped(a, b, n)))
            if select_negative:
                pos = sorted(((v - self.positive_delta_ms, v + self.positive_delta_ms) for v in ex.timestamp_label_map.values()),
                             key=lambda x: x[0])
                neg, last = [], 0
                for a, b in pos:
                    if last < a:
                        neg.append((last, a))
                    last = b
                neg.append((b, int(n / 16000 * 1000)))
                a, b = random.choice(neg)
                if b - a > self.window_size_ms:
                    a = random.randint(0, int(b - self.window_size_ms))
                    b = a + self.window_size_ms
                picked.append((self.negative_label, ex, *_clamped(int(a), int(b), n)))
        lengths = np.array([b - a for _, _, a, b in picked])
        order = np.argsort(-lengths)                     # the same (unstable) numpy sort the reference calls
        max_length = int(self.window_size_ms / 1000 * self.sample_rate) if self.pad_to_window else int(lengths.max())
        starts, counts, dst, labels = [], [], [], []
        for i in order.tolist():
            label, ex, a, b = picked[i]
            left = random.random() < 0.5                 # rand_append draw, per row in sorted order
            starts.append(ex.offset + a)
            counts.append(b - a)
            dst.append(max_length - (b - a) if left else 0)
            labels.append(label)
        return (np.array(starts, np.int64), np.array(counts, np.int64), np.array(dst, np.int64), np.array(labels, np.int64),
                max_length)

    # ------------------------------------------------------------------ device: one gather kernel
    def __call__(self, ctx: Context, device_clips: torch.Tensor, clips: Sequence[ClipRef]):
        starts, counts, dst, labels, max_length = self.plan(clips)
        dev = ctx.device
        audio = ctx.batch_gather(device_clips, torch.from_numpy(starts).to(dev), torch.from_numpy(counts).to(dev),
                                 torch.from_numpy(dst).to(dev), max_length)
        return audio, torch.from_numpy(labels).to(dev), torch.from_numpy(counts).to(dev)
