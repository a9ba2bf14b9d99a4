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
        self.last_rows = []                              # per row: (index of the clip in `clips`, window start inside that clip)
        index_of = {id(ex): i for i, ex in enumerate(clips)}
        for i in order.tolist():
            label, ex, a, b = picked[i]
            left = random.random() < 0.5                 # rand_append draw, per row in sorted order
            starts.append(ex.offset + a)
            counts.append(b - a)
            dst.append(max_length - (b - a) if left else 0)
            labels.append(label)
            self.last_rows.append((index_of[id(ex)], a))
        return (np.array(starts, np.int64), np.array(counts, np.int64), np.array(dst, np.int64), np.array(labels, np.int64),
                max_length)

    # ------------------------------------------------------------------ device: one gather kernel
    def __call__(self, ctx: Context, device_clips: torch.Tensor, clips: Sequence[ClipRef]):
        starts, counts, dst, labels, max_length = self.plan(clips)
        dev = ctx.device
        audio = ctx.batch_gather(device_clips, torch.from_numpy(starts).to(dev), torch.from_numpy(counts).to(dev),
                                 torch.from_numpy(dst).to(dev), max_length)
        return audio, torch.from_numpy(labels).to(dev), torch.from_numpy(counts).to(dev)


# =====================================================================================================================
# device-side waveform augmentation
# =====================================================================================================================
@dataclass
class ClipAugmentation:
    """What the reference's augmentation chain did to one clip, as the gather kernel needs it."""
    shift: int = 0            # samples cropped from the left (TimeshiftTransform, `audio[w:]`)
    length: int = 0           # length after the shift
    bg_clip: int = -1         # background clip mixed in (DatasetMixer), -1: none
    bg_start: int = 0         # first background sample that meets sample 0 of the ORIGINAL clip
    alpha: float = 0.0
    sigma: float = 0.0        # white-noise strength (NoiseTransform "white"), 0: off
    sp_prob: float = 0.0      # salt-and-pepper probability, 0: off
    replaced: bool = False    # DatasetMixer "replace": the clip becomes pure background and loses its labels


class DeviceWaveAugmenter:
    """Replays, on the host, the draws of the reference's training augmentation chain
        [DatasetMixer(noise).train()] -> TimeshiftTransform().train() -> NoiseTransform().train() -> batchifier
    (training/run/train.py:202-221; howl/data/transform/transform.py:120-231) from the global ``random`` in the reference's order --
    per transform one coin per parameter (``AugmentModule.forward``), then the per-clip draws -- and turns them into per-row arguments
    of ``howl_b200_batch_gather_aug``, which applies shift, mix and noise while it builds the padded batch from clips resident in HBM.
    The index arithmetic (shift width, background window, mixing weights) is bit exact with the reference; the noise SAMPLES come from
    the kernel's Philox generator (same distributions as the reference's torch host generator, another stream).
    ``TimestretchTransform`` (a phase vocoder driven by numpy's generator) is not part of the device chain: compose without it."""

    MIX_STRENGTH, SHIFT_SECONDS, WHITE, SALT_PEPPER = 0.2, 0.25, 0.001, 1 / 10000     # default magnitudes (domain[current_value_idx])

    def __init__(self, batchifier: DeviceFrameBatchifier, bg_lengths: Sequence[int] = (), do_replace: bool = False, sample_rate: int = 16000,
                 timeshift: bool = True, noise: bool = True):
        self.batchifier, self.bg_lengths, self.do_replace, self.sr = batchifier, list(bg_lengths), do_replace, sample_rate
        self.timeshift, self.noise = timeshift, noise
        self.bg_offsets = np.concatenate([[0], np.cumsum(self.bg_lengths)]).astype(np.int64) if self.bg_lengths else None

    def draw(self, clips: Sequence[ClipRef]) -> List[ClipAugmentation]:
        augs = [ClipAugmentation(length=c.length) for c in clips]
        if self.bg_lengths:        # DatasetMixer: parameters "strength" (p = 0.75) and "replace" (p = 0.1 if do_replace else 0)
            for name, prob in (("strength", 0.75), ("replace", 0.1 if self.do_replace else 0)):
                if not (random.random() < prob):
                    continue
                for c, g in zip(clips, augs):
                    k = random.choice(range(len(self.bg_lengths)))
                    while self.bg_lengths[k] < c.length:
                        k = random.choice(range(len(self.bg_lengths)))
                    b = random.randint(c.length, self.bg_lengths[k])
                    alpha = 1 if name == "replace" else random.random() * self.MIX_STRENGTH
                    if g.bg_clip >= 0:
                        raise NotImplementedError("both DatasetMixer parameters fired for one batch (p = 0.075 with do_replace): "
                                                  "two chained mixes are not fused")
                    g.bg_clip, g.bg_start, g.alpha, g.replaced = k, b - c.length, float(alpha), alpha == 1
        if self.timeshift and random.random() < 0.75:       # TimeshiftTransform
            for c, g in zip(clips, augs):
                w = min(int(random.random() * self.SHIFT_SECONDS * self.sr), int(0.5 * c.length))
                if random.random() < 0.5:
                    g.shift = w
                g.length = c.length - w
        if self.noise:                                      # NoiseTransform: "white", then "salt_pepper"
            if random.random() < 0.75:
                for g in augs:
                    g.sigma = self.WHITE * random.random()
            if random.random() < 0.75:
                for g in augs:
                    g.sp_prob = self.SALT_PEPPER * random.random()
        return augs

    def plan(self, clips: Sequence[ClipRef]):
        """-> (starts, counts, dst_off, labels, max_length, bg_starts, alpha, sigma, sp_prob) as numpy arrays, rows in batch order."""
        augs = self.draw(clips)
        shifted = [ClipRef(c.offset + g.shift, g.length, {} if g.replaced else c.timestamp_label_map) for c, g in zip(clips, augs)]
        starts, counts, dst, labels, max_length = self.batchifier.plan(shifted)
        bg_starts, alpha, sigma, sp = [], [], [], []
        for clip_idx, a in self.batchifier.last_rows:
            g = augs[clip_idx]
            bg_starts.append(int(self.bg_offsets[g.bg_clip]) + g.bg_start + g.shift + a if g.bg_clip >= 0 else -1)
            alpha.append(g.alpha)
            sigma.append(g.sigma)
            sp.append(g.sp_prob)
        return (starts, counts, dst, labels, max_length, np.array(bg_starts, np.int64), np.array(alpha, np.float64),
                np.array(sigma, np.float32), np.array(sp, np.float32))

    def __call__(self, ctx: Context, device_clips: torch.Tensor, clips: Sequence[ClipRef], device_bg: torch.Tensor = None, seed: int = 0):
        starts, counts, dst, labels, max_length, bg_starts, alpha, sigma, sp = self.plan(clips)
        dev = ctx.device
        t = lambda x: torch.from_numpy(x).to(dev)
        audio = ctx.batch_gather_aug(device_clips, t(starts), t(counts), t(dst), max_length, bg=device_bg,
                                     bg_starts=t(bg_starts) if device_bg is not None else None, alpha=t(alpha), sigma=t(sigma), sp_prob=t(sp),
                                     seed=seed)
        return audio, t(labels), t(counts)
