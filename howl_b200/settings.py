"""Environment-driven settings with the reference's variable names (howl/settings.py:27-76).

Only the groups the hot path reads are mirrored: SETTINGS.audio_transform (NUM_FFT, NUM_MELS, SAMPLE_RATE, HOP_LENGTH),
SETTINGS.audio (SAMPLE_RATE) and SETTINGS.inference_engine (INFERENCE_*, SMOOTHING_WINDOW_MS, TOLERANCE_WINDOW_MS).
Values are parsed lazily on first access, as the reference does; `SETTINGS.reset()` re-reads the environment.
"""
from __future__ import annotations

import json
import os
from dataclasses import dataclass, field
from typing import List, Optional


def _env(name: str, default, cast):
    raw = os.environ.get(name.upper(), os.environ.get(name.lower()))
    if raw is None:
        return default
    if cast in (list,):
        return json.loads(raw)
    if cast is bool:
        return raw.strip().lower() in ("1", "true", "yes", "on")
    return cast(raw)


@dataclass
class AudioSettings:
    sample_rate: int = 16000
    use_mono: bool = True


@dataclass
class AudioTransformSettings:
    num_fft: int = 512
    num_mels: int = 80
    sample_rate: int = 16000
    hop_length: int = 200
    use_meyda_spectrogram: bool = False


@dataclass
class InferenceEngineSettings:
    inference_weights: Optional[List[float]] = None
    inference_sequence: List[int] = field(default_factory=lambda: [0])
    inference_window_ms: float = 2000.0
    smoothing_window_ms: float = 50.0
    tolerance_window_ms: float = 500.0
    inference_threshold: float = 0.0


def _load(cls):
    obj = cls()
    for name, default in list(vars(obj).items()):
        cast = list if isinstance(default, list) or name in ("inference_weights", "inference_sequence") else type(default)
        setattr(obj, name, _env(name, default, cast))
    return obj


class HowlSettings:
    def __init__(self):
        self._delegate = None
        self.reset()

    def reset(self):
        self._audio = self._audio_transform = self._inference_engine = None

    def bind(self, external) -> None:
        """Delegate every group to another settings object with the same attributes -- ``plugin.install()`` binds the reference's
        ``howl.settings.SETTINGS`` so that ``Workspace.load_settings()`` (hubconf.py:54, demo.py:27) and direct assignments such as
        ``SETTINGS.inference_engine.inference_sequence = [0, 1, 2]`` reach the CUDA-backed classes.  ``bind(None)`` undoes it."""
        self._delegate = external

    def __getattr__(self, name):
        # groups this mirror does not model (training, dataset, cache, ...) exist only on a bound reference object
        d = self.__dict__.get("_delegate")
        if d is not None and not name.startswith("_"):
            return getattr(d, name)
        raise AttributeError(name)

    @property
    def audio(self) -> AudioSettings:
        if self._delegate is not None:
            return self._delegate.audio
        if self._audio is None:
            self._audio = _load(AudioSettings)
        return self._audio

    @property
    def audio_transform(self) -> AudioTransformSettings:
        if self._delegate is not None:
            return self._delegate.audio_transform
        if self._audio_transform is None:
            self._audio_transform = _load(AudioTransformSettings)
        return self._audio_transform

    @property
    def inference_engine(self) -> InferenceEngineSettings:
        if self._delegate is not None:
            return self._delegate.inference_engine
        if self._inference_engine is None:
            self._inference_engine = _load(InferenceEngineSettings)
        return self._inference_engine


SETTINGS = HowlSettings()
