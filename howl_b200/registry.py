"""The model registry of `howl.model` (howl/model/base.py:11-37, howl/utils/class_registry.py:6-19) and the flat-buffer helpers the
CUDA-backed modules share."""
from __future__ import annotations

from typing import Any, List

import torch
import torch.nn as nn


class ClassRegistry:
    registered_map = {}

    def __init_subclass__(cls, name: str = None, **kwargs):
        super().__init_subclass__(**kwargs)
        if name is not None:
            cls.registered_map[name] = cls

    @classmethod
    def registered_names(cls) -> List[str]:
        return list(cls.registered_map.keys())

    @classmethod
    def find_registered_class(cls, name: str):
        return cls.registered_map[name]


class RegisteredModel(nn.Module, ClassRegistry):
    registered_map = {}

    def __init__(self, num_labels: int):
        super().__init__()
        self.num_labels = num_labels
        self.is_streaming = False
        self.is_sequential = False

    def streaming(self):
        self.is_streaming = True
        return self

    def static(self):
        self.is_streaming = False
        return self

    def compute_length(self, length: int):
        return length

    @property
    def streaming_state(self) -> Any:
        return None

    @streaming_state.setter
    def streaming_state(self, x: Any):
        pass


def _flatten_into(owner, params, attr: str, device):
    """Make `params` (nn.Parameters, state_dict order) contiguous views of ONE flat fp32 buffer stored at owner.<attr>.
    Re-done whenever something (`.to()`, `load_state_dict` with assign, ...) broke the aliasing; the Parameter objects
    keep their identity, so optimizers built earlier stay valid."""
    flat = getattr(owner, attr)
    n = sum(p.numel() for p in params)
    ok = flat is not None and flat.device == device and flat.numel() == n
    if ok:
        off = 0
        for p in params:
            if p.data_ptr() != flat.data_ptr() + 4 * off or not p.is_contiguous():
                ok = False
                break
            off += p.numel()
    if not ok:
        flat = torch.empty(n, dtype=torch.float32, device=device)
        off = 0
        for p in params:
            flat[off:off + p.numel()].copy_(p.data.reshape(-1))
            p.data = flat[off:off + p.numel()].view(p.shape)
            off += p.numel()
        setattr(owner, attr, flat)
    return flat


def _check_unchanged(ctx, flat):
    """Backward differentiates against the flat parameter buffer: refuse if it moved, and let autograd's saved-tensor version
    check refuse if a parameter was modified in place since the forward (as torch does for its own modules)."""
    if flat.data_ptr() != ctx.flat_ptr:
        raise RuntimeError("howl_b200: the model's parameters were modified (moved) between forward and backward; "
                           "gradients would be computed against the wrong weights")
    _ = ctx.saved_tensors


