"""`howl.model` surface on top of libhowl_b200.so: the RegisteredModel registry (howl/model/base.py:11-37,
howl/utils/class_registry.py:6-19) and Res8 (howl/model/cnn.py:107-145) with the reference's state_dict keys.

Res8 keeps all trainable tensors as views into ONE flat fp32 buffer (the layout of include/howl_b200.h) so that
`torch.optim.AdamW(model.parameters())`, `Workspace.save_model` and `load_state_dict` of the reference keep working,
while forward / backward are single calls into the CUDA library.
"""
from __future__ import annotations

import math
from typing import Any, List

import torch
import torch.nn as nn

from .settings import SETTINGS
from .trainer import res8_param_shapes
from .transform import get_context


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


class _Weight(nn.Module):
    def __init__(self, shape, bias_shape=None):
        super().__init__()
        fan_in = math.prod(shape[1:])
        bound = 1.0 / math.sqrt(fan_in)
        self.weight = nn.Parameter((torch.rand(shape) * 2 - 1) * bound)   # == kaiming_uniform_(a=sqrt(5))
        if bias_shape is not None:
            self.bias = nn.Parameter((torch.rand(bias_shape) * 2 - 1) * bound)


class _BatchNormStats(nn.Module):
    def __init__(self, channels: int):
        super().__init__()
        self.register_buffer("running_mean", torch.zeros(channels))
        self.register_buffer("running_var", torch.ones(channels))
        self.register_buffer("num_batches_tracked", torch.tensor(0, dtype=torch.long))


class _Res8Function(torch.autograd.Function):
    @staticmethod
    def forward(ctx, model, feats, labels_hint, *params):
        logits = model._run_forward(feats, train=True)
        ctx.model, ctx.feats = model, feats
        return logits

    @staticmethod
    def backward(ctx, dlogits):
        model = ctx.model
        grads = model._run_backward(ctx.feats, dlogits.contiguous())
        out, off = [], 0
        for _, shape in res8_param_shapes(model.num_labels):
            n = math.prod(shape)
            out.append(grads[off:off + n].view(shape))
            off += n
        return (None, None, None, *out)


class Res8(RegisteredModel, name="res8"):
    N_MAPS = 45

    def __init__(self, num_labels: int, config=None):
        super().__init__(num_labels)
        self.conv0 = _Weight((45, 1, 3, 3))
        for i in range(1, 7):
            setattr(self, f"bn{i}", _BatchNormStats(45))
            setattr(self, f"conv{i}", _Weight((45, 45, 3, 3)))
        self.output = _Weight((num_labels, 45), (num_labels,))
        self._flat = self._bn_flat = self._nbt = self._ws = None

    # ---- flat-buffer bookkeeping -------------------------------------------------------------------
    def _param_list(self):
        ps = [self.conv0.weight] + [getattr(self, f"conv{i}").weight for i in range(1, 7)] + [self.output.weight, self.output.bias]
        return ps

    def _ensure_flat(self, device):
        ps = self._param_list()
        n = sum(p.numel() for p in ps)
        ok = self._flat is not None and self._flat.device == device and self._flat.numel() == n
        if ok:
            off = 0
            for p in ps:
                if p.data_ptr() != self._flat.data_ptr() + 4 * off or not p.is_contiguous():
                    ok = False
                    break
                off += p.numel()
        if not ok:
            flat = torch.empty(n, dtype=torch.float32, device=device)
            off = 0
            for p in ps:
                flat[off:off + p.numel()].copy_(p.data.reshape(-1))
                p.data = flat[off:off + p.numel()].view(p.shape)
                off += p.numel()
            self._flat = flat
        bns = [getattr(self, f"bn{i}") for i in range(1, 7)]
        okb = self._bn_flat is not None and self._bn_flat.device == device
        if okb:
            for i, b in enumerate(bns):
                if (b.running_mean.data_ptr() != self._bn_flat[i, 0].data_ptr() or b.running_var.data_ptr() != self._bn_flat[i, 1].data_ptr()
                        or b.num_batches_tracked.data_ptr() != self._nbt[i].data_ptr()):
                    okb = False
                    break
        if not okb:
            bn_flat = torch.empty(6, 2, 45, dtype=torch.float32, device=device)
            nbt = torch.empty(6, dtype=torch.int64, device=device)
            for i, b in enumerate(bns):
                bn_flat[i, 0].copy_(b.running_mean)
                bn_flat[i, 1].copy_(b.running_var)
                nbt[i].copy_(b.num_batches_tracked)
                b._buffers["running_mean"] = bn_flat[i, 0]
                b._buffers["running_var"] = bn_flat[i, 1]
                b._buffers["num_batches_tracked"] = nbt[i]
            self._bn_flat, self._nbt = bn_flat, nbt

    def _workspace(self, ctx, batch, frames):
        need = ctx.res8_workspace_bytes(batch, frames, self.num_labels, True)
        if self._ws is None or self._ws.numel() < need or self._ws.device != ctx.device:
            self._ws = torch.empty(need, dtype=torch.uint8, device=ctx.device)
        return self._ws

    def _run_forward(self, feats, train: bool):
        ctx = get_context(feats.device, feats.shape[2])
        self._ensure_flat(feats.device)
        ws = self._workspace(ctx, feats.shape[0], feats.shape[1])
        return ctx.res8_fwd(feats, self._flat, self._bn_flat, self._nbt, train, ws)

    def _run_backward(self, feats, dlogits):
        # generic upstream gradient: the library's backward starts from CrossEntropy; for an arbitrary dlogits the
        # head backward is re-expressed through a label-free entry point
        ctx = get_context(feats.device, feats.shape[2])
        grads = torch.empty_like(self._flat)
        ctx.res8_bwd_from_dlogits(feats, dlogits, self._flat, grads, self._ws)
        return grads

    # ---- nn.Module API ----------------------------------------------------------------------------------
    def forward(self, x, lengths=None):
        if x.device.type != "cuda":
            raise RuntimeError("howl_b200.Res8 needs CUDA tensors (no CPU fallback)")
        ctx = get_context(x.device, x.shape[2])
        feats = ctx.to_time_major(x.contiguous().float())      # x[:, :1].permute(0,1,3,2).contiguous()  (cnn.py:128-129)
        if self.training and torch.is_grad_enabled():
            return _Res8Function.apply(self, feats, None, *self._param_list())
        return self._run_forward(feats, train=self.training)
