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
        _flatten_into(self, self._param_list(), "_flat", device)
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


# =====================================================================================================================
# lstm / seq-lstm (howl/model/rnn.py:41-91)
# =====================================================================================================================
class _LstmCell(nn.Module):
    """Parameter holder with torch.nn.LSTM's names and default init (U(+-1/sqrt(hidden)))."""

    def __init__(self, n_mels: int, hidden: int):
        super().__init__()
        bound = 1.0 / math.sqrt(hidden)
        for name, shape in (("weight_ih_l0", (4 * hidden, n_mels)), ("weight_hh_l0", (4 * hidden, hidden)),
                            ("bias_ih_l0", (4 * hidden,)), ("bias_hh_l0", (4 * hidden,))):
            setattr(self, name, nn.Parameter((torch.rand(shape) * 2 - 1) * bound))


class _LstmFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, model, feats, lengths, max_steps, *params):
        ctx.model, ctx.shape, ctx.lengths, ctx.max_steps = model, tuple(feats.shape), lengths, max_steps
        return model._run(feats, lengths, max_steps, train=True)

    @staticmethod
    def backward(ctx, dlogits):
        m = ctx.model
        c = get_context(m._flat.device, ctx.shape[2])
        grads = torch.empty_like(m._flat)
        c.lstm_bwd(ctx.shape, ctx.lengths, ctx.max_steps, None, m._flat, grads, None, m._ws, dlogits=dlogits.contiguous(),
                   sequential=dlogits.dim() == 3)
        out, off = [], 0
        for p in m._param_list():
            out.append(grads[off:off + p.numel()].view(p.shape))
            off += p.numel()
        return (None, None, None, None, *out)


class _LstmBase(RegisteredModel):
    HIDDEN = 128

    def __init__(self, num_labels: int, config=None):
        super().__init__(num_labels)
        n_mels = getattr(config, "num_mels", 40) if config is not None else 40
        self.n_mels = n_mels
        self.lstm = _LstmCell(n_mels, self.HIDDEN)
        self.dnn = nn.Sequential(_Weight((2 * self.HIDDEN, self.HIDDEN), (2 * self.HIDDEN,)), nn.Identity(),
                                 _Weight((num_labels, 2 * self.HIDDEN), (num_labels,)))
        self._flat = self._ws = None

    def _param_list(self):
        return [self.lstm.weight_ih_l0, self.lstm.weight_hh_l0, self.lstm.bias_ih_l0, self.lstm.bias_hh_l0,
                self.dnn[0].weight, self.dnn[0].bias, self.dnn[2].weight, self.dnn[2].bias]

    def _prepare(self, x, lengths):
        if x.device.type != "cuda":
            raise RuntimeError("howl_b200 lstm models need CUDA tensors (no CPU fallback)")
        ctx = get_context(x.device, x.shape[2])
        feats = ctx.to_time_major(x.contiguous().float())      # x[:, 0].permute(2, 0, 1) of rnn.py:61-65,86-88, kept [B, F, M]
        if lengths is None:
            lengths = torch.full((x.shape[0],), feats.shape[1], dtype=torch.int64)
        max_steps = int(lengths.max())
        lengths = lengths.to(device=x.device, dtype=torch.int64).contiguous()
        _flatten_into(self, self._param_list(), "_flat", x.device)
        return ctx, feats, lengths, max_steps

    def _workspace(self, ctx, batch, max_steps, train, sequential):
        need = ctx.lstm_workspace_bytes(batch, max_steps, self.num_labels, train, sequential)
        if self._ws is None or self._ws.numel() < need or self._ws.device != ctx.device:
            self._ws = torch.empty(need, dtype=torch.uint8, device=ctx.device)
        return self._ws


class SimpleLstm(_LstmBase, name="lstm"):
    """dnn(h_n) -> [B, L]; the reference never carries state for this model (its streaming_state setter is a no-op)."""

    def _run(self, feats, lengths, max_steps, train):
        ctx = get_context(feats.device, feats.shape[2])
        ws = self._workspace(ctx, feats.shape[0], max_steps, train, False)
        return ctx.lstm_fwd(feats, lengths, max_steps, self._flat, ws, sequential=False, train=train)

    def forward(self, x, lengths):
        ctx, feats, lengths, max_steps = self._prepare(x, lengths)
        if self.training and torch.is_grad_enabled():
            return _LstmFunction.apply(self, feats, lengths, max_steps, *self._param_list())
        return self._run(feats, lengths, max_steps, train=False)


class SequentialLstm(_LstmBase, name="seq-lstm"):
    """dnn(h_t) for every frame -> [T', B, L]; carries (h, c) between calls when streaming (rnn.py:60-71)."""

    def __init__(self, num_labels: int, config=None):
        super().__init__(num_labels, config)
        self.hc = None

    @property
    def streaming_state(self) -> Any:
        return self.hc

    @streaming_state.setter
    def streaming_state(self, x: Any):
        self.hc = x

    def _run(self, feats, lengths, max_steps, train):
        ctx = get_context(feats.device, feats.shape[2])
        b = feats.shape[0]
        ws = self._workspace(ctx, b, max_steps, train, True)
        state_in = None
        if self.is_streaming and self.hc is not None:
            state_in = torch.stack([self.hc[0].reshape(b, self.HIDDEN), self.hc[1].reshape(b, self.HIDDEN)]).contiguous().float()
        state_out = torch.empty(2, b, self.HIDDEN, dtype=torch.float32, device=feats.device) if self.is_streaming else None
        out = ctx.lstm_fwd(feats, lengths, max_steps, self._flat, ws, sequential=True, train=train, state_in=state_in,
                           state_out=state_out)
        if self.is_streaming:      # detached by construction (rnn.py:66-68)
            self.hc = (state_out[0].unsqueeze(0).clone(), state_out[1].unsqueeze(0).clone())
        return out

    def forward(self, x, lengths):
        ctx, feats, lengths, max_steps = self._prepare(x, lengths)
        if self.training and torch.is_grad_enabled():
            return _LstmFunction.apply(self, feats, lengths, max_steps, *self._param_list())
        return self._run(feats, lengths, max_steps, train=False)
