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


from .registry import ClassRegistry, RegisteredModel, _check_unchanged, _flatten_into  # noqa: F401  (re-exported)


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
    """Everything backward needs travels on the autograd ctx: the activation workspace of THIS forward (a fresh allocation per
    training forward -- a later forward, an eval pass or a larger batch cannot overwrite it) and the identity / version of the
    flat parameter buffer (checked in backward, as torch does for saved tensors)."""

    @staticmethod
    def forward(ctx, model, feats, labels_hint, *params):
        c = get_context(feats.device, feats.shape[2])
        ws = torch.empty(c.res8_workspace_bytes(feats.shape[0], feats.shape[1], model.num_labels, True), dtype=torch.uint8,
                         device=feats.device)
        logits = model._run_forward(feats, train=True, ws=ws)
        ctx.model, ctx.feats, ctx.ws = model, feats, ws
        ctx.flat_ptr = model._flat.data_ptr()
        ctx.save_for_backward(*params)
        return logits

    @staticmethod
    def backward(ctx, dlogits):
        model = ctx.model
        _check_unchanged(ctx, model._flat)
        grads = model._run_backward(ctx.feats, dlogits.contiguous(), ctx.ws)
        ctx.ws = None
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

    def _run_forward(self, feats, train: bool, ws=None):
        ctx = get_context(feats.device, feats.shape[2])
        if ws is None:      # inference / no-grad passes share one scratch workspace; training forwards bring their own
            ws = self._workspace(ctx, feats.shape[0], feats.shape[1])
        return ctx.res8_fwd(feats, self._flat, self._bn_flat, self._nbt, train, ws)

    def _run_backward(self, feats, dlogits, ws):
        # generic upstream gradient: the library's backward starts from CrossEntropy; for an arbitrary dlogits the
        # head backward is re-expressed through a label-free entry point
        ctx = get_context(feats.device, feats.shape[2])
        grads = torch.empty_like(self._flat)
        ctx.res8_bwd_from_dlogits(feats, dlogits, self._flat, grads, ws)
        return grads

    # ---- nn.Module API ----------------------------------------------------------------------------------
    def forward(self, x, lengths=None):
        if x.device.type != "cuda":
            raise RuntimeError("howl_b200.Res8 needs CUDA tensors (no CPU fallback)")
        ctx = get_context(x.device, x.shape[2])
        feats = ctx.to_time_major(x.contiguous().float())      # x[:, :1].permute(0,1,3,2).contiguous()  (cnn.py:128-129)
        self._ensure_flat(feats.device)
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
        c = get_context(feats.device, feats.shape[2])
        ws = torch.empty(c.lstm_workspace_bytes(feats.shape[0], max_steps, model.num_labels, True, model.SEQUENTIAL), dtype=torch.uint8,
                         device=feats.device)          # per-forward workspace, kept on the autograd ctx (see _Res8Function)
        ctx.model, ctx.shape, ctx.lengths, ctx.max_steps, ctx.ws = model, tuple(feats.shape), lengths, max_steps, ws
        ctx.flat_ptr = model._flat.data_ptr()
        ctx.save_for_backward(*params)
        return model._run(feats, lengths, max_steps, train=True, ws=ws)

    @staticmethod
    def backward(ctx, dlogits):
        m = ctx.model
        _check_unchanged(ctx, m._flat)
        c = get_context(m._flat.device, ctx.shape[2])
        grads = torch.empty_like(m._flat)
        c.lstm_bwd(ctx.shape, ctx.lengths, ctx.max_steps, None, m._flat, grads, None, ctx.ws, dlogits=dlogits.contiguous(),
                   sequential=dlogits.dim() == 3)
        ctx.ws = None
        out, off = [], 0
        for p in m._param_list():
            out.append(grads[off:off + p.numel()].view(p.shape))
            off += p.numel()
        return (None, None, None, None, *out)


class _LstmBase(RegisteredModel):
    HIDDEN = 128
    SEQUENTIAL = False

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

    def _run(self, feats, lengths, max_steps, train, ws=None):
        ctx = get_context(feats.device, feats.shape[2])
        ws = self._workspace(ctx, feats.shape[0], max_steps, train, False) if ws is None else ws
        return ctx.lstm_fwd(feats, lengths, max_steps, self._flat, ws, sequential=False, train=train)

    def forward(self, x, lengths):
        ctx, feats, lengths, max_steps = self._prepare(x, lengths)
        if self.training and torch.is_grad_enabled():
            return _LstmFunction.apply(self, feats, lengths, max_steps, *self._param_list())
        return self._run(feats, lengths, max_steps, train=False)


class SequentialLstm(_LstmBase, name="seq-lstm"):
    """dnn(h_t) for every frame -> [T', B, L]; carries (h, c) between calls when streaming (rnn.py:60-71)."""

    SEQUENTIAL = True

    def __init__(self, num_labels: int, config=None):
        super().__init__(num_labels, config)
        self.hc = None

    @property
    def streaming_state(self) -> Any:
        return self.hc

    @streaming_state.setter
    def streaming_state(self, x: Any):
        self.hc = x

    def _run(self, feats, lengths, max_steps, train, ws=None):
        ctx = get_context(feats.device, feats.shape[2])
        b = feats.shape[0]
        ws = self._workspace(ctx, b, max_steps, train, True) if ws is None else ws
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


# =====================================================================================================================
# ConvertedStaticModel (howl/model/base.py:40-62): a static model applied over sliding windows of the time axis
# =====================================================================================================================
class ConvertedStaticModel(RegisteredModel, name="converted"):
    """Runs ``model`` on windows of ``frame_window_size`` frames every ``frame_stride_size`` frames and stacks the outputs
    ``[n_windows, B, L]``.  Window arithmetic follows the reference literally, INCLUDING its first window, which is the tail
    ``x[..., frame_window_size:]`` rather than the head (base.py:55) -- flagged, kept.  In eval mode all equally sized windows go
    through the model as ONE batch (n_windows x B utterances per launch instead of n_windows launches)."""

    def __init__(self, model: RegisteredModel, frame_window_size: int, frame_stride_size: int):
        super().__init__(model.num_labels)
        self.model, self.frame_window_size, self.frame_stride_size = model, frame_window_size, frame_stride_size

    def compute_length(self, length: int):
        return None if length is None else max(1, (length - self.frame_window_size) // self.frame_stride_size)

    def _window_list(self, x):
        w, s, total = self.frame_window_size, self.frame_stride_size, x.size(3)
        out, idx, cur = [x[:, :, :, w:]], s, None
        while True:
            cur = x[:, :, :, idx:idx + w]
            if cur.size(3) != w:
                break
            out.append(cur)
            idx += s
        return out

    def forward(self, x, lengths):
        wins = self._window_list(x)
        if self.training or len(wins) < 3:
            return torch.stack([self.model(w, lengths) for w in wins])
        head = self.model(wins[0], lengths)                          # the odd-sized first window
        b = x.size(0)
        rep = None if lengths is None else lengths.repeat(len(wins) - 1)
        rest = self.model(torch.cat([w.contiguous() for w in wins[1:]], 0), rep)
        return torch.cat([head.unsqueeze(0), rest.view(len(wins) - 1, b, -1)], 0)


def _not_accelerated(name: str, where: str):
    class _Stub(RegisteredModel, name=name):
        def __init__(self, *a, **k):
            raise NotImplementedError(f"howl_b200: model '{name}' ({where}) is not on the accelerated hot path (SURVEY.md §2/§8a, "
                                      "DESIGN.md §1); inside a reference checkout plugin.install() keeps the reference's torch module")
    _Stub.__name__ = f"NotAccelerated_{name.replace('-', '_')}"
    return _Stub


# registry names of the reference without a CUDA path here: constructible only through the reference itself
for _n, _w in (("small-cnn", "howl/model/cnn.py:40-74"), ("seq-cnn", "howl/model/cnn.py:77-104"), ("gru", "howl/model/rnn.py:94-130")):
    _not_accelerated(_n, _w)

# name -> class of every model with a CUDA forward/backward in this package (what plugin.install() re-points)
ACCELERATED = {"res8": Res8, "lstm": SimpleLstm, "seq-lstm": SequentialLstm}


from .mobilenet import MobileNetClassifier  # noqa: E402  (own module; shares registry.py with this one)

ACCELERATED["mobilenet"] = MobileNetClassifier

from .las import LASClassifier  # noqa: E402  (forward only)

ACCELERATED["las"] = LASClassifier
