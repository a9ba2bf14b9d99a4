"""`las` of the howl registry (howl/model/rnn.py:133-215: two small convolutions over the three stacked feature channels, a bidirectional
LSTM(352 -> 96), a fixed-context 4-head attention and an MLP) on libhowl_b200.so: forward (eval / batch statistics) and the autograd
backward, exact fp32 (csrc/las.cu).  Dropout(0.1) of the fc block uses the library's counter-based mask, not torch's generator."""
from __future__ import annotations

import ctypes as C
import math

import numpy as np
import torch
import torch.nn as nn

from .registry import RegisteredModel, _check_unchanged, _flatten_into
from .runtime import HowlB200Error, _check, _ptr
from .transform import get_context


def param_shapes(num_labels: int, n_mels: int = 40):
    """Trainable tensors in `parameters()` order (the encoder's convolutions appear a second time in the state_dict under
    `encoder.conv_encoder.{0,4}`: the reference registers the same modules twice)."""
    inp = 8 * (n_mels + 4)
    out = [("encoder.conv1.weight", (8, 3, 3, 3)), ("encoder.conv1.bias", (8,)), ("encoder.conv2.weight", (8, 8, 3, 3)), ("encoder.conv2.bias", (8,)),
           ("encoder.conv_encoder.1.weight", (8,)), ("encoder.conv_encoder.1.bias", (8,)), ("encoder.conv_encoder.5.weight", (8,)),
           ("encoder.conv_encoder.5.bias", (8,))]
    for suffix in ("", "_reverse"):
        out += [(f"encoder.lstm_encoder.weight_ih_l0{suffix}", (384, inp)), (f"encoder.lstm_encoder.weight_hh_l0{suffix}", (384, 96)),
                (f"encoder.lstm_encoder.bias_ih_l0{suffix}", (384,)), (f"encoder.lstm_encoder.bias_hh_l0{suffix}", (384,))]
    out += [("attn.context_vec", (192,)), ("attn.v_proj.weight", (192, 192)), ("attn.v_proj.bias", (192,)), ("attn.k_proj.weight", (192, 192)),
            ("attn.k_proj.bias", (192,)), ("fc.0.weight", (256, 192)), ("fc.0.bias", (256,)), ("fc.3.weight", (num_labels, 256)), ("fc.3.bias", (num_labels,))]
    return out


def las_lengths(ctx, lengths: torch.Tensor) -> torch.Tensor:
    """LASEncoder.forward's length arithmetic (rnn.py:163-168) on the host, bit exact (float floors as the reference)."""
    src = np.ascontiguousarray(lengths.detach().cpu().numpy(), dtype=np.int64)
    out = np.zeros_like(src)
    ctx._rc(ctx.lib.howl_b200_las_lengths(src.ctypes.data_as(C.c_void_p), src.size, out.ctypes.data_as(C.c_void_p)), "las_lengths")
    return torch.from_numpy(out)


def workspace_bytes(ctx, batch: int, frames: int, n_mels: int, num_labels: int, train: bool) -> int:
    need = int(ctx.lib.howl_b200_las_workspace_bytes(batch, frames, n_mels, num_labels, int(train)))
    if need < 0:
        raise HowlB200Error(f"las: unsupported shape B={batch} frames={frames} mels={n_mels}")
    return need


def forward(ctx, feats, lengths, params, bn_running, nbt, train: bool, ws=None, num_labels=None, dropout_p: float = 0.0, seed: int = 0,
            enc_lengths=None):
    """feats [B, 3, n_mels, frames] f32 (the stacked frontend layout); lengths [B] i64 frames per clip (host or device) or None.
    train=True keeps the activations in `ws` (sized with train=True) for `backward`."""
    _check(feats, torch.float32, ctx.device, "feats")
    _check(params, torch.float32, ctx.device, "params")
    b, c, m, f = feats.shape
    if c != 3:
        raise HowlB200Error("las: needs the three stacked feature channels [B, 3, M, F]")
    if enc_lengths is None:
        enc_lengths = encoder_lengths(ctx, lengths, b, f)
    L = num_labels
    need = workspace_bytes(ctx, b, f, m, L, train)
    if ws is None:
        ws = torch.empty(need, dtype=torch.uint8, device=ctx.device)
    elif ws.numel() < need:
        raise HowlB200Error(f"las: workspace {ws.numel()} < required {need}")
    logits = torch.empty(b, L, dtype=torch.float32, device=ctx.device)
    ctx._rc(ctx.lib.howl_b200_las_fwd(ctx.handle, ctx._stream(), _ptr(feats), _ptr(enc_lengths), b, f, m, L, _ptr(params), _ptr(bn_running),
                                      _ptr(nbt), int(train), float(dropout_p), int(seed), _ptr(logits), _ptr(ws), ws.numel()), "las_fwd")
    return logits


def encoder_lengths(ctx, lengths, batch: int, frames: int) -> torch.Tensor:
    if lengths is None:
        lengths = torch.full((batch,), frames, dtype=torch.int64)
    return las_lengths(ctx, lengths).to(ctx.device)


def backward(ctx, feats, enc_lengths, params, grads, ws, num_labels: int, labels=None, dlogits=None, loss=None, dropout_p: float = 0.0,
             loss_scale_batch=None):
    """Gradients (flat layout, overwritten) of the train-mode forward kept in `ws`: CrossEntropyLoss(mean) over `labels`, or a caller's
    `dlogits` [B, L]."""
    b, _, m, f = feats.shape
    _check(grads, torch.float32, ctx.device, "grads")
    _check(enc_lengths, torch.int64, ctx.device, "enc_lengths")
    if dlogits is not None:
        _check(dlogits, torch.float32, ctx.device, "dlogits")
        ctx._rc(ctx.lib.howl_b200_las_bwd_dlogits(ctx.handle, ctx._stream(), _ptr(feats), _ptr(enc_lengths), _ptr(dlogits), b, f, m, num_labels,
                                                  _ptr(params), _ptr(grads), float(dropout_p), _ptr(ws), ws.numel()), "las_bwd_dlogits")
    else:
        _check(labels, torch.int64, ctx.device, "labels")
        ctx._rc(ctx.lib.howl_b200_las_bwd(ctx.handle, ctx._stream(), _ptr(feats), _ptr(enc_lengths), _ptr(labels), b, f, m, num_labels,
                                          loss_scale_batch or b, _ptr(params), _ptr(grads), float(dropout_p), _ptr(loss) if loss is not None else None,
                                          _ptr(ws), ws.numel()), "las_bwd")


class _Holder(nn.Module):
    pass


def _descend(root, path):
    node = root
    for part in path.split("."):
        if part not in node._modules:
            node.add_module(part, _Holder())
        node = node._modules[part]
    return node


class _LasFunction(torch.autograd.Function):
    """Train-mode forward + backward through the C ABI; each forward owns its workspace, so several forwards may precede a backward."""

    @staticmethod
    def forward(ctx, model, feats, enc_lengths, seed, *params):
        c = get_context(feats.device, feats.shape[2])
        b, _, m, f = feats.shape
        ws = torch.empty(workspace_bytes(c, b, f, m, model.num_labels, True), dtype=torch.uint8, device=feats.device)
        ctx.model, ctx.feats, ctx.enc, ctx.ws, ctx.p = model, feats, enc_lengths, ws, model.dropout_p
        ctx.flat_ptr = model._flat.data_ptr()
        ctx.save_for_backward(*params)
        return forward(c, feats, None, model._flat, model._bn_flat, model._nbt, True, ws, model.num_labels, model.dropout_p, seed, enc_lengths)

    @staticmethod
    def backward(ctx, dlogits):
        model = ctx.model
        _check_unchanged(ctx, model._flat)
        c = get_context(ctx.feats.device, ctx.feats.shape[2])
        grads = torch.empty_like(model._flat)
        backward(c, ctx.feats, ctx.enc, model._flat, grads, ctx.ws, model.num_labels, dlogits=dlogits.contiguous().float(), dropout_p=ctx.p)
        ctx.ws = None
        out, off = [], 0
        for p in model._param_list():
            out.append(grads[off:off + p.numel()].view(p.shape))
            off += p.numel()
        return (None, None, None, None, *out)


class LASClassifier(RegisteredModel, name="las"):
    def __init__(self, num_labels: int, config=None):
        super().__init__(num_labels)
        self.dropout_p = float(getattr(config, "dropout", 0.1)) if config is not None else 0.1       # LASClassifierConfig.dropout (rnn.py:30)
        self._step = 0
        g = torch.Generator().manual_seed(torch.initial_seed() % (2 ** 31))
        self._names = []
        for path in ("encoder.conv1", "encoder.conv2", "encoder.conv_encoder.0", "encoder.conv_encoder.1", "encoder.conv_encoder.4",
                     "encoder.conv_encoder.5", "encoder.lstm_encoder"):      # module order of the reference (state_dict key order)
            _descend(self, path)
        for name, shape in param_shapes(num_labels):
            mod, leaf = name.rsplit(".", 1)
            if name == "attn.context_vec":
                init = torch.rand(shape, generator=g) * 0.5 - 0.25
            elif ".conv_encoder." in name:
                init = torch.ones(shape) if leaf == "weight" else torch.zeros(shape)
            else:
                fan = 96 if "lstm_encoder" in name else (math.prod(shape[1:]) if len(shape) > 1 else {"encoder.conv1.bias": 27, "encoder.conv2.bias": 72,
                                                                                                    "attn.v_proj.bias": 192, "attn.k_proj.bias": 192,
                                                                                                    "fc.0.bias": 192, "fc.3.bias": 256}[name])
                init = (torch.rand(shape, generator=g) * 2 - 1) / math.sqrt(fan)
            _descend(self, mod).register_parameter(leaf, nn.Parameter(init))
            self._names.append(name)
        # the reference registers conv1 / conv2 a second time inside conv_encoder: same Parameter objects under the alias keys
        enc = self._modules["encoder"]
        for alias, src in (("0", "conv1"), ("4", "conv2")):
            node = _descend(enc, f"conv_encoder.{alias}")
            node._parameters["weight"], node._parameters["bias"] = enc._modules[src]._parameters["weight"], enc._modules[src]._parameters["bias"]
        for idx in ("1", "5"):
            node = _descend(enc, f"conv_encoder.{idx}")
            node.register_buffer("running_mean", torch.zeros(8))
            node.register_buffer("running_var", torch.ones(8))
            node.register_buffer("num_batches_tracked", torch.tensor(0, dtype=torch.long))
        self._flat = self._bn_flat = self._nbt = self._ws = None

    def _param_list(self):
        return [_descend(self, n.rsplit(".", 1)[0])._parameters[n.rsplit(".", 1)[1]] for n in self._names]

    def _ensure_flat(self, device):
        _flatten_into(self, self._param_list(), "_flat", device)
        enc = self._modules["encoder"]._modules["conv_encoder"]
        nodes = [enc._modules["1"], enc._modules["5"]]
        ok = self._bn_flat is not None and self._bn_flat.device == device and all(
            nd.running_mean.data_ptr() == self._bn_flat[i, 0].data_ptr() and nd.running_var.data_ptr() == self._bn_flat[i, 1].data_ptr() and
            nd.num_batches_tracked.data_ptr() == self._nbt[i].data_ptr() for i, nd in enumerate(nodes))
        if not ok:
            bn = torch.empty(2, 2, 8, dtype=torch.float32, device=device)
            nbt = torch.empty(2, dtype=torch.int64, device=device)
            for i, nd in enumerate(nodes):
                bn[i, 0].copy_(nd.running_mean)
                bn[i, 1].copy_(nd.running_var)
                nbt[i].copy_(nd.num_batches_tracked)
                nd._buffers["running_mean"], nd._buffers["running_var"], nd._buffers["num_batches_tracked"] = bn[i, 0], bn[i, 1], nbt[i]
            self._bn_flat, self._nbt = bn, nbt

    def forward(self, x, lengths=None):
        if x.device.type != "cuda":
            raise RuntimeError("howl_b200.LASClassifier needs CUDA tensors (no CPU fallback)")
        ctx = get_context(x.device, x.shape[2])
        self._ensure_flat(x.device)
        feats = x.contiguous().float()
        if self.training and torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            self._step += 1
            enc = encoder_lengths(ctx, lengths, feats.shape[0], feats.shape[3])
            return _LasFunction.apply(self, feats, enc, self._step, *self._param_list())
        return forward(ctx, feats, lengths, self._flat, self._bn_flat, self._nbt, self.training, num_labels=self.num_labels)
