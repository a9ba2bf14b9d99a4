"""`mobilenet` of the howl registry (howl/model/cnn.py:15-29: a 1->3 channel stem, torchvision's MobileNetV2 and a Linear head) on top
of libhowl_b200.so: bf16 activations / gradients on the tensor cores, fp32 master weights and BatchNorm statistics.

The module keeps the reference's state_dict keys (so `howl-models/.../mobilenet/0/model-best.pt.bin` loads); all trainable tensors
are views into ONE flat fp32 buffer in state_dict order and all BatchNorm running statistics views into one [2][17059] buffer -- the
layouts of include/howl_b200.h.
"""
from __future__ import annotations

import math
from typing import List, Optional, Tuple

import torch
import torch.nn as nn

from .registry import RegisteredModel, _check_unchanged, _flatten_into
from .runtime import Context, HowlB200Error, _check, _ptr
from .transform import get_context

SETTING = ((1, 16, 1, 1), (6, 24, 2, 2), (6, 32, 3, 2), (6, 64, 4, 2), (6, 96, 3, 1), (6, 160, 3, 2), (6, 320, 1, 1))
LAST = 1280


def layer_plan(num_labels: int):
    """[(kind, prefix of the conv, prefix of its BatchNorm, weight shape, has_bias)] in state_dict order, then the classifier."""
    plan = [("stem", "downsample.0", "downsample.1", (3, 1, 3, 3), True), ("entry", "model.features.0.0", "model.features.0.1", (32, 3, 3, 3), False)]
    inp, idx = 32, 1
    for t, c, n, s in SETTING:
        for i in range(n):
            p, hidden, j = f"model.features.{idx}.conv", inp * t, 0
            if t != 1:
                plan.append(("pw", f"{p}.0.0", f"{p}.0.1", (hidden, inp, 1, 1), False))
                j = 1
            plan.append(("dw", f"{p}.{j}.0", f"{p}.{j}.1", (hidden, 1, 3, 3), False))
            plan.append(("pw", f"{p}.{j + 1}", f"{p}.{j + 2}", (c, hidden, 1, 1), False))
            inp, idx = c, idx + 1
    plan.append(("pw", "model.features.18.0", "model.features.18.1", (LAST, inp, 1, 1), False))
    return plan


def param_shapes(num_labels: int) -> List[Tuple[str, Tuple[int, ...]]]:
    out = []
    for _, conv, bn, shape, bias in layer_plan(num_labels):
        out.append((conv + ".weight", shape))
        if bias:
            out.append((conv + ".bias", (shape[0],)))
        out += [(bn + ".weight", (shape[0],)), (bn + ".bias", (shape[0],))]
    out += [("model.classifier.1.weight", (num_labels, LAST)), ("model.classifier.1.bias", (num_labels,))]
    return out


def init_flat(num_labels: int, seed: int = 0) -> torch.Tensor:
    """torchvision's MobileNetV2 initialisation (kaiming_normal fan_out for convolutions, BatchNorm 1 / 0, Linear N(0, 0.01) / 0) for
    the backbone and PyTorch's defaults for the stem, into the flat layout."""
    g = torch.Generator().manual_seed(seed)
    parts = []
    for name, shape in param_shapes(num_labels):
        if name.startswith("downsample.0"):
            bound = 1.0 / 3.0                                      # fan_in 9
            parts.append((torch.rand(shape, generator=g) * 2 - 1) * bound)
        elif name.endswith(".weight") and len(shape) == 4:
            fan_out = shape[0] * shape[2] * shape[3]
            parts.append(torch.randn(shape, generator=g) * math.sqrt(2.0 / fan_out))
        elif name == "model.classifier.1.weight":
            parts.append(torch.randn(shape, generator=g) * 0.01)
        elif name.endswith(".weight"):
            parts.append(torch.ones(shape))
        else:
            parts.append(torch.zeros(shape))
    return torch.cat([p.reshape(-1) for p in parts])


def algorithmic_flops(samples: int, num_labels: int, batch: int, n_mels: int = 40, hop: int = 200):
    """Algorithmic flops / HBM bytes per utterance of one train step (forward + data gradient + weight gradient = 3 x the forward MACs)."""
    frames = 1 + samples // hop
    h, w = n_mels, (frames + 4) // 2
    mac = 3 * 9 * n_mels * (frames + 4)
    h, w = (h - 1) // 2 + 1, (w - 1) // 2 + 1
    mac += 32 * 27 * h * w
    inp, gemm = 32, 32 * 27 * h * w
    for t, c, n, s in SETTING:
        for i in range(n):
            hidden, stride = inp * t, (s if i == 0 else 1)
            if t != 1:
                mac += inp * hidden * h * w
                gemm += inp * hidden * h * w
            if stride == 2:
                h, w = (h - 1) // 2 + 1, (w - 1) // 2 + 1
            mac += hidden * 9 * h * w + hidden * c * h * w
            gemm += hidden * c * h * w
            inp = c
    mac += inp * LAST * h * w + LAST * num_labels
    gemm += inp * LAST * h * w
    nparam = sum(math.prod(s) for _, s in param_shapes(num_labels))
    return {"flop": 3 * 2.0 * mac, "gemm_flop": 3 * 2.0 * gemm, "bytes": samples * 4 + 8 + 4 * num_labels + 24.0 * nparam / batch}


# ---------------------------------------------------------------------------------------------------------------------
# Context-level wrappers
# ---------------------------------------------------------------------------------------------------------------------
def _labels_from_params(ctx: Context, params: torch.Tensor) -> int:
    base = int(ctx.lib.howl_b200_mobilenet_param_count(1)) - (LAST + 1)
    n = params.numel() - base
    if n <= 0 or n % (LAST + 1):
        raise HowlB200Error(f"mobilenet: flat parameter buffer of {params.numel()} floats is not a mobilenet layout")
    return n // (LAST + 1)


def bn_channels(ctx: Context) -> int:
    return int(ctx.lib.howl_b200_mobilenet_bn_channels())


def bn_layers(ctx: Context) -> int:
    return int(ctx.lib.howl_b200_mobilenet_bn_layers())


def workspace_bytes(ctx: Context, batch: int, frames: int, num_labels: int) -> int:
    n = int(ctx.lib.howl_b200_mobilenet_workspace_bytes(batch, frames, ctx.n_mels, num_labels))
    if n < 0:
        raise HowlB200Error(f"mobilenet: unsupported shape B={batch} frames={frames} L={num_labels}")
    return n


def forward(ctx: Context, feats, params, bn_running, nbt, train: bool, ws, dropout_p: float = 0.0, seed: int = 0, logits=None):
    """feats [B, n_mels, frames] f32 ('mels' layout of the frontend)."""
    _check(feats, torch.float32, ctx.device, "feats")
    _check(params, torch.float32, ctx.device, "params")
    _check(bn_running, torch.float32, ctx.device, "bn_running")
    b, m, f = feats.shape
    L = _labels_from_params(ctx, params)
    if logits is None:
        logits = torch.empty(b, L, dtype=torch.float32, device=ctx.device)
    ctx._rc(ctx.lib.howl_b200_mobilenet_fwd(ctx.handle, ctx._stream(), _ptr(feats), b, f, m, L, _ptr(params), _ptr(bn_running), _ptr(nbt),
                                            int(train), float(dropout_p), int(seed), _ptr(logits), _ptr(ws), ws.numel()), "mobilenet_fwd")
    return logits


def backward(ctx: Context, feats, labels, params, grads, loss, ws, dropout_p: float = 0.0, seed: int = 0, loss_scale_batch=None, dlogits=None):
    b, m, f = feats.shape
    L = _labels_from_params(ctx, params)
    if dlogits is not None:
        _check(dlogits, torch.float32, ctx.device, "dlogits")
        ctx._rc(ctx.lib.howl_b200_mobilenet_bwd_dlogits(ctx.handle, ctx._stream(), _ptr(feats), _ptr(dlogits), b, f, m, L, _ptr(params),
                                                        _ptr(grads), float(dropout_p), int(seed), _ptr(ws), ws.numel()), "mobilenet_bwd_dlogits")
    else:
        _check(labels, torch.int64, ctx.device, "labels")
        ctx._rc(ctx.lib.howl_b200_mobilenet_bwd(ctx.handle, ctx._stream(), _ptr(feats), _ptr(labels), b, f, m, L, loss_scale_batch or b,
                                                _ptr(params), _ptr(grads), float(dropout_p), int(seed), _ptr(loss), _ptr(ws), ws.numel()),
                "mobilenet_bwd")


def train_step(ctx: Context, pcm, labels, fb, zmuv, params, bn_running, nbt, grads, m, v, step, lr, weight_decay, dropout_p, seed, loss,
               logits, ws):
    ctx._note_pcm(pcm)
    _check(labels, torch.int64, ctx.device, "labels")
    for name, t_ in (("fb", fb), ("params", params), ("bn_running", bn_running), ("grads", grads), ("m", m), ("v", v), ("loss", loss),
                     ("logits", logits)):
        _check(t_, torch.float32, ctx.device, name)
    b, t = pcm.shape
    L = _labels_from_params(ctx, params)
    ctx._note_fb(fb)
    ctx._rc(ctx.lib.howl_b200_mobilenet_train_step(ctx.handle, ctx._stream(), _ptr(pcm), _ptr(labels), b, t, _ptr(fb), float(zmuv[0]),
                                                   float(zmuv[1]), L, _ptr(params), _ptr(bn_running), _ptr(nbt), _ptr(grads), _ptr(m), _ptr(v),
                                                   step, lr, weight_decay, float(dropout_p), int(seed), _ptr(loss), _ptr(logits), _ptr(ws),
                                                   ws.numel()), "mobilenet_train_step")


# ---------------------------------------------------------------------------------------------------------------------
# nn.Module with the reference's state_dict
# ---------------------------------------------------------------------------------------------------------------------
class _Holder(nn.Module):
    """Container whose children are set by dotted path; numeric names are fine for nn.Module attributes."""


def _descend(root: nn.Module, path: str) -> nn.Module:
    node = root
    for part in path.split("."):
        if part not in node._modules:
            node.add_module(part, _Holder())
        node = node._modules[part]
    return node


class _MobileNetFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, model, feats, seed, *params):
        c = get_context(feats.device, feats.shape[1])
        ws = torch.empty(workspace_bytes(c, feats.shape[0], feats.shape[2], model.num_labels), dtype=torch.uint8, device=feats.device)
        ctx.model, ctx.feats, ctx.ws, ctx.seed, ctx.p = model, feats, ws, seed, model.dropout_p
        ctx.flat_ptr = model._flat.data_ptr()
        ctx.save_for_backward(*params)
        return forward(c, feats, model._flat, model._bn_flat, model._nbt, True, ws, model.dropout_p, seed)

    @staticmethod
    def backward(ctx, dlogits):
        model = ctx.model
        _check_unchanged(ctx, model._flat)
        c = get_context(ctx.feats.device, ctx.feats.shape[1])
        grads = torch.empty_like(model._flat)
        backward(c, ctx.feats, None, model._flat, grads, None, ctx.ws, ctx.p, ctx.seed, dlogits=dlogits.contiguous())
        ctx.ws = None
        out, off = [], 0
        for p in model._param_list():
            out.append(grads[off:off + p.numel()].view(p.shape))
            off += p.numel()
        return (None, None, None, *out)


class MobileNetClassifier(RegisteredModel, name="mobilenet"):
    """`pretrained` is accepted for signature compatibility: the reference downloads ImageNet weights at construction (cnn.py:22), which
    needs a network; here the backbone starts from torchvision's initialiser and `load_state_dict` brings trained weights."""

    def __init__(self, num_labels: int, config=None, pretrained: bool = False):
        super().__init__(num_labels)
        self.dropout_p = 0.2
        flat = init_flat(num_labels, seed=torch.initial_seed() % (2 ** 31))
        off = 0
        self._names = []
        for name, shape in param_shapes(num_labels):
            n = math.prod(shape)
            mod_path, leaf = name.rsplit(".", 1)
            _descend(self, mod_path).register_parameter(leaf, nn.Parameter(flat[off:off + n].view(shape).clone()))
            self._names.append(name)
            off += n
        self._bn_names = []
        for _, _, bn, shape, _ in layer_plan(num_labels):
            node = _descend(self, bn)
            node.register_buffer("running_mean", torch.zeros(shape[0]))
            node.register_buffer("running_var", torch.ones(shape[0]))
            node.register_buffer("num_batches_tracked", torch.tensor(0, dtype=torch.long))
            self._bn_names.append(bn)
        self._flat = self._bn_flat = self._nbt = self._ws = None
        self._step = 0

    def _param_list(self):
        return [_descend(self, n.rsplit(".", 1)[0])._parameters[n.rsplit(".", 1)[1]] for n in self._names]

    def _ensure_flat(self, device):
        _flatten_into(self, self._param_list(), "_flat", device)
        nodes = [_descend(self, n) for n in self._bn_names]
        total = sum(nd.running_mean.numel() for nd in nodes)
        ok = self._bn_flat is not None and self._bn_flat.device == device
        if ok:
            off = 0
            for i, nd in enumerate(nodes):
                c = nd.running_mean.numel()
                if (nd.running_mean.data_ptr() != self._bn_flat[0, off:off + c].data_ptr() or nd.running_var.data_ptr() != self._bn_flat[1, off:off + c].data_ptr()
                        or nd.num_batches_tracked.data_ptr() != self._nbt[i].data_ptr()):
                    ok = False
                    break
                off += c
        if not ok:
            bn_flat = torch.empty(2, total, dtype=torch.float32, device=device)
            nbt = torch.empty(len(nodes), dtype=torch.int64, device=device)
            off = 0
            for i, nd in enumerate(nodes):
                c = nd.running_mean.numel()
                bn_flat[0, off:off + c].copy_(nd.running_mean)
                bn_flat[1, off:off + c].copy_(nd.running_var)
                nbt[i].copy_(nd.num_batches_tracked)
                nd._buffers["running_mean"] = bn_flat[0, off:off + c]
                nd._buffers["running_var"] = bn_flat[1, off:off + c]
                nd._buffers["num_batches_tracked"] = nbt[i]
                off += c
            self._bn_flat, self._nbt = bn_flat, nbt

    def forward(self, x, lengths=None):
        if x.device.type != "cuda":
            raise RuntimeError("howl_b200.MobileNetClassifier needs CUDA tensors (no CPU fallback)")
        feats = x[:, 0].contiguous().float() if x.dim() == 4 else x.contiguous().float()      # log-Mels only (cnn.py:27)
        self._ensure_flat(feats.device)
        if self.training and torch.is_grad_enabled():
            self._step += 1
            return _MobileNetFunction.apply(self, feats, self._step, *self._param_list())
        c = get_context(feats.device, feats.shape[1])
        need = workspace_bytes(c, feats.shape[0], feats.shape[2], self.num_labels)
        if self._ws is None or self._ws.numel() < need or self._ws.device != feats.device:
            self._ws = torch.empty(need, dtype=torch.uint8, device=feats.device)
        return forward(c, feats, self._flat, self._bn_flat, self._nbt, self.training, self._ws)
