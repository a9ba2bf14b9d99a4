"""Res8TrainStep -- the fused train step of training/run/train.py:287-302 behind one object.

Owns the flat parameter / optimizer buffers and the activation workspace; every computation is a call into
libhowl_b200.so.  Data parallelism: each rank runs its shard, one NCCL all-reduce of the flat fp32
gradient (res8: 110,307 floats at L=12) sits between the backward and the fused AdamW; BatchNorm statistics stay
per rank (DDP semantics).
"""
from __future__ import annotations

import math
from typing import Optional, Tuple

import torch

from .runtime import Context


def res8_param_shapes(num_labels: int):
    """state_dict order of Res8's trainable tensors (howl/model/cnn.py:114-125)."""
    shapes = [("conv0.weight", (45, 1, 3, 3))]
    shapes += [(f"conv{i}.weight", (45, 45, 3, 3)) for i in range(1, 7)]
    shapes += [("output.weight", (num_labels, 45)), ("output.bias", (num_labels,))]
    return shapes


def init_res8_flat(num_labels: int, seed: int, device) -> torch.Tensor:
    """PyTorch-default initialisation (kaiming_uniform(a=sqrt 5) == U(+-1/sqrt(fan_in))) into the flat layout."""
    g = torch.Generator().manual_seed(seed)
    parts = []
    for _, shape in res8_param_shapes(num_labels):
        fan_in = math.prod(shape[1:]) if len(shape) > 1 else 45
        parts.append(((torch.rand(shape, generator=g) * 2 - 1) / math.sqrt(fan_in)).reshape(-1))
    return torch.cat(parts).to(device)


def mel_filterbank(n_mels: int, sample_rate: int = 16000, n_freqs: int = 257) -> torch.Tensor:
    """HTK mel triangles [n_freqs, n_mels] with the torch op sequence of torchaudio's melscale_fbanks (host side;
    the reference builds the same matrix in MelSpectrogram.__init__, transform.py:249-254)."""
    all_freqs = torch.linspace(0, sample_rate // 2, n_freqs)
    m_max = 2595.0 * math.log10(1.0 + (float(sample_rate // 2) / 700.0))
    m_pts = torch.linspace(0.0, m_max, n_mels + 2)
    f_pts = 700.0 * (10 ** (m_pts / 2595.0) - 1.0)
    f_diff = f_pts[1:] - f_pts[:-1]
    slopes = f_pts.unsqueeze(0) - all_freqs.unsqueeze(1)
    down = (-1.0 * slopes[:, :-2]) / f_diff[:-1]
    up = slopes[:, 2:] / f_diff[1:]
    return torch.clamp(torch.minimum(down, up), min=0.0)


class _HostPipeline:
    """Host-input front of a fused step (SURVEY §8 row a1 = ``ClassificationBatch.to``, howl/data/common/batch.py:27-32):
    ``step_host(*pinned_host_tensors)`` copies the step's inputs H2D on a copy stream into one of ``HOST_SLOTS`` device slots
    (so the copy of step k+1.. overlaps the compute of step k), runs ``self.step`` on the slot and reads the loss back (D2H)
    asynchronously; ``flush_host()`` drains and returns the last loss."""

    HOST_SLOTS = 3

    def _init_host_pipeline(self):
        self._copy_stream = torch.cuda.Stream(self.device)
        self._slots = None
        self._host_loss = torch.zeros(1).pin_memory()
        self._host_i = 0

    def step_host(self, *host_tensors: torch.Tensor) -> None:
        dev = self.device
        if self._slots is None or any(tuple(d.shape) != tuple(h.shape) or d.dtype != h.dtype for d, h in zip(self._slots[0], host_tensors)):
            self._slots = [tuple(torch.empty(h.shape, dtype=h.dtype, device=dev) for h in host_tensors) for _ in range(self.HOST_SLOTS)]
            self._slot_free = [torch.cuda.Event() for _ in range(self.HOST_SLOTS)]
            self._slot_ready = [torch.cuda.Event() for _ in range(self.HOST_SLOTS)]
            for e in self._slot_free:
                e.record(torch.cuda.current_stream(dev))
        k = self._host_i % self.HOST_SLOTS
        self._host_i += 1
        slot = self._slots[k]
        with torch.cuda.stream(self._copy_stream):
            self._copy_stream.wait_event(self._slot_free[k])
            for d, h in zip(slot, host_tensors):
                d.copy_(h, non_blocking=True)
            self._slot_ready[k].record(self._copy_stream)
        cur = torch.cuda.current_stream(dev)
        cur.wait_event(self._slot_ready[k])
        self.step(*slot)
        self._slot_free[k].record(cur)
        self._host_loss.copy_(self.loss, non_blocking=True)

    def flush_host(self) -> float:
        torch.cuda.current_stream(self.device).synchronize()
        return float(self._host_loss.item())

    def profile_groups(self, *inputs, reps: int = 3):
        """Device time of every kernel label over `reps` steps (CUDA events on the launching stream)."""
        acc, cnt = {}, {}
        for _ in range(reps):
            self.ctx.profile_begin()
            self.step(*inputs)
            for name, ms in self.ctx.profile_end():
                acc[name] = acc.get(name, 0.0) + ms
                cnt[name] = cnt.get(name, 0) + 1
        return [{"name": k, "ms": acc[k] / reps, "launches_per_step": cnt[k] // reps} for k in acc]


class Res8TrainStep(_HostPipeline):
    def __init__(self, device, num_labels: int, batch: int, samples: int, n_mels: int = 40, lr: float = 0.01,
                 weight_decay: float = 1e-5, zmuv: Tuple[float, float] = (0.0, 1.0), seed: int = 0, world_size: int = 1):
        self.ctx = Context(device, n_mels=n_mels)
        dev = self.ctx.device
        self.device, self.num_labels, self.batch, self.samples = dev, num_labels, batch, samples
        self.lr, self.weight_decay, self.zmuv, self.world = lr, weight_decay, zmuv, world_size
        self.params = init_res8_flat(num_labels, seed, dev)
        n = self.params.numel()
        assert n == self.ctx.res8_param_count(num_labels)
        self.grads, self.m, self.v = (torch.zeros(n, device=dev) for _ in range(3))
        self.bn_running = torch.stack([torch.zeros(6, 45), torch.ones(6, 45)], 1).contiguous().to(dev)  # [6][2][45]
        self.nbt = torch.zeros(6, dtype=torch.int64, device=dev)
        self.loss = torch.zeros(1, device=dev)
        self.logits = torch.zeros(batch, num_labels, device=dev)
        self.fb = mel_filterbank(n_mels).to(dev)
        self.frames = self.ctx.num_frames(samples)
        self.feat_bytes = (batch * self.frames * n_mels * 4 + 255) // 256 * 256
        self.ws = torch.empty(self.ctx.train_step_workspace_bytes(batch, samples, num_labels), dtype=torch.uint8, device=dev)
        self.step_count = 0
        self._init_host_pipeline()

    # ------------------------------------------------------------------ device-resident step
    def _ensure_capacity(self, b: int, t: int):
        """Workspace / logits follow the batch actually passed: a larger or longer batch than the one the step was built for grows
        them (the last, smaller batch of an epoch just uses a prefix)."""
        if t != self.samples or b > self.batch:
            self.batch, self.samples = max(b, self.batch), t
            self.frames = self.ctx.num_frames(t)
            self.ws = None
            self.ws = torch.empty(self.ctx.train_step_workspace_bytes(self.batch, t, self.num_labels), dtype=torch.uint8, device=self.device)
            self.logits = torch.zeros(self.batch, self.num_labels, device=self.device)

    def step(self, pcm: torch.Tensor, labels: torch.Tensor, rects: Optional[torch.Tensor] = None,
             fb: Optional[torch.Tensor] = None) -> torch.Tensor:
        """One train step on device tensors; returns the (device) loss tensor of this step.  `rects` [B,4] int32 (device): this
        step's SpecAugment rectangles; `fb` [257, M]: this step's mel filterbank (the VTLP matrix drawn for the step), default the
        standard bank."""
        if pcm.dim() != 2 or labels.shape[0] != pcm.shape[0]:
            raise ValueError("Res8TrainStep.step: pcm must be [B, T] and labels [B]")
        b, t = pcm.shape
        self._ensure_capacity(b, t)
        self.step_count += 1
        c = self.ctx
        fb = self.fb if fb is None else fb
        if self.world == 1:
            c.res8_train_step(pcm, labels, fb, self.zmuv, self.params, self.bn_running, self.nbt, self.grads, self.m,
                              self.v, self.step_count, self.lr, self.weight_decay, self.loss, self.logits, self.ws, rects)
        else:
            from .parallel import allreduce_flat_grads

            nfeat = b * self.frames * c.n_mels * 4
            feats = self.ws[:nfeat].view(torch.float32).view(b, self.frames, c.n_mels)
            ws = self.ws[(nfeat + 255) // 256 * 256:]
            c.frontend(pcm, fb, "time_major", zmuv=self.zmuv, rects=rects, out=feats)
            c.res8_fwd(feats, self.params, self.bn_running, self.nbt, True, ws, logits=self.logits)
            c.res8_bwd(feats, labels, self.params, self.grads, self.loss, ws, loss_scale_batch=b * self.world)
            allreduce_flat_grads(self.grads)     # one NCCL all-reduce(SUM) of the flat gradient over NVLink
            c.adamw(self.params, self.grads, self.m, self.v, self.step_count, self.lr, self.weight_decay)
        return self.loss

    # ------------------------------------------------------------------ state_dict interop
    def state_dict(self):
        out, off = {}, 0
        for name, shape in res8_param_shapes(self.num_labels):
            n = math.prod(shape)
            out[name] = self.params[off:off + n].view(shape).clone()
            off += n
        for i in range(6):
            out[f"bn{i + 1}.running_mean"] = self.bn_running[i, 0].clone()
            out[f"bn{i + 1}.running_var"] = self.bn_running[i, 1].clone()
            out[f"bn{i + 1}.num_batches_tracked"] = self.nbt[i].clone()
        return out


def lstm_param_shapes(num_labels: int, n_mels: int = 40):
    """state_dict order of SimpleLstm / SequentialLstm (howl/model/rnn.py:41-91)."""
    return [("lstm.weight_ih_l0", (512, n_mels)), ("lstm.weight_hh_l0", (512, 128)), ("lstm.bias_ih_l0", (512,)),
            ("lstm.bias_hh_l0", (512,)), ("dnn.0.weight", (256, 128)), ("dnn.0.bias", (256,)),
            ("dnn.2.weight", (num_labels, 256)), ("dnn.2.bias", (num_labels,))]


class LstmTrainStep(_HostPipeline):
    """Fused train step of the `lstm` model (frame objective): frontend -> LSTM -> MLP -> CE -> BPTT -> AdamW."""

    def __init__(self, device, num_labels: int, batch: int, samples: int, n_mels: int = 40, lr: float = 0.01,
                 weight_decay: float = 1e-5, zmuv: Tuple[float, float] = (0.0, 1.0), seed: int = 0, world_size: int = 1):
        self.ctx = Context(device, n_mels=n_mels)
        dev = self.ctx.device
        self.device, self.num_labels, self.batch, self.samples = dev, num_labels, batch, samples
        self.lr, self.weight_decay, self.zmuv, self.world = lr, weight_decay, zmuv, world_size
        g = torch.Generator().manual_seed(seed)
        parts = []
        for name, shape in lstm_param_shapes(num_labels, n_mels):
            fan = 128 if name.startswith("lstm.") else (shape[1] if len(shape) > 1 else {"dnn.0.bias": 128, "dnn.2.bias": 256}[name])
            parts.append(((torch.rand(shape, generator=g) * 2 - 1) / math.sqrt(fan)).reshape(-1))
        self.params = torch.cat(parts).to(dev)
        n = self.params.numel()
        assert n == self.ctx.lstm_param_count(num_labels)
        self.grads, self.m, self.v = (torch.zeros(n, device=dev) for _ in range(3))
        self.loss = torch.zeros(1, device=dev)
        self.logits = torch.zeros(batch, num_labels, device=dev)
        self.fb = mel_filterbank(n_mels).to(dev)
        self.frames = self.ctx.num_frames(samples)
        self.steps = (samples - self.ctx.n_fft) // self.ctx.hop + 1          # compute_lengths of a full clip
        self.lengths = torch.full((batch,), self.steps, dtype=torch.int64, device=dev)
        self.feat_bytes = (batch * self.frames * n_mels * 4 + 255) // 256 * 256
        self.ws = torch.empty(self.ctx.lstm_train_step_workspace_bytes(batch, samples, self.steps, num_labels), dtype=torch.uint8,
                              device=dev)
        self.step_count = 0
        self._init_host_pipeline()

    def step(self, pcm: torch.Tensor, labels: torch.Tensor, rects: Optional[torch.Tensor] = None, fb: Optional[torch.Tensor] = None) -> torch.Tensor:
        """rects [B,4] int32 (SpecAugment rectangles) / fb [257,M] (a VTLP-warped bank for this step) select the stage-by-stage path
        (the fused entry point takes the plain features); a smaller last batch uses a prefix of the buffers."""
        self.step_count += 1
        c = self.ctx
        b = pcm.shape[0]
        if b > self.batch:
            raise ValueError(f"batch of {b} exceeds the step's capacity {self.batch}")
        lengths, logits = self.lengths[:b], self.logits[:b]
        if self.world == 1 and rects is None and fb is None:
            c.lstm_train_step(pcm, labels, lengths, self.steps, self.fb, self.zmuv, self.params, self.grads, self.m, self.v,
                              self.step_count, self.lr, self.weight_decay, self.loss, logits, self.ws)
        else:
            feats = self.ws[: b * self.frames * c.n_mels * 4].view(torch.float32).view(b, self.frames, c.n_mels)
            ws = self.ws[self.feat_bytes:]
            c.frontend(pcm, self.fb if fb is None else fb, "time_major", zmuv=self.zmuv, rects=rects, out=feats)
            c.lstm_fwd(feats, lengths, self.steps, self.params, ws, train=True, out=logits)
            c.lstm_bwd(tuple(feats.shape), lengths, self.steps, labels, self.params, self.grads, self.loss, ws,
                       loss_scale_batch=b * self.world)
            if self.world > 1:
                from .parallel import allreduce_flat_grads

                allreduce_flat_grads(self.grads)
            c.adamw(self.params, self.grads, self.m, self.v, self.step_count, self.lr, self.weight_decay)
        return self.loss


class SeqLstmCtcTrainStep(LstmTrainStep):
    """Fused train step of the streaming `seq-lstm` with the CTC objective (training/run/train.py:294-302,
    envs/seq-lstm.env): frontend -> LSTM (state carried, detached) -> MLP per frame -> log_softmax + CTC -> BPTT -> AdamW."""

    def __init__(self, device, num_labels: int, batch: int, samples: int, blank: int, max_target_len: int = 3, **kw):
        super().__init__(device, num_labels, batch, samples, **kw)
        dev = self.device
        self.blank = blank
        self.state = torch.zeros(2, batch, 128, device=dev)
        self.scores = torch.zeros(self.steps, batch, num_labels, device=dev)
        self.ws = torch.empty(self.ctx.seq_lstm_train_step_workspace_bytes(batch, samples, self.steps, num_labels),
                              dtype=torch.uint8, device=dev)

    def step(self, pcm: torch.Tensor, targets: torch.Tensor, target_lengths: torch.Tensor) -> torch.Tensor:
        self.step_count += 1
        c = self.ctx
        if self.world == 1:
            c.seq_lstm_ctc_train_step(pcm, targets, target_lengths, self.blank, self.lengths, self.steps, self.fb, self.zmuv,
                                      self.params, self.state, self.grads, self.m, self.v, self.step_count, self.lr,
                                      self.weight_decay, self.loss, self.scores, self.ws)
        else:
            from .parallel import allreduce_flat_grads

            feats = self.ws[: self.batch * self.frames * c.n_mels * 4].view(torch.float32).view(self.batch, self.frames, c.n_mels)
            ws = self.ws[self.feat_bytes:]
            c.frontend(pcm, self.fb, "time_major", zmuv=self.zmuv, out=feats)
            c.lstm_fwd(feats, self.lengths, self.steps, self.params, ws, sequential=True, train=True, state_in=self.state,
                       state_out=self.state, out=self.scores)
            c.lstm_ctc_bwd(tuple(feats.shape), self.lengths, self.steps, targets, target_lengths, self.blank, self.params,
                           self.grads, self.loss, ws, loss_scale_batch=self.batch * self.world)
            allreduce_flat_grads(self.grads)
            c.adamw(self.params, self.grads, self.m, self.v, self.step_count, self.lr, self.weight_decay)
        return self.loss


class MobileNetTrainStep(_HostPipeline):
    """Fused train step of the `mobilenet` model (BASELINE.json configs[2]): frontend -> MobileNetV2 (bf16 tensor-core GEMMs, fp32
    master weights / BatchNorm statistics) -> CE -> backward -> [allreduce] -> AdamW."""

    def __init__(self, device, num_labels: int, batch: int, samples: int, n_mels: int = 40, lr: float = 0.01, weight_decay: float = 1e-5,
                 zmuv: Tuple[float, float] = (0.0, 1.0), seed: int = 0, world_size: int = 1, dropout_p: float = 0.2):
        from . import mobilenet as mb

        self.mb = mb
        self.ctx = Context(device, n_mels=n_mels)
        dev = self.ctx.device
        self.device, self.num_labels, self.batch, self.samples = dev, num_labels, batch, samples
        self.lr, self.weight_decay, self.zmuv, self.world, self.dropout_p, self.seed = lr, weight_decay, zmuv, world_size, dropout_p, seed
        self.params = mb.init_flat(num_labels, seed).to(dev)
        n = self.params.numel()
        assert n == int(self.ctx.lib.howl_b200_mobilenet_param_count(num_labels))
        self.grads, self.m, self.v = (torch.zeros(n, device=dev) for _ in range(3))
        nc = mb.bn_channels(self.ctx)
        self.bn_running = torch.cat([torch.zeros(1, nc), torch.ones(1, nc)]).contiguous().to(dev)       # [2][channels]: means, variances
        self.nbt = torch.zeros(mb.bn_layers(self.ctx), dtype=torch.int64, device=dev)
        self.loss = torch.zeros(1, device=dev)
        self.logits = torch.zeros(batch, num_labels, device=dev)
        self.fb = mel_filterbank(n_mels).to(dev)
        self.frames = self.ctx.num_frames(samples)
        self.feat_bytes = (batch * self.frames * n_mels * 4 + 255) // 256 * 256
        self.ws = torch.empty(self.feat_bytes + mb.workspace_bytes(self.ctx, batch, self.frames, num_labels), dtype=torch.uint8, device=dev)
        self.step_count = 0
        self._init_host_pipeline()

    def step(self, pcm: torch.Tensor, labels: torch.Tensor, rects: Optional[torch.Tensor] = None, fb: Optional[torch.Tensor] = None) -> torch.Tensor:
        """rects / fb as in LstmTrainStep.step."""
        self.step_count += 1
        c, mb = self.ctx, self.mb
        seed = self.seed * 1000003 + self.step_count
        b = pcm.shape[0]
        if b > self.batch:
            raise ValueError(f"batch of {b} exceeds the step's capacity {self.batch}")
        logits = self.logits[:b]
        if self.world == 1 and rects is None and fb is None:
            mb.train_step(c, pcm, labels, self.fb, self.zmuv, self.params, self.bn_running, self.nbt, self.grads, self.m, self.v,
                          self.step_count, self.lr, self.weight_decay, self.dropout_p, seed, self.loss, logits, self.ws)
        else:
            feats = self.ws[: b * self.frames * c.n_mels * 4].view(torch.float32).view(b, c.n_mels, self.frames)
            ws = self.ws[self.feat_bytes:]
            c.frontend(pcm, self.fb if fb is None else fb, "mels", zmuv=self.zmuv, rects=rects, out=feats)
            mb.forward(c, feats, self.params, self.bn_running, self.nbt, True, ws, self.dropout_p, seed, logits=logits)
            mb.backward(c, feats, labels, self.params, self.grads, self.loss, ws, self.dropout_p, seed, loss_scale_batch=b * self.world)
            if self.world > 1:
                from .parallel import allreduce_flat_grads

                allreduce_flat_grads(self.grads)
            c.adamw(self.params, self.grads, self.m, self.v, self.step_count, self.lr, self.weight_decay)
        return self.loss


class LasTrainStep(_HostPipeline):
    """Train step of the `las` model: frontend (stacked log-mel / delta / delta-delta, ZMUV) -> LASClassifier forward
    (batch statistics) -> CE -> backward -> [allreduce] -> AdamW, exact fp32, every stage a C-ABI call on one stream."""

    def __init__(self, device, num_labels: int, batch: int, samples: int, n_mels: int = 40, lr: float = 0.01, weight_decay: float = 1e-5,
                 zmuv: Tuple[float, float] = (0.0, 1.0), seed: int = 0, world_size: int = 1, dropout_p: float = 0.1):
        from . import las

        self.las = las
        self.ctx = Context(device, n_mels=n_mels)
        dev = self.ctx.device
        self.device, self.num_labels, self.batch, self.samples = dev, num_labels, batch, samples
        self.lr, self.weight_decay, self.zmuv, self.world, self.dropout_p, self.seed = lr, weight_decay, zmuv, world_size, dropout_p, seed
        g = torch.Generator().manual_seed(seed)
        init = []
        for name, shape in las.param_shapes(num_labels, n_mels):
            if ".conv_encoder." in name:
                init.append(torch.ones(shape) if name.endswith("weight") else torch.zeros(shape))
            elif name == "attn.context_vec":
                init.append(torch.rand(shape, generator=g) * 0.5 - 0.25)
            else:
                fan = 96 if "lstm_encoder" in name else {"encoder.conv1": 27, "encoder.conv2": 72, "attn.v_proj": 192, "attn.k_proj": 192,
                                                        "fc.0": 192, "fc.3": 256}[name.rsplit(".", 1)[0]]
                init.append((torch.rand(shape, generator=g) * 2 - 1) / math.sqrt(fan))
        self.params = torch.cat([t.reshape(-1) for t in init]).to(dev)
        n = self.params.numel()
        assert n == int(self.ctx.lib.howl_b200_las_param_count(num_labels, n_mels))
        self.grads, self.m, self.v = (torch.zeros(n, device=dev) for _ in range(3))
        self.bn_running = torch.stack([torch.stack([torch.zeros(8), torch.ones(8)])] * 2).contiguous().to(dev)
        self.nbt = torch.zeros(2, dtype=torch.int64, device=dev)
        self.loss = torch.zeros(1, device=dev)
        self.fb = mel_filterbank(n_mels).to(dev)
        self.frames = self.ctx.num_frames(samples)
        self.feats = torch.empty(batch, 3, n_mels, self.frames, device=dev)
        self.ws = torch.empty(las.workspace_bytes(self.ctx, batch, self.frames, n_mels, num_labels, True), dtype=torch.uint8, device=dev)
        # train.py:290-291 hands the model `audio_transform.compute_lengths(batch.lengths)` = floor((samples - 512) / 200) + 1 frames
        full = torch.full((batch,), (samples - self.ctx.n_fft) // self.ctx.hop + 1, dtype=torch.int64)
        self.full_lengths = las.encoder_lengths(self.ctx, full, batch, self.frames)
        self.logits = None
        self.step_count = 0
        self._init_host_pipeline()

    def step(self, pcm: torch.Tensor, labels: torch.Tensor, lengths=None, rects: Optional[torch.Tensor] = None,
             fb: Optional[torch.Tensor] = None) -> torch.Tensor:
        """pcm [B, T] f32 / int16, labels [B] i64 on the device; lengths [B] i64 = samples per clip (None: full clips); rects / fb as in
        LstmTrainStep.step.  The stacked layout applies the SpecAugment mask after the delta channels, as the reference does
        (spectrogram_augmentations run on the transform's output, train.py:289-290)."""
        self.step_count += 1
        c, las = self.ctx, self.las
        b = pcm.shape[0]
        if b > self.batch:
            raise ValueError(f"batch of {b} exceeds the step's capacity {self.batch}")
        seed = self.seed * 1000003 + self.step_count
        feats = self.feats[:b]
        c.frontend(pcm, self.fb if fb is None else fb, "stacked", zmuv=self.zmuv, out=feats)
        if rects is not None:
            c.spec_mask(feats, rects)
        if lengths is None:
            enc = self.full_lengths[:b]
        else:
            frames = torch.div(lengths.cpu().long() - c.n_fft, c.hop, rounding_mode="floor") + 1
            enc = las.encoder_lengths(c, frames, b, self.frames)
        self.logits = las.forward(c, feats, None, self.params, self.bn_running, self.nbt, True, self.ws, self.num_labels, self.dropout_p, seed, enc)
        las.backward(c, feats, enc, self.params, self.grads, self.ws, self.num_labels, labels=labels, loss=self.loss, dropout_p=self.dropout_p,
                     loss_scale_batch=b * self.world)
        if self.world > 1:
            from .parallel import allreduce_flat_grads

            allreduce_flat_grads(self.grads)
        c.adamw(self.params, self.grads, self.m, self.v, self.step_count, self.lr, self.weight_decay)
        return self.loss


class Trainer:
    """``howl.trainer.Trainer`` (``howl/trainer.py:11-31``): constructible from a ``TrainingConfig`` exactly like the reference's
    (which is a stub without a training loop); ``train`` is the addition the reference leaves open -- the epoch loop of
    ``training/run/train.py:280-307`` over the fused CUDA step.

    Construction touches neither CUDA nor the datasets (the reference's does not either); the fused step is built on the first
    ``train`` call, for the clip length and batch size of the first batch.
    """

    def __init__(self, training_cfg, logger=None, device="cuda:0"):
        import logging

        from .inference import SimpleContext

        self.training_cfg = training_cfg
        self.context_cfg = training_cfg.context_config
        if self.context_cfg.token_type != "word":
            raise NotImplementedError("Trainer: only word-level contexts (token_type='word') are supported")
        vocab = list(self.context_cfg.vocab or [])
        self.context = SimpleContext.for_vocab(vocab)           # labels: vocab..., then the negative label (context.py:86-97)
        sequence = self.context_cfg.sequence if self.context_cfg.sequence is not None else range(len(vocab))
        self.wake_word = " ".join(vocab[i] for i in sequence)        # Vocab.wakeword (howl/data/common/vocab.py:96-98): no stripping
        self.logger = logger or logging.getLogger(self.__class__.__name__)
        self.device = device
        self.step_obj = None
        self._aug = None
        self.epoch = 0
        if training_cfg.model_config.architecture not in self.STEPS:
            raise NotImplementedError(f"Trainer: a CUDA training step exists for {sorted(self.STEPS)}, not "
                                      f"{training_cfg.model_config.architecture!r}")

    # frame-objective steps by registry name (the CTC objective of seq-lstm needs targets: SeqLstmCtcTrainStep, driven directly)
    STEPS = {"res8": "Res8TrainStep", "mobilenet": "MobileNetTrainStep", "lstm": "LstmTrainStep", "las": "LasTrainStep"}

    def _ensure_step(self, pcm: torch.Tensor, zmuv):
        """The step object is built for the first batch; res8 grows its buffers with the batch, the other models are sized for the
        largest batch seen at construction (pass the largest batch first, or construct the step yourself)."""
        if self.step_obj is None or (not isinstance(self.step_obj, Res8TrainStep) and pcm.shape[0] > self.step_obj.batch):
            cfg = self.training_cfg
            cls = globals()[self.STEPS[cfg.model_config.architecture]]
            old = self.step_obj
            self.step_obj = cls(self.device, num_labels=self.context.num_labels, batch=pcm.shape[0], samples=pcm.shape[1],
                                lr=cfg.learning_rate if old is None else old.lr, weight_decay=cfg.weight_decay, zmuv=zmuv,
                                seed=self.context_cfg.seed)
            if old is not None:       # a larger batch arrived: keep the trained state
                for name in ("params", "m", "v", "bn_running", "nbt"):
                    if hasattr(old, name):
                        getattr(self.step_obj, name).copy_(getattr(old, name))
                self.step_obj.step_count = old.step_count
        return self.step_obj

    def train_epoch(self, batches, zmuv=(0.0, 1.0), augment: bool = True) -> float:
        """One pass over ``batches`` (iterable of (pcm [B,T] float32, labels [B] int64), host or device tensors); returns the mean
        loss.  As in the reference loop (``training/run/train.py:280-307``) every step draws, from the global ``random`` and in the
        reference's order, (1) the ``audio_transform.train()`` coin: with p = 0.75 a VTLP-warped filterbank (alpha ~ U[0.9, 1.1)) for
        the whole batch (``transform.py:93,441``), (2) the two ``SpecAugmentTransform().train()`` coins and, when they fall, one
        frequency / time rectangle per clip (``transform.py:310-326``); both are applied inside the fused frontend kernel.
        ``augment=False`` trains on the plain features (the parity-comparable recipe).  The learning rate decays by ``lr_decay``
        after the epoch (``train.py:306-307``)."""
        from .transform import SpecAugmentTransform, StandardAudioTransform, vtlp_filterbank

        if augment and self._aug is None:
            self._aug = (StandardAudioTransform().train(), SpecAugmentTransform().train())
        total, n = 0.0, 0
        for pcm, labels in batches:
            step = self._ensure_step(pcm, zmuv)
            fb = rects = None
            if augment:
                std, spec = self._aug
                if std.rand.random() < std.augment_params[0].prob:          # AugmentModule.forward's coin (transform.py:93)
                    import random as _random

                    alpha = _random.random() * 0.2 + 0.9                      # VtlpMelScale.forward (transform.py:441)
                    fb = vtlp_filterbank(alpha, std.num_mels, std.sample_rate, std.num_fft // 2 + 1).to(step.device)
                rects = spec.draw_rects(pcm.shape[0], std.num_mels, 1 + pcm.shape[1] // std.hop_length).to(step.device)
            loss = step.step(pcm.to(step.device, torch.float32), labels.to(step.device, torch.int64), rects=rects, fb=fb)
            total, n = total + float(loss.item()), n + 1
        if self.step_obj is not None:
            self.step_obj.lr *= self.training_cfg.lr_decay
        self.epoch += 1
        return total / max(n, 1)

    def train(self, make_batches, zmuv=(0.0, 1.0), augment: bool = True):
        """``num_epochs`` epochs; ``make_batches(epoch)`` returns that epoch's batch iterable.  Returns the per-epoch mean losses."""
        losses = []
        for epoch in range(self.training_cfg.num_epochs):
            losses.append(self.train_epoch(make_batches(epoch), zmuv, augment))
            self.logger.info("epoch %d: mean loss %.5f, lr %.6f", epoch, losses[-1], self.step_obj.lr if self.step_obj else float("nan"))
        return losses
