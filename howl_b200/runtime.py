"""Thin torch-tensor wrapper over the C ABI: owns a howl_ctx_t and hands device pointers to the library.

PyTorch is plumbing here (device memory, streams); every computation is done by libhowl_b200.so.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _lib
from ._lib import FE_MELS_ONLY, FE_STACKED, FE_TIME_MAJOR, FE_ZMUV, HowlB200Error


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _check(t: torch.Tensor, dtype, device, name: str):
    if t.dtype != dtype or t.device != device or not t.is_contiguous():
        raise HowlB200Error(f"{name}: need contiguous {dtype} tensor on {device}, got {t.dtype} {t.device} "
                            f"contiguous={t.is_contiguous()}")


class Context:
    """One per (device, thread).  Mirrors howl_ctx_t."""

    def __init__(self, device="cuda:0", n_mels: int = 40, sample_rate: int = 16000, n_fft: int = 512, hop: int = 200):
        self.lib = _lib.load()
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise HowlB200Error("howl_b200 runs on CUDA devices only (no CPU fallback)")
        index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self.device = torch.device("cuda", index)
        self.n_mels, self.hop, self.n_fft, self.sample_rate = n_mels, hop, n_fft, sample_rate
        cfg = _lib.FrontendCfg(sample_rate, n_fft, hop, n_mels)
        handle = C.c_void_p()
        rc = self.lib.howl_b200_create(index, C.byref(cfg), C.byref(handle))
        if rc != 0:
            raise HowlB200Error(f"howl_b200_create failed ({rc}): {self.lib.howl_b200_last_error(None).decode()}")
        self.handle = handle
        self._ws = None
        self._fb_seen = None      # (tensor, version) of the filterbank of the previous frontend call

    def close(self):
        if getattr(self, "handle", None):
            self.lib.howl_b200_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ helpers
    def _rc(self, rc: int, what: str):
        if rc != 0:
            raise HowlB200Error(f"{what} failed ({rc}): {self.lib.howl_b200_last_error(self.handle).decode()}")

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    @property
    def launch_count(self) -> int:
        return int(self.lib.howl_b200_launch_count(self.handle))

    @property
    def sm_count(self) -> int:
        return int(self.lib.howl_b200_sm_count(self.handle))

    def set_option(self, name: str, value: int):
        self._rc(self.lib.howl_b200_set_option(self.handle, name.encode(), int(value)), "set_option")

    def selftest_umma(self, a: torch.Tensor, b: torch.Tensor, mn_major: bool, variant: int = 0) -> torch.Tensor:
        d = torch.empty(128, 48, dtype=torch.float32, device=self.device)
        self._rc(self.lib.howl_b200_selftest_umma(self.handle, self._stream(), _ptr(a), _ptr(b), _ptr(d), int(mn_major),
                                                  variant), "selftest_umma")
        return d

    def debug_umma_bench(self, mode: int, n: int, iters: int = 4096) -> float:
        """Tuning aid: cycles per M=128 x n x 16 bf16 tcgen05.mma (mode bit 0: A from TMEM, 1: B MN-major, 2: M=64)."""
        out = torch.zeros(1, dtype=torch.int64, device=self.device)
        self._rc(self.lib.howl_b200_debug_umma_bench(self.handle, self._stream(), mode, n, iters, _ptr(out)), "debug_umma_bench")
        return float(out.item()) / iters

    def debug_stream_profile(self, buf, kind: int):
        """Tuning aid: per-CTA wait-cycle counters of the tensor-core conv kernels (kind 1 fwd, 2 dgrad) -> buf[sm][16] int64."""
        self._rc(self.lib.howl_b200_debug_stream_profile(self.handle, _ptr(buf) if buf is not None else None, kind),
                 "debug_stream_profile")

    def profile_begin(self):
        self._rc(self.lib.howl_b200_profile_begin(self.handle, self._stream()), "profile_begin")

    def profile_end(self):
        """-> list of (kernel label, device ms) for every launch since profile_begin()."""
        cap = 1024
        names = C.create_string_buffer(cap * 32)
        ms = (C.c_float * cap)()
        n = self.lib.howl_b200_profile_end(self.handle, names, len(names), ms, cap)
        if n < 0:
            self._rc(n, "profile_end")
        labels = names.raw.split(b"\0")[:n]
        return [(labels[i].decode(), float(ms[i])) for i in range(n)]

    def _note_pcm(self, pcm: torch.Tensor, name: str = "pcm"):
        """PCM may be float32 or int16 (the wav files' own format: K1 converts with x / 32768); tells the library which."""
        if pcm.dtype not in (torch.float32, torch.int16) or pcm.device != self.device or not pcm.is_contiguous():
            raise HowlB200Error(f"{name}: need a contiguous float32 or int16 tensor on {self.device}, got {pcm.dtype} {pcm.device}")
        i16 = pcm.dtype == torch.int16
        if i16 != getattr(self, "_pcm_i16", False):
            self.set_option("pcm_i16", int(i16))
            self._pcm_i16 = i16

    def _note_fb(self, fb: torch.Tensor):
        """Tell the library when this call's filterbank is the very tensor (same object, unmodified) of the previous call, so that
        the compact bank / work plan on the device are reused.  The tensor is kept referenced: its storage cannot be recycled for
        another bank in between."""
        seen = self._fb_seen
        if seen is not None and seen[0] is fb and seen[1] == fb._version:
            self.lib.howl_b200_set_option(self.handle, b"fb_unchanged", 1)
        self._fb_seen = (fb, fb._version)

    def num_frames(self, samples: int) -> int:
        return int(self.lib.howl_b200_num_frames(samples, self.hop))

    def workspace(self, nbytes: int) -> torch.Tensor:
        if self._ws is None or self._ws.numel() < nbytes:
            self._ws = None
            self._ws = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
        return self._ws

    # ------------------------------------------------------------------ frontend
    def frontend(self, pcm: torch.Tensor, fb: torch.Tensor, layout: str = "stacked", zmuv=None,
                 rects: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """layout: 'stacked' [B,3,M,F] | 'mels' [B,M,F] | 'time_major' [B,F,M]; zmuv = (mean, std) or None."""
        self._note_pcm(pcm)
        _check(fb, torch.float32, self.device, "fb")
        if pcm.dim() != 2 or tuple(fb.shape) != (self.n_fft // 2 + 1, self.n_mels):
            raise HowlB200Error(f"frontend: pcm must be [B,T] and fb [{self.n_fft // 2 + 1},{self.n_mels}]")
        b, t = pcm.shape
        f = self.num_frames(t)
        flag = {"stacked": FE_STACKED, "mels": FE_MELS_ONLY, "time_major": FE_TIME_MAJOR}[layout]
        shape = {"stacked": (b, 3, self.n_mels, f), "mels": (b, self.n_mels, f), "time_major": (b, f, self.n_mels)}[layout]
        if out is None:
            out = torch.empty(shape, dtype=torch.float32, device=self.device)
        _check(out, torch.float32, self.device, "out")
        mean, std = 0.0, 1.0
        if zmuv is not None:
            mean, std = float(zmuv[0]), float(zmuv[1])
            flag |= FE_ZMUV
        if rects is not None:
            _check(rects, torch.int32, self.device, "rects")
        self._note_fb(fb)
        self._rc(self.lib.howl_b200_frontend_fwd(self.handle, self._stream(), _ptr(pcm), b, t, _ptr(fb), mean, std,
                                                 _ptr(rects), flag, _ptr(out)), "frontend_fwd")
        return out

    def deltas(self, x: torch.Tensor) -> torch.Tensor:
        """``_execute_op(deltas_only=True)`` (transform.py:272-280): log-mels x [..., M, F] -> [B, 3, M, F] (B = the leading axes folded)."""
        _check(x, torch.float32, self.device, "x")
        if x.dim() < 2:
            raise HowlB200Error("deltas: x must be [..., M, F]")
        m, f = x.shape[-2], x.shape[-1]
        b = x.numel() // (m * f) if m * f else 0
        out = torch.empty((b, 3, m, f), dtype=torch.float32, device=self.device)
        self._rc(self.lib.howl_b200_deltas_fwd(self.handle, self._stream(), _ptr(x), b, m, f, _ptr(out)), "deltas_fwd")
        return out

    def sum_sumsq(self, x: torch.Tensor, sums: torch.Tensor) -> None:
        _check(x, torch.float32, self.device, "x")
        _check(sums, torch.float64, self.device, "sums")
        self._rc(self.lib.howl_b200_sum_sumsq(self.handle, self._stream(), _ptr(x), x.numel(), _ptr(sums)), "sum_sumsq")

    def zmuv(self, x: torch.Tensor, mean: float, std: float, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        _check(x, torch.float32, self.device, "x")
        out = torch.empty_like(x) if out is None else out
        self._rc(self.lib.howl_b200_zmuv_fwd(self.handle, self._stream(), _ptr(x), x.numel(), mean, std, _ptr(out)), "zmuv_fwd")
        return out

    def spec_mask(self, x: torch.Tensor, rects: torch.Tensor) -> torch.Tensor:
        _check(x, torch.float32, self.device, "x")
        _check(rects, torch.int32, self.device, "rects")
        b, c, m, f = x.shape
        self._rc(self.lib.howl_b200_spec_mask(self.handle, self._stream(), _ptr(x), b, c, m, f, _ptr(rects)), "spec_mask")
        return x

    def to_time_major(self, x: torch.Tensor) -> torch.Tensor:
        _check(x, torch.float32, self.device, "x")
        b, c, m, f = x.shape
        out = torch.empty(b, f, m, dtype=torch.float32, device=self.device)
        self._rc(self.lib.howl_b200_to_time_major(self.handle, self._stream(), _ptr(x), b, c, m, f, _ptr(out)), "to_time_major")
        return out

    def batch_gather(self, clips: torch.Tensor, starts: torch.Tensor, counts: torch.Tensor, dst_off: torch.Tensor,
                     max_length: int) -> torch.Tensor:
        _check(clips, torch.float32, self.device, "clips")
        for name, t in (("starts", starts), ("counts", counts), ("dst_off", dst_off)):
            _check(t, torch.int64, self.device, name)
        out = torch.empty(starts.numel(), max_length, dtype=torch.float32, device=self.device)
        self._rc(self.lib.howl_b200_batch_gather(self.handle, self._stream(), _ptr(clips), _ptr(starts), _ptr(counts),
                                                 _ptr(dst_off), starts.numel(), max_length, _ptr(out)), "batch_gather")
        return out

    def batch_gather_aug(self, clips, starts, counts, dst_off, max_length: int, bg=None, bg_starts=None, alpha=None, sigma=None,
                         sp_prob=None, seed: int = 0) -> torch.Tensor:
        """batch_gather with the waveform augmentations applied in the same pass; per-row parameter tensors on the
        device (float64 alpha, float32 sigma / sp_prob, int64 bg_starts), any of them None."""
        _check(clips, torch.float32, self.device, "clips")
        for name, t in (("starts", starts), ("counts", counts), ("dst_off", dst_off)) + ((("bg_starts", bg_starts),) if bg_starts is not None else ()):
            _check(t, torch.int64, self.device, name)
        for name, t in (("bg", bg), ("sigma", sigma), ("sp_prob", sp_prob)):
            if t is not None:
                _check(t, torch.float32, self.device, name)
        if alpha is not None:
            _check(alpha, torch.float64, self.device, "alpha")
        out = torch.empty(starts.numel(), max_length, dtype=torch.float32, device=self.device)
        self._rc(self.lib.howl_b200_batch_gather_aug(self.handle, self._stream(), _ptr(clips), _ptr(starts), _ptr(counts), _ptr(dst_off),
                                                     starts.numel(), max_length, _ptr(bg), _ptr(bg_starts), _ptr(alpha), _ptr(sigma),
                                                     _ptr(sp_prob), int(seed), _ptr(out)), "batch_gather_aug")
        return out

    # ------------------------------------------------------------------ res8
    def res8_param_count(self, num_labels: int) -> int:
        return int(self.lib.howl_b200_res8_param_count(num_labels))

    def res8_workspace_bytes(self, batch: int, frames: int, num_labels: int, train: bool = True) -> int:
        n = int(self.lib.howl_b200_res8_workspace_bytes(batch, frames, self.n_mels, num_labels, int(train)))
        if n < 0:
            raise HowlB200Error(f"res8: unsupported shape B={batch} frames={frames} mels={self.n_mels} L={num_labels}")
        return n

    def res8_fwd(self, feats, params, bn_running, nbt, train: bool, ws: torch.Tensor, logits=None):
        b, f, m = feats.shape
        num_labels = self._labels_from_params(params)
        if logits is None:
            logits = torch.empty(b, num_labels, dtype=torch.float32, device=self.device)
        self._rc(self.lib.howl_b200_res8_fwd(self.handle, self._stream(), _ptr(feats), b, f, m, num_labels, _ptr(params),
                                             _ptr(bn_running), _ptr(nbt), int(train), _ptr(logits), _ptr(ws),
                                             ws.numel()), "res8_fwd")
        return logits

    def res8_bwd(self, feats, labels, params, grads, loss, ws, loss_scale_batch: Optional[int] = None):
        b, f, m = feats.shape
        num_labels = self._labels_from_params(params)
        _check(labels, torch.int64, self.device, "labels")
        self._rc(self.lib.howl_b200_res8_bwd(self.handle, self._stream(), _ptr(feats), _ptr(labels), b, f, m, num_labels,
                                             loss_scale_batch or b, _ptr(params), _ptr(grads), _ptr(loss), _ptr(ws),
                                             ws.numel()), "res8_bwd")

    def res8_bwd_from_dlogits(self, feats, dlogits, params, grads, ws):
        b, f, m = feats.shape
        num_labels = self._labels_from_params(params)
        _check(dlogits, torch.float32, self.device, "dlogits")
        self._rc(self.lib.howl_b200_res8_bwd_dlogits(self.handle, self._stream(), _ptr(feats), _ptr(dlogits), b, f, m,
                                                     num_labels, _ptr(params), _ptr(grads), _ptr(ws), ws.numel()),
                 "res8_bwd_dlogits")

    def res8_debug_masks(self, feats, params, ws, conv0: bool = True):
        """Test hook (include/howl_b200_debug.h): the ReLU decisions of the backward for the forward kept in `ws`.
        -> (mask0 [B,45,3H,M] uint8 or None, masks [6,B,45,H,10] uint8)."""
        b, f, m = feats.shape
        h = f // 3
        num_labels = self._labels_from_params(params)
        mask0 = torch.empty(b, 45, 3 * h, m, dtype=torch.uint8, device=self.device) if conv0 else None
        masks = torch.empty(6, b, 45, h, 10, dtype=torch.uint8, device=self.device)
        self._rc(self.lib.howl_b200_res8_debug_masks(self.handle, self._stream(), _ptr(feats), _ptr(params), b, f, m, num_labels,
                                                     _ptr(ws), ws.numel(), _ptr(mask0), _ptr(masks)), "res8_debug_masks")
        return mask0, masks

    # ------------------------------------------------------------------ lstm / seq-lstm
    def lstm_param_count(self, num_labels: int) -> int:
        return int(self.lib.howl_b200_lstm_param_count(num_labels, self.n_mels))

    def lstm_workspace_bytes(self, batch: int, max_steps: int, num_labels: int, train: bool = True, sequential: bool = False) -> int:
        n = int(self.lib.howl_b200_lstm_workspace_bytes(batch, max_steps, self.n_mels, num_labels, int(train), int(sequential)))
        if n < 0:
            raise HowlB200Error(f"lstm: unsupported shape B={batch} steps={max_steps} L={num_labels}")
        return n

    def lstm_labels_from_params(self, params: torch.Tensor) -> int:
        n = params.numel() - (512 * self.n_mels + 512 * 128 + 1024 + 256 * 128 + 256)
        if n <= 0 or n % 257:
            raise HowlB200Error(f"lstm: flat parameter buffer of {params.numel()} floats is not an lstm layout")
        return n // 257

    def lstm_fwd(self, feats, lengths, max_steps: int, params, ws, sequential=False, train=False, state_in=None,
                 state_out=None, out=None):
        b, f, m = feats.shape
        num_labels = self.lstm_labels_from_params(params)
        _check(lengths, torch.int64, self.device, "lengths")
        if out is None:
            shape = (max_steps, b, num_labels) if sequential else (b, num_labels)
            out = torch.empty(shape, dtype=torch.float32, device=self.device)
        self._rc(self.lib.howl_b200_lstm_fwd(self.handle, self._stream(), _ptr(feats), _ptr(lengths), b, f, m, num_labels,
                                             max_steps, _ptr(params), _ptr(state_in), _ptr(state_out), int(sequential),
                                             int(train), _ptr(out), _ptr(ws), ws.numel()), "lstm_fwd")
        return out

    def lstm_bwd(self, feats_shape, lengths, max_steps: int, labels, params, grads, loss, ws, loss_scale_batch=None, dlogits=None,
                 sequential=False):
        b, f, m = feats_shape
        num_labels = self.lstm_labels_from_params(params)
        if dlogits is not None:
            _check(dlogits, torch.float32, self.device, "dlogits")
            self._rc(self.lib.howl_b200_lstm_bwd_dlogits(self.handle, self._stream(), _ptr(lengths), _ptr(dlogits),
                                                         int(sequential), b, f, m, num_labels, max_steps, _ptr(params),
                                                         _ptr(grads), _ptr(ws), ws.numel()), "lstm_bwd_dlogits")
        else:
            _check(labels, torch.int64, self.device, "labels")
            self._rc(self.lib.howl_b200_lstm_bwd(self.handle, self._stream(), _ptr(lengths), _ptr(labels), b, f, m, num_labels,
                                                 max_steps, loss_scale_batch or b, _ptr(params), _ptr(grads), _ptr(loss),
                                                 _ptr(ws), ws.numel()), "lstm_bwd")

    def lstm_train_step(self, pcm, labels, lengths, max_steps, fb, zmuv, params, grads, m, v, step, lr, weight_decay, loss,
                        logits, ws):
        b, t = pcm.shape
        num_labels = self.lstm_labels_from_params(params)
        self._note_pcm(pcm)
        _check(labels, torch.int64, self.device, "labels")
        _check(lengths, torch.int64, self.device, "lengths")
        for name, t_ in (("fb", fb), ("params", params), ("grads", grads), ("m", m), ("v", v), ("loss", loss), ("logits", logits)):
            _check(t_, torch.float32, self.device, name)
        self._note_fb(fb)
        self._rc(self.lib.howl_b200_lstm_train_step(
            self.handle, self._stream(), _ptr(pcm), _ptr(labels), _ptr(lengths), b, t, _ptr(fb), float(zmuv[0]),
            float(zmuv[1]), num_labels, max_steps, _ptr(params), _ptr(grads), _ptr(m), _ptr(v), step, lr, weight_decay,
            _ptr(loss), _ptr(logits), _ptr(ws), ws.numel()), "lstm_train_step")

    def lstm_ctc_bwd(self, feats_shape, lengths, max_steps, targets, target_lengths, blank, params, grads, loss, ws,
                     loss_scale_batch=None):
        b, f, m = feats_shape
        num_labels = self.lstm_labels_from_params(params)
        _check(targets, torch.int64, self.device, "targets")
        _check(target_lengths, torch.int64, self.device, "target_lengths")
        self._rc(self.lib.howl_b200_lstm_ctc_bwd(self.handle, self._stream(), _ptr(lengths), _ptr(targets), _ptr(target_lengths),
                                                 targets.shape[1], blank, b, f, m, num_labels, max_steps, loss_scale_batch or b,
                                                 _ptr(params), _ptr(grads), _ptr(loss), _ptr(ws), ws.numel()), "lstm_ctc_bwd")

    def seq_lstm_ctc_train_step(self, pcm, targets, target_lengths, blank, lengths, max_steps, fb, zmuv, params, state, grads, m,
                                v, step, lr, weight_decay, loss, scores, ws):
        b, t = pcm.shape
        num_labels = self.lstm_labels_from_params(params)
        self._note_pcm(pcm)
        for name, t_ in (("targets", targets), ("target_lengths", target_lengths), ("lengths", lengths)):
            _check(t_, torch.int64, self.device, name)
        for name, t_ in (("fb", fb), ("params", params), ("state", state), ("grads", grads), ("m", m), ("v", v), ("loss", loss),
                         ("scores", scores)):
            _check(t_, torch.float32, self.device, name)
        self._note_fb(fb)
        self._rc(self.lib.howl_b200_seq_lstm_ctc_train_step(
            self.handle, self._stream(), _ptr(pcm), _ptr(targets), _ptr(target_lengths), targets.shape[1], blank, _ptr(lengths),
            b, t, _ptr(fb), float(zmuv[0]), float(zmuv[1]), num_labels, max_steps, _ptr(params), _ptr(state), _ptr(grads),
            _ptr(m), _ptr(v), step, lr, weight_decay, _ptr(loss), _ptr(scores), _ptr(ws), ws.numel()), "seq_lstm_ctc_train_step")

    def seq_lstm_train_step_workspace_bytes(self, batch: int, samples: int, max_steps: int, num_labels: int) -> int:
        f = self.num_frames(samples)
        feat = (batch * f * self.n_mels * 4 + 255) // 256 * 256
        return feat + self.lstm_workspace_bytes(batch, max_steps, num_labels, True, True)

    def lstm_train_step_workspace_bytes(self, batch: int, samples: int, max_steps: int, num_labels: int) -> int:
        f = self.num_frames(samples)
        feat = (batch * f * self.n_mels * 4 + 255) // 256 * 256
        return feat + self.lstm_workspace_bytes(batch, max_steps, num_labels, True, False)

    def adamw(self, params, grads, m, v, step: int, lr: float, weight_decay: float, betas=(0.9, 0.999), eps=1e-8):
        self._rc(self.lib.howl_b200_adamw(self.handle, self._stream(), _ptr(params), _ptr(grads), _ptr(m), _ptr(v),
                                          params.numel(), step, lr, betas[0], betas[1], eps, weight_decay), "adamw")

    def res8_train_step(self, pcm, labels, fb, zmuv, params, bn_running, nbt, grads, m, v, step, lr, weight_decay,
                        loss, logits, ws, rects=None):
        self._note_pcm(pcm)
        _check(labels, torch.int64, self.device, "labels")
        for name, t_ in (("fb", fb), ("params", params), ("bn_running", bn_running), ("grads", grads), ("m", m), ("v", v)):
            _check(t_, torch.float32, self.device, name)
        if pcm.dim() != 2 or labels.numel() != pcm.shape[0]:
            raise HowlB200Error("res8_train_step: pcm must be [B,T] and labels [B]")
        if rects is not None:
            _check(rects, torch.int32, self.device, "rects")
            if tuple(rects.shape) != (pcm.shape[0], 4):
                raise HowlB200Error("res8_train_step: rects must be [B,4]")
        _check(nbt, torch.int64, self.device, "nbt")
        _check(loss, torch.float32, self.device, "loss")
        _check(logits, torch.float32, self.device, "logits")
        _check(ws, torch.uint8, self.device, "ws")
        if logits.numel() < pcm.shape[0] * self._labels_from_params(params):
            raise HowlB200Error("res8_train_step: logits buffer smaller than [B, num_labels]")
        b, t = pcm.shape
        num_labels = self._labels_from_params(params)
        self._note_fb(fb)
        self._rc(self.lib.howl_b200_res8_train_step(
            self.handle, self._stream(), _ptr(pcm), _ptr(labels), b, t, _ptr(fb), float(zmuv[0]), float(zmuv[1]),
            _ptr(rects), num_labels, _ptr(params), _ptr(bn_running), _ptr(nbt), _ptr(grads), _ptr(m), _ptr(v), step,
            lr, weight_decay, _ptr(loss), _ptr(logits), _ptr(ws), ws.numel()), "res8_train_step")

    def train_step_workspace_bytes(self, batch: int, samples: int, num_labels: int) -> int:
        f = self.num_frames(samples)
        feat = (batch * f * self.n_mels * 4 + 255) // 256 * 256
        return feat + self.res8_workspace_bytes(batch, f, num_labels, True)

    @staticmethod
    def _labels_from_params(params: torch.Tensor) -> int:
        n = params.numel() - 45 * 9 - 6 * 45 * 45 * 9
        if n <= 0 or n % 46:
            raise HowlB200Error(f"res8: flat parameter buffer of {params.numel()} floats is not a res8 layout")
        return n // 46
