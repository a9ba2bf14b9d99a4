#!/usr/bin/env python
"""bench.py -- train-step throughput of the howl hot path on N x B200 (utterances / s), with roofline and baselines.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--model res8|lstm|seq-lstm|mobilenet|las] [--batch B] [--seconds S] [--scaling weak|strong] [--global-batch G]

A "step" is one pass of the hot path over one batch of synthetic 16 kHz clips (training/run/train.py:287-302 of the reference):
frontend (STFT -> mel -> log -> ZMUV) -> model forward -> loss -> backward -> [allreduce] -> AdamW.

Default workload = BASELINE.json configs[1]: res8, NUM_MELS=40, 1 s clips, batch 4096 per GPU (weak scaling; configs[4] at N=8;
`--scaling strong --global-batch 32768` is the fixed-global-batch half of configs[4]).  The other configs are reachable with flags:
  configs[0]  --model res8 --seconds 0.5 --batch 64            (the reference's own CPU-runnable case; `--impl reference` times it on the CPU)
  configs[2]  --model mobilenet                                (MobileNetV2, batch 8192, bf16)
  configs[3]  --model seq-lstm                                 (streaming seq-lstm + CTC, 0.5 s clips, batch 2048)

  value                : whole-job utterances/s with PCM already resident in HBM (device-timed with CUDA events, max over ranks)
  e2e                  : same metric through the public Python API with HOST (pinned) inputs, H2D copies and a D2H read of the loss
                         inside the timed region (copy stream, triple buffered)
  roofline             : the dominant kernel FAMILY of the step against its binding roof: tensor pipe for the conv / GEMM
                         kernels, HBM for the frontend; whole-step fractions beside it
  cpu_baseline         : the oracle port of the reference's torch-CPU path on this box's host cores (rank 0, N=1, bounded sample)
  gpu_library_baseline : the same oracle graph on cuda:0 through stock torch (cuDNN / cuFFT / cuBLAS), TF32 off and on -- the
                         "library-kernel bar" of BASELINE.md §2 step 5
  --impl reference     : the reference's CPU path (oracle port) alone, same metric / config.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

N_MELS = 40
LR, WD = 0.01, 1e-5
ZMEAN, ZSTD = -2.0166, 3.9955   # ZMUV constants of the reference's GSC res8 run
FP32_PEAK_TFLOPS = 148 * 128 * 2 * 1.965e9 / 1e12      # FFMA roof: 148 SMs x 128 lanes x 2 flop x max clock
SR, N_FFT, HOP = 16000, 512, 200

MODELS = {
    # labels: GSC-12 semantics for res8; hey-fire-fox + blank for the CTC config (config 4)
    "res8": {"seconds": 1.0, "batch": 4096, "labels": 12, "dtype": "f32"},
    "lstm": {"seconds": 0.5, "batch": 2048, "labels": 5, "dtype": "f32"},
    "seq-lstm": {"seconds": 0.5, "batch": 2048, "labels": 5, "blank": 4, "dtype": "f32"},
    "mobilenet": {"seconds": 1.0, "batch": 8192, "labels": 12, "dtype": "bf16"},
    "las": {"seconds": 1.0, "batch": 2048, "labels": 12, "dtype": "f32"},     # not a BASELINE config: SURVEY §8 row a12, for completeness
}


def frames_of(samples):
    return 1 + samples // HOP


def algorithmic(model, samples, labels, batch):
    """Algorithmic flops / HBM bytes per utterance of one train step (derivations in DESIGN.md §4)."""
    F = frames_of(samples)
    if model == "res8":
        H = F // 3
        mac_fwd = 45 * 9 * F * 40 + 6 * 45 * 45 * 9 * H * 10 + 45 * labels
        mac = mac_fwd + 45 * 9 * F * 40 + 2 * 6 * 45 * 45 * 9 * H * 10 + 2 * 45 * labels   # bwd = wgrad(conv0) + 2 x conv1-6
        nparam = 45 * 9 + 6 * 45 * 45 * 9 + 46 * labels
        return {"flop": 2.0 * mac, "bytes": samples * 4 + 8 + 4 * labels + 24.0 * nparam / batch,
                "conv_layer_flop": 2.0 * 45 * 45 * 9 * H * 10}
    if model in ("lstm", "seq-lstm"):
        steps = (samples - N_FFT) // HOP + 1
        head = (128 * 256 + 256 * labels) * (steps if model == "seq-lstm" else 1)
        mac = steps * 512 * (N_MELS + 128) + head
        nparam = 512 * (N_MELS + 128) + 1024 + 128 * 256 + 256 + 257 * labels
        return {"flop": 3 * 2.0 * mac, "bytes": samples * 4 + 8 + 4 * labels + 24.0 * nparam / batch, "steps": steps}
    if model == "mobilenet":
        from howl_b200.mobilenet import algorithmic_flops   # exists once the a10 row is built
        return algorithmic_flops(samples, labels, batch)
    if model == "las":
        h1, w1 = N_MELS + 2, F + 2
        h2, w2 = h1 + 2, w1 // 2 + 2
        T, inp = w2 // 2, 8 * h2
        mac = 8 * 27 * h1 * w1 + 8 * 72 * h2 * w2 + T * 2 * 384 * (inp + 96) + 2 * T * 192 * 192 + 192 * 256 + 256 * labels
        nparam = 8 * 27 + 8 * 72 + 32 + 2 * (384 * (inp + 96) + 768) + 192 + 2 * 192 * 193 + 256 * 193 + 257 * labels
        return {"flop": 3 * 2.0 * mac, "bytes": samples * 4 + 8 + 4 * labels + 24.0 * nparam / batch}
    raise ValueError(model)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--model", default="res8", choices=sorted(MODELS))
    ap.add_argument("--batch", type=int, default=None, help="utterances per GPU per step (default: the BASELINE config of the model)")
    ap.add_argument("--seconds", type=float, default=None, help="clip length (default: the BASELINE config of the model)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--global-batch", type=int, default=32768, help="--scaling strong: the fixed global batch (configs[4])")
    ap.add_argument("--cpu-sample", type=int, default=256, help="utterances per CPU-baseline step")
    ap.add_argument("--pcm32", action="store_true", help="e2e leg ships float32 PCM instead of int16 (the wav files' own sample format, "
                                                        "which the C ABI accepts directly: K1 converts with x / 32768 like the reference's loader)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gpu-library-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-module-path", action="store_true", help="skip the module-by-module (train.py loop body) leg")
    ap.add_argument("--breakdown", action="store_true", help="print per-kernel-group CUDA-event times to stderr")
    a = ap.parse_args()
    spec = MODELS[a.model]
    a.seconds = spec["seconds"] if a.seconds is None else a.seconds
    a.samples = int(round(a.seconds * SR))
    a.labels = spec["labels"]
    a.pcm16 = not a.pcm32
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if a.scaling == "strong":
        if a.global_batch % world:
            ap.error("--global-batch must divide by the number of GPUs")
        a.batch = a.global_batch // world
    elif a.batch is None:
        a.batch = spec["batch"]
    return a


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"], "bf16_tflops_sustained": d.get("bf16_tflops_sustained"),
                "source": "measured (MEASURED_PEAKS.json)"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """nvidia-smi clock / throttle sampling during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                          "-lms", "50"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            for n, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        mx = max((int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()), default=None)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def synth(model, batch, samples, labels, seed):
    """Synthetic inputs of SURVEY §8(d): noise at speech-like RMS clamped to [-1, 1], uniform labels (CTC: 1..3 target labels)."""
    gen = torch.Generator(device="cpu").manual_seed(seed)
    pcm = (torch.randn(batch, samples, generator=gen) * 0.1).clamp_(-1, 1)
    if model == "seq-lstm":
        blank = MODELS[model]["blank"]
        return (pcm, torch.randint(0, blank, (batch, 3), generator=gen), torch.randint(1, 4, (batch,), generator=gen))
    return (pcm, torch.randint(0, labels, (batch,), generator=gen))


# ---------------------------------------------------------------------------------------------------------------------
# baselines: the oracle port of the reference's torch path, on the host cores (cpu_baseline / --impl reference) and on cuda:0
# through stock torch kernels (gpu_library_baseline)
# ---------------------------------------------------------------------------------------------------------------------
def oracle_step_factory(model, batch, samples, labels, device="cpu"):
    from oracle import howl_oracle as O  # the checker / baseline -- never on the product path

    dev = torch.device(device)
    inputs = [t.to(dev) for t in synth(model, batch, samples, labels, seed=0)]
    fb = O.mel_filterbank(N_MELS).to(dev)
    zm, zm2 = torch.tensor([ZMEAN], device=dev), torch.tensor([ZMEAN ** 2 + ZSTD ** 2], device=dev)
    state = {"step": 0, "hc": None}
    if model == "mobilenet":
        from howl_b200 import mobilenet as mb   # host-side shapes / initialiser only (no CUDA)

        flat, sd, off = mb.init_flat(labels, 0), {}, 0
        for name, shape in mb.param_shapes(labels):
            n = int(torch.tensor(shape).prod())
            sd[name] = flat[off:off + n].view(shape).clone().to(dev)
            off += n
        for _, _, bnn, shape, _ in mb.layer_plan(labels):
            sd[bnn + ".running_mean"], sd[bnn + ".running_var"] = torch.zeros(shape[0], device=dev), torch.ones(shape[0], device=dev)
        params = {k: sd[k] for k in O.mobilenet_param_names(sd)}
    elif model == "las":
        from howl_b200 import las   # host-side shapes only (no CUDA)

        g = torch.Generator().manual_seed(0)
        sd = {n: ((torch.rand(sh, generator=g) * 2 - 1) / (max(int(torch.tensor(sh[1:]).prod()), 1) ** 0.5 if len(sh) > 1 else 10.0)).to(dev)
              for n, sh in las.param_shapes(labels, N_MELS)}
        for idx in ("1", "5"):
            sd[f"encoder.conv_encoder.{idx}.weight"] = torch.ones(8, device=dev)
        params = dict(sd)
        las_len = torch.full((batch,), (samples - N_FFT) // HOP + 1, dtype=torch.int64)
    elif model == "res8":
        params = {k: v.to(dev) for k, v in O.res8_init(labels, seed=0).items()}
        bn = {k: v.to(dev) for k, v in O.res8_bn_init().items()}
    else:
        params = {k: v.to(dev) for k, v in O.lstm_init(labels, seed=0).items()}
        steps = (samples - N_FFT) // HOP + 1
        lengths = torch.full((batch,), steps, dtype=torch.int64, device=dev)
    m = {k: torch.zeros_like(p) for k, p in params.items()}
    v = {k: torch.zeros_like(p) for k, p in params.items()}

    def step():
        state["step"] += 1
        feats = O.hot_path_features(inputs[0], fb, zm, zm2)
        if model == "mobilenet":
            with torch.autocast(dev.type, dtype=torch.bfloat16, enabled=state.get("autocast", False)):
                loss, _, grads = O.mobilenet_grads(feats, inputs[1], sd)
            with torch.no_grad():
                O.adamw_step(params, grads, m, v, state["step"], LR, WD)
                for k in params:
                    sd[k] = params[k]
        elif model == "las":
            loss, _, grads = O.las_grads(feats, inputs[1], params, las_len, dtype=torch.float32)
            with torch.no_grad():
                O.adamw_step(params, grads, m, v, state["step"], LR, WD)
        elif model == "res8":
            loss, _, _ = O.res8_train_step(feats, inputs[1], params, bn, m, v, state["step"], LR, WD)
        elif model == "lstm":
            loss, _, _ = O.lstm_train_step(feats, inputs[1], lengths, params, m, v, state["step"], LR, WD)
        else:
            loss, _, _, hc = O.seq_lstm_ctc_step(feats, inputs[1], inputs[2], lengths, params, state["hc"], MODELS[model]["blank"], m, v,
                                                 state["step"], 1e-4, WD)
            state["hc"] = tuple(t.detach() for t in hc)
        return loss

    step.state = state
    return step


def run_cpu(model, batch, samples, labels, steps, warmup):
    """Times the oracle port on the host cores with torch's intra-op pool at a FIXED size chosen once: on a many-core box the
    small convolutions of res8 run slower with every core than with a few dozen, so {8, 16, 32, 64, all} are each timed over
    3 steps after a warm-up step and the fastest is used (and reported as `cores`)."""
    ncpu = os.cpu_count() or 1
    step = oracle_step_factory(model, batch, samples, labels)
    forced = os.environ.get("HOWL_CPU_THREADS")
    best, best_t = (int(forced), 0.0) if forced else (ncpu, None)
    if not forced:
        for n in sorted({min(c, ncpu) for c in (8, 16, 32, 64, ncpu)}):
            torch.set_num_threads(n)
            step()
            t0 = time.perf_counter()
            for _ in range(3):
                step()
            dt = time.perf_counter() - t0
            if best_t is None or dt < best_t:
                best, best_t = n, dt
    torch.set_num_threads(best)
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    return batch * steps / dt, dt / steps * 1e3, best


def run_gpu_library(model, batch, samples, labels, dev, steps=5, warmup=2):
    """The oracle graph through stock torch on the GPU (cuFFT stft, cuDNN convolutions / cuBLAS GEMMs, torch autograd, the oracle's
    AdamW restatement as elementwise torch ops), TF32 off (fp32 parity arithmetic) and on (torch's default for convolutions)."""
    out = {}
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    try:
        for tf32 in (False, True):
            torch.backends.cudnn.allow_tf32 = tf32
            torch.backends.cuda.matmul.allow_tf32 = tf32
            step = oracle_step_factory(model, batch, samples, labels, device=str(dev))
            step.state["autocast"] = bool(tf32 and model == "mobilenet")        # mobilenet: the "on" leg is torch's bf16 autocast
            for _ in range(warmup):
                step()
            torch.cuda.synchronize(dev)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                step()
            e1.record()
            torch.cuda.synchronize(dev)
            ms = e0.elapsed_time(e1) / steps
            out["tf32_on" if tf32 else "tf32_off"] = {"value": batch / ms * 1e3, "unit": "utterances/s", "ms_per_step": ms}
            del step
            torch.cuda.empty_cache()
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old
    out["what"] = (f"oracle graph of the same train step on cuda:0 through stock torch {torch.__version__} (cuFFT / cuDNN / cuBLAS, eager "
                   f"autograd), batch {batch}, {steps} steps after {warmup} warm-up; PCM resident on the device"
                   + ("; tf32_on = TF32 convolutions + torch.autocast(bfloat16)" if model == "mobilenet" else ""))
    return out


def run_module_path(a, dev, pcm, labels, steps=10, warmup=3):
    """The loop body of training/run/train.py:288-302 through the drop-in MODULES (what the reference's own train.py executes after
    `plugin.install()`): StandardAudioTransform -> ZmuvTransform -> registry model (nn.Module, autograd Function) -> torch's
    CrossEntropyLoss -> backward -> torch.optim.AdamW over model.parameters().  Same batch as the fused step; PCM resident in HBM."""
    from howl_b200.model import RegisteredModel
    from howl_b200.transform import StandardAudioTransform, ZmuvTransform

    os.environ["NUM_MELS"] = str(N_MELS)
    from howl_b200.settings import SETTINGS
    SETTINGS.reset()
    model = RegisteredModel.find_registered_class(a.model)(a.labels).to(dev).streaming()
    std = StandardAudioTransform().to(dev).eval()
    zmuv = ZmuvTransform().to(dev)
    zmuv.mean = torch.tensor([ZMEAN], device=dev)
    zmuv.mean2 = torch.tensor([ZMEAN ** 2 + ZSTD ** 2], device=dev)
    opt = torch.optim.AdamW(model.parameters(), LR, weight_decay=WD)
    crit = torch.nn.CrossEntropyLoss()
    model.train()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for i in range(warmup + steps):
        if i == warmup:
            torch.cuda.synchronize(dev)
            ev0.record()
        lengths = std.compute_lengths(torch.full((pcm.size(0),), pcm.size(1)))
        scores = model(zmuv(std(pcm)), lengths)
        loss = crit(scores, labels)
        opt.zero_grad()
        model.zero_grad()
        loss.backward()
        opt.step()
    ev1.record()
    torch.cuda.synchronize(dev)
    ms = ev0.elapsed_time(ev1) / steps
    return {"value": pcm.size(0) / (ms / 1e3), "unit": "utterances/s", "ms_per_step": ms, "last_loss": float(loss.item()),
            "what": "train.py:288-302 loop body through the drop-in modules (StandardAudioTransform, ZmuvTransform, registry nn.Module under "
                    f"autograd, torch CrossEntropyLoss + torch.optim.AdamW), batch {pcm.size(0)}, {steps} steps after {warmup} warm-up; the "
                    "stacked [B,3,M,F] features and torch's optimizer are part of it -- the fused step (`value`) is the same math in one call"}


def workload_text(a, world):
    per = {"res8": "fused STFT->mel->conv train step", "lstm": "frontend + LSTM(40->128) + MLP train step (frame objective)",
           "seq-lstm": "frontend + streaming seq-lstm + CTC train step", "mobilenet": "frontend + MobileNetV2 train step",
           "las": "frontend + LASClassifier (convs, BiLSTM(352->96), attention, MLP) train step"}[a.model]
    return (f"{a.model} NUM_MELS={N_MELS} batch={a.batch}/GPU {per}, synthetic GSC-shaped {a.seconds:g} s clips, L={a.labels}")


def main_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warmup = max(1, a.steps), max(0, a.warmup)
    sample = min(a.cpu_sample, a.batch)
    value, ms, cores = run_cpu(a.model, sample, a.samples, a.labels, steps, warmup)
    line = {
        "impl": "reference", "metric": f"{a.model} train-step throughput ({a.seconds:g} s @16 kHz clips)", "value": value,
        "unit": "utterances/s", "n_gpus": a.gpus, "steps": steps, "warmup": warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": a.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_text(a, 1) + f"; oracle port of the reference torch-CPU path, {sample} utterances per step "
                                                     f"(bounded sample of the batch-{a.batch} workload)"},
        "cpu_baseline": {"value": value, "unit": "utterances/s", "cores": cores, "kind": "port",
                         "sample": f"{steps} steps x {sample} utterances after {warmup} warm-up, torch CPU {torch.__version__}, "
                                   f"{cores} intra-op threads of {os.cpu_count()} host cores"},
        "e2e": {"value": value, "unit": "utterances/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------------------------
def make_step(a, dev, world):
    from howl_b200 import trainer as T

    kw = dict(num_labels=a.labels, batch=a.batch, samples=a.samples, n_mels=N_MELS, weight_decay=WD, zmuv=(ZMEAN, ZSTD), seed=0,
              world_size=world)
    if a.model == "res8":
        return T.Res8TrainStep(dev, lr=LR, **kw)
    if a.model == "lstm":
        return T.LstmTrainStep(dev, lr=LR, **kw)
    if a.model == "seq-lstm":
        return T.SeqLstmCtcTrainStep(dev, blank=MODELS[a.model]["blank"], lr=1e-4, **kw)
    if a.model == "mobilenet":
        return T.MobileNetTrainStep(dev, lr=LR, **kw)
    if a.model == "las":
        return T.LasTrainStep(dev, lr=LR, **kw)
    raise ValueError(a.model)


def family_of(name):
    if name.startswith("conv3x3"):
        return "conv3x3"
    if name.startswith(("lstm_", "ctc")):
        return "lstm"
    if name.startswith(("mbn_gemm", "mbn_wgrad")):
        return "mbn_gemm"
    if name.startswith(("las_lstm", "las_gemm")):
        return "las_lstm"
    return name


def load_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full` capture of this round
    (profiles/r02_traffic.json, written by tools/ncu_traffic.py from the .ncu-rep); null when absent."""
    for name in ("r02_traffic.json",):
        p = os.path.join(ROOT, "profiles", name)
        if os.path.exists(p):
            return json.load(open(p)), "profiles/" + name
    return {}, None


def main_ours(a):
    import torch.distributed as dist

    import howl_b200  # noqa: F401

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")   # keep stdout to the one JSON line
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    B = a.batch
    # one distinct batch per in-flight slot
    nbuf = 3
    host = [tuple(t.pin_memory() for t in synth(a.model, B, a.samples, a.labels, 1234 + 17 * rank + i)) for i in range(nbuf)]
    devi = [tuple(t.to(dev) for t in h) for h in host]
    if a.pcm16:
        host = [((h[0] * 32768.0).round().clamp_(-32768, 32767).to(torch.int16).pin_memory(),) + tuple(h[1:]) for h in host]

    step_obj = make_step(a, dev, world)
    if os.environ.get("HOWL_CONV_ENGINE"):   # tuning experiments only
        step_obj.ctx.set_option("conv_engine", int(os.environ["HOWL_CONV_ENGINE"]))
    if os.environ.get("HOWL_LSTM_ENGINE"):   # A/B of the recurrences: 0 = plain, 1 = software pipelined, 2 = pipelined + 2 x 8 register tile (default)
        step_obj.ctx.set_option("lstm_engine", int(os.environ["HOWL_LSTM_ENGINE"]))
    ctx = step_obj.ctx
    warmup = max(a.warmup, 3)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def maxed(ms):
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident timing (value)
    for i in range(warmup):
        step_obj.step(*devi[i % nbuf])
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    launches0 = ctx.launch_count
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for i in range(a.steps):
        step_obj.step(*devi[i % nbuf])
    ev1.record()
    barrier()
    launches = ctx.launch_count - launches0
    ms_total = maxed(ev0.elapsed_time(ev1))
    clocks = sampler.stop() if sampler else None
    ms_step = ms_total / a.steps
    value = B * world * a.steps / (ms_total / 1e3)
    last_loss_dev = float(step_obj.loss.item())

    # ---- per-kernel device times for the roofline (live, CUDA events on the launching stream)
    groups = step_obj.profile_groups(*devi[0], reps=3)
    barrier()

    # ---- end to end through the public API with host buffers
    e2e = None
    if not a.no_e2e:
        for i in range(3):
            step_obj.step_host(*host[i % nbuf])
        step_obj.flush_host()
        barrier()
        t0 = time.perf_counter()
        ev0.record()
        for i in range(a.steps):
            step_obj.step_host(*host[i % nbuf])
        last_loss = step_obj.flush_host()
        ev1.record()
        barrier()
        e_ms = maxed(ev0.elapsed_time(ev1))
        e2e = {"value": B * world * a.steps / (e_ms / 1e3), "unit": "utterances/s",
               "h2d_bytes_per_step": sum(t.numel() * t.element_size() for t in host[0]), "d2h_bytes_per_step": 4,
               "wall_s": time.perf_counter() - t0, "last_loss": last_loss, "pcm_dtype": str(host[0][0].dtype).replace("torch.", "")}

    if rank == 0:
        peaks = load_peaks()
        alg = algorithmic(a.model, a.samples, a.labels, B)
        tensor_peak = peaks["bf16_tflops_sustained"] or peaks["bf16_tflops"]
        traffic, traffic_src = load_traffic()
        fam = {}
        for g in groups:
            f = fam.setdefault(family_of(g["name"]), {"ms": 0.0, "launches": 0, "kernels": []})
            f["ms"] += g["ms"]
            f["launches"] += g["launches_per_step"]
            f["kernels"].append(g["name"])
        step_tf = B * alg["flop"] / (ms_step / 1e3) / 1e12
        step_gbs = B * alg["bytes"] / (ms_step / 1e3) / 1e9
        if a.model == "res8":
            engine = "tcgen05 bf16x3-split MMA, fp32 TMEM accumulate" if any("_tc" in g["name"] for g in groups) else "fp32 FFMA"
            cf = fam["conv3x3"]
            per_launch_ms = cf["ms"] / cf["launches"]
            ach = B * alg["conv_layer_flop"] / (per_launch_ms / 1e3) / 1e12
            tr = [traffic[k] for k in cf["kernels"] if k in traffic]
            roofline = {"bound": "tensor", "kernel": f"conv3x3 45->45 family ({cf['launches']} launches/step: fwd + dgrad + wgrad x 6 layers), {engine}",
                        "achieved": ach, "peak": tensor_peak, "unit": "TFLOP/s", "frac": ach / tensor_peak,
                        "traffic": (sum(tr) / len(tr) if tr else None), "traffic_source": traffic_src if tr else None,
                        "algorithmic_flop_per_launch": B * alg["conv_layer_flop"], "avg_launch_ms": per_launch_ms,
                        "family_ms_per_step": cf["ms"], "family_share_of_step": cf["ms"] / sum(g["ms"] for g in groups),
                        "peak_source": peaks["source"] + ", dense bf16 cuBLAS sustained",
                        "note": "fp32-equivalent flops; the issued bf16 MMA flops are 3x (hi*hi + hi*lo + lo*hi)"}
        else:
            key = {"lstm": "lstm", "seq-lstm": "lstm", "las": "las_lstm"}.get(a.model, "mbn_gemm")
            cf = fam.get(key, {"ms": ms_step, "launches": 1, "kernels": []})
            ach = B * alg.get("gemm_flop", alg["flop"]) / (cf["ms"] / 1e3) / 1e12
            roofline = {"bound": "tensor", "kernel": f"{key} family ({cf['launches']} launches/step)", "achieved": ach, "peak": tensor_peak,
                        "unit": "TFLOP/s", "frac": ach / tensor_peak, "traffic": None, "family_ms_per_step": cf["ms"],
                        "peak_source": peaks["source"] + ", dense bf16 cuBLAS sustained"}
        roofline["step"] = {"tflops_algorithmic": step_tf, "tensor_frac": step_tf / tensor_peak, "hbm_gbs_algorithmic": step_gbs,
                            "hbm_frac": step_gbs / peaks["hbm_gbs"], "fp32_ffma_peak_tflops": FP32_PEAK_TFLOPS,
                            "dram_bytes_measured": traffic.get("__step_total__"), "dram_bytes_source": traffic_src,
                            "binding": "tensor (SURVEY 8d: 2.8 kFLOP/B algorithmic intensity)"}
        fe = next((g for g in groups if g["name"] == "frontend"), None)
        if fe:
            gbs = B * (a.samples * 4 + frames_of(a.samples) * N_MELS * 4) / (fe["ms"] / 1e3) / 1e9
            # 1.0 MFLOP per 1 s utterance (FFT 0.93 + sparse mel 0.08), scaled with the frame count
            fe_flop = B * 1.0e6 * frames_of(a.samples) / 81.0
            roofline["frontend"] = {"bound": "hbm", "ms": fe["ms"], "achieved": gbs, "unit": "GB/s", "peak": peaks["hbm_gbs"],
                                    "frac": gbs / peaks["hbm_gbs"], "traffic": traffic.get("frontend"),
                                    "fp32_tflops_algorithmic": fe_flop / (fe["ms"] / 1e3) / 1e12,
                                    "fp32_frac": fe_flop / (fe["ms"] / 1e3) / 1e12 / FP32_PEAK_TFLOPS,
                                    "note": "HBM is the roof SURVEY 8(d) names for K1; at 13-34 FLOP/B (twice that with int16 PCM) the kernel sits at "
                                            "the FP32 ridge and is bound by instruction issue of the FFT (DESIGN.md section 4)"}
        cpu = None
        if not a.no_cpu_baseline and world == 1:
            sample = min(a.cpu_sample, B)
            v, ms, cores = run_cpu(a.model, sample, a.samples, a.labels, 10, 2)
            cpu = {"value": v, "unit": "utterances/s", "cores": cores, "kind": "port",
                   "sample": f"10 steps x {sample} utterances of the same workload after 2 warm-up (oracle port, torch CPU), "
                             f"{cores} intra-op threads of {os.cpu_count()} host cores"}
        modules = None
        if a.model == "res8" and world == 1 and not a.no_module_path:
            try:
                modules = run_module_path(a, dev, devi[0][0], devi[0][1])
            except Exception as exc:   # noqa: BLE001 -- a side leg must not take the measurement down
                modules = {"error": repr(exc)[:300]}
        gpu_lib = None
        if not a.no_gpu_library_baseline and world == 1:
            del step_obj
            torch.cuda.empty_cache()
            try:
                gpu_lib = run_gpu_library(a.model, B, a.samples, a.labels, dev)
                gpu_lib["speedup_over_tf32_off"] = value / gpu_lib["tf32_off"]["value"]
                gpu_lib["speedup_over_tf32_on"] = value / gpu_lib["tf32_on"]["value"]
            except Exception as exc:   # noqa: BLE001 -- a baseline leg must not take the measurement down
                gpu_lib = {"error": repr(exc)[:300]}
        line = {
            "metric": f"{a.model} train-step throughput ({a.seconds:g} s @16 kHz clips)", "value": value, "unit": "utterances/s",
            "n_gpus": world, "steps": a.steps, "warmup": warmup, "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": a.scaling, "vs_baseline": None, "dtype": MODELS[a.model]["dtype"], "data": "synthetic",
            "config": {"workload": workload_text(a, world), "global_batch": B * world, "parallelism": f"dp{world}",
                       "l2_policy": f"inputs ({B * a.samples * 4 / 1e6:.0f} MB PCM + activations per step) exceed the 126 MB L2; "
                                    f"{nbuf} alternating batches"},
            "e2e": e2e, "gpu_launches": launches, "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
            "gpu_library_baseline": gpu_lib, "drop_in_modules": modules, "last_loss": last_loss_dev,
            "groups_ms": {g["name"]: round(g["ms"], 4) for g in groups},
        }
        if a.breakdown:
            for g in sorted(groups, key=lambda g: -g["ms"]):
                print(f"{g['name']:28s} {g['ms']:8.4f} ms  x{g['launches_per_step']}", file=sys.stderr)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    args = parse()
    if args.impl == "reference":
        main_reference(args)
    else:
        main_ours(args)
