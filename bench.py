#!/usr/bin/env python
"""bench.py -- res8 train-step throughput on N x B200 (utterances / s), with roofline and CPU baseline.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch B] [--seconds S]

A "step" is one pass of the hot path over one batch of synthetic 16 kHz clips:
frontend (STFT -> mel -> log -> ZMUV) -> Res8 forward -> CrossEntropy -> backward -> [allreduce] -> AdamW
(training/run/train.py:287-302 of the reference).  Workload at every N: BASELINE.json configs[1]
(res8, NUM_MELS=40, 1 s clips, batch 4096 per GPU; weak scaling, configs[4] at N=8).

  value  : whole-job utterances/s with PCM already resident in HBM (device-timed, max over ranks)
  e2e    : same metric through the public Python API with HOST (pinned) PCM/labels, H2D copies and a D2H read
           of the loss inside the timed region (double-buffered on a copy stream)
  --impl reference : the oracle port of the reference's torch CPU path timed on this box's host cores.
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

NUM_LABELS = 12          # GSC-12 semantics
SAMPLES = 16000          # 1 s @ 16 kHz
N_MELS = 40
LR, WD = 0.01, 1e-5
ZMEAN, ZSTD = -2.0166, 3.9955   # ZMUV constants of the reference's GSC res8 run
# algorithmic figures per utterance for the roofline (derivations in DESIGN.md)
STEP_FLOP_PER_UTT = 182.4e6
STEP_BYTES_PER_UTT = 64.7e3
CONV_LAYER_FLOP_PER_UTT = 2 * 45 * 45 * 9 * 270        # one 45->45 3x3 layer on 27x10 pixels (fwd or dgrad or wgrad)
FP32_PEAK_TFLOPS = 148 * 128 * 2 * 1.965e9 / 1e12      # FFMA roof: 148 SMs x 128 lanes x 2 flop x max clock


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=4096, help="utterances per GPU per step")
    ap.add_argument("--cpu-sample", type=int, default=256, help="utterances per CPU-baseline step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--breakdown", action="store_true", help="print per-kernel-group CUDA-event times to stderr")
    return ap.parse_args()


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"], "bf16_tflops_sustained": d.get("bf16_tflops_sustained"),
                "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clock / throttle sampling during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
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


# ---------------------------------------------------------------------------------------------------------------------
# reference arm / CPU baseline: the oracle port of the reference's torch-CPU path on the host cores
# ---------------------------------------------------------------------------------------------------------------------
def cpu_step_factory(batch):
    from oracle import howl_oracle as O  # the checker / CPU baseline -- never on the product path

    pcm, labels = O.synthetic_batch(batch, SAMPLES, NUM_LABELS, seed=0)
    fb = O.mel_filterbank(N_MELS)
    zm, zm2 = torch.tensor([ZMEAN]), torch.tensor([ZMEAN ** 2 + ZSTD ** 2])
    params, bn = O.res8_init(NUM_LABELS, seed=0), O.res8_bn_init()
    m = {k: torch.zeros_like(p) for k, p in params.items()}
    v = {k: torch.zeros_like(p) for k, p in params.items()}
    state = {"step": 0}

    def step():
        state["step"] += 1
        feats = O.hot_path_features(pcm, fb, zm, zm2)
        loss, _, _ = O.res8_train_step(feats, labels, params, bn, m, v, state["step"], LR, WD)
        return float(loss)

    return step


def run_cpu(batch, steps, warmup):
    """Times the oracle port on the host cores.  torch's intra-op pool is tuned first: on a 128-core box the small
    convolutions of res8 run slower with every core than with a few dozen, so the thread count that gives the best
    single-step time among {8, 16, 32, 64, all} is used and reported as `cores`."""
    ncpu = os.cpu_count() or 1
    step = cpu_step_factory(batch)
    best, best_t = ncpu, None
    for n in sorted({min(c, ncpu) for c in (8, 16, 32, 64, ncpu)}):
        torch.set_num_threads(n)
        step()
        t0 = time.perf_counter()
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


def main_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warmup = max(1, args.steps), max(1, min(args.warmup, 3))
    value, ms, cores = run_cpu(args.cpu_sample, steps, warmup)
    line = {
        "impl": "reference", "metric": "res8 train-step throughput (1 s @16 kHz clips)", "value": value, "unit": "utterances/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"res8 NUM_MELS=40 1 s clips L={NUM_LABELS}, oracle port of the reference torch-CPU path, "
                               f"{args.cpu_sample} utterances per step (bounded sample of the batch-4096 workload)"},
        "cpu_baseline": {"value": value, "unit": "utterances/s", "cores": cores, "kind": "port",
                         "sample": f"{steps} steps x {args.cpu_sample} utterances, torch CPU {torch.__version__}"},
        "e2e": {"value": value, "unit": "utterances/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------------------------
def main_ours(args):
    import torch.distributed as dist

    import howl_b200
    from howl_b200.trainer import Res8TrainStep

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")   # keep stdout to the one JSON line
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    B = args.batch
    gen = torch.Generator(device="cpu").manual_seed(1234 + rank)
    # synthetic inputs: noise at speech-like RMS, uniform labels; one distinct batch per in-flight slot
    nbuf = 2
    host_pcm = [(torch.randn(B, SAMPLES, generator=gen) * 0.1).clamp_(-1, 1).pin_memory() for _ in range(nbuf)]
    host_lab = [torch.randint(0, NUM_LABELS, (B,), generator=gen).pin_memory() for _ in range(nbuf)]
    dev_pcm = [h.to(dev) for h in host_pcm]
    dev_lab = [h.to(dev) for h in host_lab]

    step_obj = Res8TrainStep(dev, num_labels=NUM_LABELS, batch=B, samples=SAMPLES, n_mels=N_MELS, lr=LR, weight_decay=WD,
                             zmuv=(ZMEAN, ZSTD), seed=0, world_size=world)
    for opt in ("conv_engine",):   # tuning experiments only
        if os.environ.get("HOWL_" + opt.upper()):
            step_obj.ctx.set_option(opt, int(os.environ["HOWL_" + opt.upper()]))
    ctx = step_obj.ctx

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---- device-resident timing (value)
    for i in range(max(args.warmup, 3)):
        step_obj.step(dev_pcm[i % nbuf], dev_lab[i % nbuf])
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    launches0 = ctx.launch_count
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for i in range(args.steps):
        step_obj.step(dev_pcm[i % nbuf], dev_lab[i % nbuf])
    ev1.record()
    barrier()
    launches = ctx.launch_count - launches0
    ms_total = ev0.elapsed_time(ev1)
    clocks = sampler.stop() if sampler else None
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    ms_step = ms_total / args.steps
    value = B * world * args.steps / (ms_total / 1e3)

    # ---- per-kernel-group device times for the roofline (live, CUDA events on the launching stream)
    groups = step_obj.profile_groups(dev_pcm[0], dev_lab[0], reps=3)
    barrier()

    # ---- end to end through the public API with host buffers
    e2e = None
    if not args.no_e2e:
        for i in range(2):
            step_obj.step_host(host_pcm[i % nbuf], host_lab[i % nbuf])
        step_obj.flush_host()
        barrier()
        t0 = time.perf_counter()
        ev0.record()
        for i in range(args.steps):
            step_obj.step_host(host_pcm[i % nbuf], host_lab[i % nbuf])
        last_loss = step_obj.flush_host()
        ev1.record()
        barrier()
        e_ms = ev0.elapsed_time(ev1)
        t = torch.tensor([e_ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e = {"value": B * world * args.steps / (float(t.item()) / 1e3), "unit": "utterances/s",
               "h2d_bytes_per_step": B * SAMPLES * 4 + B * 8, "d2h_bytes_per_step": 4, "wall_s": time.perf_counter() - t0,
               "last_loss": last_loss}

    if rank == 0:
        peaks = load_peaks()
        by = {g["name"]: g for g in groups}
        conv_launches = 18   # 6 fwd + 6 dgrad + 6 wgrad launches of the 45->45 3x3 kernels per step
        conv_ms = sum(g["ms"] for g in groups if g["name"].startswith("conv3x3"))
        conv_tf = B * CONV_LAYER_FLOP_PER_UTT * conv_launches / (conv_ms / 1e3) / 1e12
        tensor_peak = peaks["bf16_tflops_sustained"] or peaks["bf16_tflops"]
        tj = {}
        tpath = os.path.join(ROOT, "profiles", "r01_traffic.json")
        if os.path.exists(tpath):   # dram__bytes_read + write per launch, from the committed `ncu --set full` capture
            tj = json.load(open(tpath))
        engine = "tcgen05 bf16x3-split MMA, fp32 TMEM accumulate" if any("_tc" in g["name"] for g in groups) else "fp32 FFMA"
        # the dominant kernel group of the step decides the headline bound; the other groups are listed beside it
        dom = max((g for g in groups if g["name"].startswith(("conv3x3", "bn_bwd_apply"))), key=lambda g: g["ms"])
        HW_ = 27 * 10
        plane, opb = 45 * HW_ * 4, 12 * 320 * 16          # fp32 activation plane set / operand-format block, bytes per utterance
        # BatchNorm-backward kernel, algorithmic bytes per utterance summed over its six launches:
        #   layer 6: u + mask in, gu + dc_op + dc_opT out; layers 4, 2: g + u + mask + gu in, same out; layers 5, 3, 1: g + u in, 2 operands out
        apply_bytes = (2 * plane + plane + 2 * opb) + 2 * (4 * plane + plane + 2 * opb) + 3 * (2 * plane + 2 * opb)
        if dom["name"].startswith("bn_bwd_apply"):
            launches_dom = max(1, dom.get("launches_per_step", 6))
            ach = B * apply_bytes / (dom["ms"] / 1e3) / 1e9
            roofline = {"bound": "hbm", "kernel": "bn_bwd_apply_op (BatchNorm backward + residual fan-in + ReLU mask -> both gradient "
                                                  "operand formats), 6 launches/step",
                        "achieved": ach, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": ach / peaks["hbm_gbs"],
                        "traffic": tj.get("bn_bwd_apply_op"), "algorithmic_bytes_per_launch": B * apply_bytes / launches_dom,
                        "peak_source": peaks["source"] + " (copy bandwidth)"}
        else:
            per_launch_ms = dom["ms"] / 6
            ach = B * CONV_LAYER_FLOP_PER_UTT / (per_launch_ms / 1e3) / 1e12
            roofline = {"bound": "tensor", "kernel": f"{dom['name']} (conv3x3 45->45, 6 launches/step), {engine}", "achieved": ach,
                        "peak": tensor_peak, "unit": "TFLOP/s", "frac": ach / tensor_peak, "traffic": tj.get(dom["name"]),
                        "peak_source": peaks["source"] + " (dense bf16 cuBLAS, sustained)"}
        roofline.update({
            "dominant_launch": dom["name"], "dominant_launch_ms": dom["ms"],
            "conv": {"kernel": f"conv3x3 45->45, 18 launches/step (6 fwd + 6 dgrad + 6 wgrad), {engine}", "tflops": conv_tf,
                     "frac_of_bf16_peak": conv_tf / tensor_peak, "fp32_ffma_peak_tflops": FP32_PEAK_TFLOPS,
                     "frac_of_fp32_ffma_peak": conv_tf / FP32_PEAK_TFLOPS, "ms_per_step": conv_ms,
                     "note": "issued bf16 MMA flops are 3x the fp32-equivalent figure (hi*hi + hi*lo + lo*hi); the N=48 MMAs are bound "
                             "by shared-memory operand reads, see DESIGN.md"},
            "apply": ({"hbm_gbs": B * apply_bytes / (by["bn_bwd_apply_op"]["ms"] / 1e3) / 1e9,
                       "hbm_frac": B * apply_bytes / (by["bn_bwd_apply_op"]["ms"] / 1e3) / 1e9 / peaks["hbm_gbs"]}
                      if "bn_bwd_apply_op" in by else None),
            "step_hbm_gbs_algorithmic": B * STEP_BYTES_PER_UTT / (ms_step / 1e3) / 1e9,
            "step_hbm_frac": B * STEP_BYTES_PER_UTT / (ms_step / 1e3) / 1e9 / peaks["hbm_gbs"],
            "step_tflops_algorithmic": B * STEP_FLOP_PER_UTT / (ms_step / 1e3) / 1e12,
            "frontend": next((g for g in groups if g["name"] == "frontend"), None),
        })
        fe = roofline["frontend"]
        if fe:
            fe["hbm_gbs"] = B * 76960 / (fe["ms"] / 1e3) / 1e9
            fe["hbm_frac"] = fe["hbm_gbs"] / peaks["hbm_gbs"]
        cpu = None
        if not args.no_cpu_baseline and world == 1:
            v, ms, cores = run_cpu(args.cpu_sample, 3, 1)
            cpu = {"value": v, "unit": "utterances/s", "cores": cores, "kind": "port",
                   "sample": f"3 steps x {args.cpu_sample} utterances of the same workload (oracle port, torch CPU)"}
        line = {
            "metric": "res8 train-step throughput (1 s @16 kHz clips)", "value": value, "unit": "utterances/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"res8 NUM_MELS=40 batch={B}/GPU fused STFT->mel->conv train step, synthetic GSC-shaped 1 s clips, "
                                   f"L={NUM_LABELS}", "global_batch": B * world, "parallelism": f"dp{world}",
                       "l2_policy": "inputs (262 MB PCM + 4.5 GB activations and operands per step) exceed the 126 MB L2; two alternating batches"},
            "e2e": e2e, "gpu_launches": launches, "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
            "groups_ms": {g["name"]: round(g["ms"], 4) for g in groups},
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        main_reference(a)
    else:
        main_ours(a)
