"""Manual GPU aid (not a test, not the headline bench): lstm frame-objective train-step throughput, B=2048, 0.5 s clips
(the shape of BASELINE.json configs[3]; the CTC objective of that config is not built yet)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from howl_b200.trainer import LstmTrainStep, SeqLstmCtcTrainStep

B, T, L = int(os.environ.get("B", 2048)), 8000, 5
dev = torch.device("cuda:0")
tr = LstmTrainStep(dev, L, B, T, zmuv=(-1.78896, 3.93389))
g = torch.Generator().manual_seed(0)
pcm = (torch.randn(B, T, generator=g) * 0.1).clamp_(-1, 1).to(dev)
lab = torch.randint(0, L, (B,), generator=g).to(dev)
for _ in range(3): tr.step(pcm, lab)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): tr.step(pcm, lab)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
groups = tr.profile_groups(pcm, lab)
print(json.dumps({"model": "lstm (frame objective)", "batch": B, "ms_per_step": ms, "utt_per_s": B / ms * 1e3,
                  "groups_ms": {x["name"]: round(x["ms"], 4) for x in groups}}))

# ---- BASELINE.json configs[3]: seq-lstm, streaming state, CTC (blank = 4, L = 5), B = 2048, 0.5 s clips
tr = SeqLstmCtcTrainStep(dev, 5, B, T, blank=4, zmuv=(-1.78896, 3.93389), lr=1e-4)
tg = torch.randint(0, 4, (B, 3), generator=g).to(dev)
tl = torch.randint(1, 4, (B,), generator=g).to(dev)
for _ in range(3): tr.step(pcm, tg, tl)
torch.cuda.synchronize()
e0.record()
for _ in range(10): tr.step(pcm, tg, tl)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
groups = tr.profile_groups(pcm, tg, tl)
print(json.dumps({"model": "seq-lstm (streaming, CTC)", "batch": B, "ms_per_step": ms, "utt_per_s": B / ms * 1e3,
                  "loss": tr.loss.item(), "groups_ms": {x["name"]: round(x["ms"], 4) for x in groups}}))
