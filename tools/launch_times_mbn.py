"""Tuning aid: device time of every kernel launch of one MobileNetV2 train step (CUDA events on the launching stream), in launch order."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from howl_b200.trainer import MobileNetTrainStep

B, T, L = int(os.environ.get("B", 8192)), int(os.environ.get("T", 16000)), 12
dev = torch.device("cuda:0")
tr = MobileNetTrainStep(dev, L, B, T, zmuv=(-2.0166, 3.9955))
g = torch.Generator().manual_seed(0)
pcm = (torch.randn(B, T, generator=g) * 0.1).clamp_(-1, 1).to(dev)
lab = torch.randint(0, L, (B,), generator=g).to(dev)
for _ in range(3):
    tr.step(pcm, lab)
acc = None
N = 3
for _ in range(N):
    tr.ctx.profile_begin()
    tr.step(pcm, lab)
    rows = tr.ctx.profile_end()
    acc = rows if acc is None else [(n, a + b) for (n, a), (_, b) in zip(acc, rows)]
tot = 0.0
for n, ms in acc:
    print(f"{n:24s} {ms / N * 1e3:9.1f} us")
    tot += ms / N
print(f"total {tot:.3f} ms")
