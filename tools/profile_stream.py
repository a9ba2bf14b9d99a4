"""Manual GPU tuning aid (not a test): where do the issuer / epilogue of the tensor-core stream conv kernels wait?

Prints, per kind (forward / data gradient), the mean over CTAs of the cycle counters written by the LAST launch of that kind
in one training step (layer 6 forward, layer 1 data gradient): issuer waits on [operand slot 0 ready, slot-1 release,
slot 1 ready, accumulator free, slot-0 release], issuer loop total, epilogue warp 0 wait-for-tile and loop total.
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from howl_b200.trainer import Res8TrainStep

dev = torch.device("cuda:0")
B, T, L = 4096, 16000, 12
step = Res8TrainStep(dev, num_labels=L, batch=B, samples=T, zmuv=(-2.0, 4.0))
pcm = (torch.randn(B, T) * 3000).round().to(dev)     # float32 PCM in int16 range
lab = torch.randint(0, L, (B,), device=dev)
for _ in range(3):
    step.step(pcm, lab)
names = ["wait ready0", "wait ready1", "-", "wait tmem free", "-", "issuer total", "sum load->ready (exposed only)", "n exposed",
         "epi wait tile", "epi total"]
for kind, label in ((1, "forward"), (2, "data gradient")):
    buf = torch.zeros(step.ctx.sm_count * 16, dtype=torch.int64, device=dev)
    step.ctx.debug_stream_profile(buf, kind)
    step.step(pcm, lab)
    torch.cuda.synchronize()
    step.ctx.debug_stream_profile(None, 0)
    m = buf.view(-1, 16).double().mean(0).cpu()
    print(label, {n: int(m[i]) for i, n in enumerate(names) if n != "-"}, flush=True)
