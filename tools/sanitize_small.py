"""Manual GPU aid: small-shape pass over every kernel family for compute-sanitizer (memcheck / racecheck / synccheck)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import howl_b200
from howl_b200.trainer import Res8TrainStep, LstmTrainStep, SeqLstmCtcTrainStep

dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
# (engine, batch, samples): 0.5 s clips (R = 192: three tiles, one straddling the two ring slots), 1 s clips on more CTAs than
# utterances, a 1000-sample clip (R = 64: one tile), 450 clips (3-4 utterances per CTA: ring, TMEM and dC-slot reuse), the fast mode and the fp32 engine
for engine, bsz, t in ((1, 5, 8000), (1, 3, 16000), (1, 2, 1000), (1, 450, 8000), (2, 5, 8000), (0, 5, 8000)):
    tr = Res8TrainStep(dev, num_labels=4, batch=bsz, samples=t, zmuv=(-1.8, 3.9))
    tr.ctx.set_option("conv_engine", engine)
    pcm = (torch.randn(bsz, t, generator=g) * 0.1).to(dev)
    lab = torch.randint(0, 4, (bsz,), generator=g).to(dev)
    for _ in range(2):
        tr.step(pcm, lab)
    rects = torch.tensor([[1, 3, 2, 5]] * bsz, dtype=torch.int32, device=dev)
    tr.ctx.frontend(pcm, tr.fb, "stacked", zmuv=(-1.8, 3.9), rects=rects)
    tr.ctx.frontend(pcm[:, : min(t, 4567)].contiguous(), tr.fb, "mels")
    torch.cuda.synchronize()
    print("res8 engine", engine, "loss", tr.loss.item(), flush=True)
lt = LstmTrainStep(dev, 5, 19, 8000, zmuv=(-1.8, 3.9))
p2 = (torch.randn(19, 8000, generator=g) * 0.1).to(dev)
for _ in range(2):
    lt.step(p2, torch.randint(0, 5, (19,), generator=g).to(dev))
st = SeqLstmCtcTrainStep(dev, 5, 19, 8000, blank=4, zmuv=(-1.8, 3.9))
for _ in range(2):
    st.step(p2, torch.randint(0, 4, (19, 3), generator=g).to(dev), torch.randint(1, 4, (19,), generator=g).to(dev))
torch.cuda.synchronize()
print("lstm", lt.loss.item(), "ctc", st.loss.item(), flush=True)
