import sys, os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch, numpy as np
import howl_b200
from howl_b200 import mobilenet as mb
from oracle import howl_oracle as O
from test_gpu_mobilenet import _flat_of, _bn_of
DEV = torch.device('cuda:0')
ctx = howl_b200.Context('cuda:0', n_mels=40)
g = dict(np.load('tests/golden/mobilenet_ckpt.npz'))
sd = {}
for k, v in g.items():
    if k.endswith("::bf16"):
        sd[k[:-6]] = torch.from_numpy(v.view(np.int16).copy()).view(torch.bfloat16).to(torch.float32)
    elif k.startswith(("downsample.", "model.")):
        sd[k] = torch.from_numpy(v)
L = 30
pcm, labels = torch.from_numpy(g["pcm"]), torch.from_numpy(g["labels"])
mean, mean2 = torch.from_numpy(g["zmuv_mean"]), torch.from_numpy(g["zmuv_mean2"])
fb = O.mel_filterbank(40)
feats = ctx.frontend(pcm.to(DEV), fb.to(DEV), "mels", zmuv=(float(mean[0]), float((mean2 - mean ** 2).sqrt()[0])))
flat = _flat_of(sd, L).to(DEV); bn = _bn_of(sd, L).to(DEV)
nbt = torch.zeros(mb.bn_layers(ctx), dtype=torch.int64, device=DEV)
ws = torch.empty(mb.workspace_bytes(ctx, pcm.shape[0], feats.shape[2], L), dtype=torch.uint8, device=DEV)
mb.forward(ctx, feats, flat, bn, nbt, True, ws)
grads, loss = torch.zeros_like(flat), torch.zeros(1, device=DEV)
mb.backward(ctx, feats, labels.to(DEV), flat, grads, loss, ws)
x = O.hot_path_features(pcm, fb, mean, mean2)
_, _, ograds = O.mobilenet_grads(x, labels, sd)
got, off = grads.cpu(), 0
rows = []
for name, shape in mb.param_shapes(L):
    n = int(np.prod(shape)); gg, w = got[off:off + n], ograds[name].reshape(-1); off += n
    rows.append((name, ((gg - w).norm() / (w.norm() + 1e-30)).item(), w.norm().item(), gg.norm().item()))
for r in rows[::-1]:
    print(f"{r[0]:44s} rel {r[1]:8.4f}  |ref| {r[2]:10.4e} |got| {r[3]:10.4e}")
