"""Manual GPU debugging aid (not a test): per-stage comparison of the res8 step against the oracle's intermediates."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import howl_b200
from oracle import howl_oracle as O

def al(x, a=256): return (x + a - 1) // a * a

def carve(B, H, L):
    off = 0; out = {}
    def take(name, nbytes):
        nonlocal off
        out[name] = off; off += al(nbytes)
    take("stats_fwd", 8 * 6 * 90); take("stats_bwd", 8 * 6 * 90); take("loss", 16); take("mean_rstd", 4 * 6 * 90)
    take("wT", 4 * 6 * 18225); take("pooled", 4 * B * 45); take("dh", 4 * B * 45); take("dlogits", 4 * B * L); take("logits", 4 * B * L)
    n = 4 * B * 45 * H * 10
    take("a0", n)
    for i in range(1, 7): take(f"u{i}", n)
    Kp = ((H + 2) * 11 + 15) // 16 * 16
    take("g", n); take("dc", max(n, B * 12 * Kp * 16)); take("gu0", n); take("gu1", n)
    return out, off

def main():
    dev = torch.device("cuda:0")
    ctx = howl_b200.Context(dev)
    B, T, L = int(os.environ.get("B", 4)), int(os.environ.get("T", 16000)), 12
    pcm, labels = O.synthetic_batch(B, T, L, seed=1)
    fb = O.mel_filterbank(40)
    zmean, zstd = -2.0166, 3.9955
    params, bn = O.res8_init(L, seed=2), O.res8_bn_init()
    flat = O.flatten(params, L).to(dev)
    bnd = torch.stack([torch.stack([bn[f"bn{i}.running_mean"], bn[f"bn{i}.running_var"]]) for i in range(1, 7)]).to(dev)
    nbt = torch.zeros(6, dtype=torch.int64, device=dev)
    feats = ctx.frontend(pcm.to(dev), fb.to(dev), "time_major", zmuv=(zmean, zstd))
    ofe = O.hot_path_features(pcm, fb, torch.tensor([zmean]), torch.tensor([zmean**2 + zstd**2]))
    print("feats err", (feats.cpu() - ofe[:, 0].transpose(1, 2)).abs().max().item())
    F = feats.shape[1]; H = F // 3
    nb = ctx.res8_workspace_bytes(B, F, L)
    offs, tot = carve(B, H, L)
    print("ws bytes", nb, "py carve", tot)
    ws = torch.zeros(nb, dtype=torch.uint8, device=dev)
    logits = ctx.res8_fwd(feats, flat, bnd, nbt, True, ws)
    torch.cuda.synchronize()
    taps = {}
    leaves = {k: p.clone().requires_grad_(True) for k, p in params.items()}
    ol = O.res8_forward(ofe, leaves, bn, True, taps)
    def view(name, shape):
        n = int(np.prod(shape))
        return ws[offs[name]:offs[name] + 4 * n].view(torch.float32).view(shape).cpu()
    for i in range(0, 7):
        name = "a0" if i == 0 else f"u{i}"
        got = view(name, (B, 45, H, 10)); want = taps[f"u{i}"].detach()
        print(name, "max err", (got - want).abs().max().item(), "scale", want.abs().max().item())
    print("pooled err", (view("pooled", (B, 45)) - taps["pooled"].detach()).abs().max().item())
    print("logits err", (logits.cpu() - ol.detach()).abs().max().item())
    loss = torch.nn.functional.cross_entropy(ol, labels); loss.backward()
    grads = torch.zeros_like(flat); gl = torch.zeros(1, device=dev)
    ctx.res8_bwd(feats, labels.to(dev), flat, grads, gl, ws)
    torch.cuda.synchronize()
    print("loss", gl.item(), loss.item())
    og = {k: leaves[k].grad for k in leaves}
    gg = O.unflatten(grads.cpu(), L)
    for k in gg:
        e = (gg[k] - og[k]).abs().max().item(); s = og[k].abs().max().item()
        print(f"grad {k}: max err {e:.3e} scale {s:.3e} rel {e / s:.3e}")

if __name__ == "__main__":
    main()
