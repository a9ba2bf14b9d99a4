"""Manual GPU debugging aid (not a test): UMMA descriptor self-test + tensor-core vs fp32 engine comparison."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import howl_b200
from oracle import howl_oracle as O

dev = torch.device("cuda:0")
ctx = howl_b200.Context(dev)
torch.manual_seed(0)
sel = sys.argv[1] if len(sys.argv) > 1 else "all"
cases = {"k0": [(0, 0)], "mn0": [(1, 0)], "mn1": [(1, 1)], "ts": [(2, 0)], "all": [(0, 0), (1, 0)], "conv": []}[sel]
for mn, variant in cases:
    if True:
        if mn == 2:     # A [128][32] via tcgen05.cp into tensor memory, B [32][48] MN-major from shared memory
            A = torch.randn(128, 32, device=dev); Bm = torch.randn(32, 48, device=dev)
            ref = A.bfloat16().float() @ Bm.bfloat16().float()
        elif mn == 0:
            A = torch.randn(128, 32, device=dev); Bm = torch.randn(48, 32, device=dev)
            ref = A.bfloat16().float() @ Bm.bfloat16().float().t()
        else:
            A = torch.randn(32, 128, device=dev); Bm = torch.randn(32, 48, device=dev)
            ref = A.bfloat16().float().t() @ Bm.bfloat16().float()
        D = ctx.selftest_umma(A.contiguous(), Bm.contiguous(), int(mn), variant)
        torch.cuda.synchronize()
        print(f"selftest mn_major={mn} variant={variant}: max err {(D - ref).abs().max().item():.3e} (ref scale {ref.abs().max().item():.2f})", flush=True)

def run(engine, B, T, L):
    ctx.set_option("conv_engine", engine)
    pcm, labels = O.synthetic_batch(B, T, L, seed=1)
    params, bn = O.res8_init(L, seed=2), O.res8_bn_init()
    flat = O.flatten(params, L).to(dev)
    bnd = torch.stack([torch.stack([bn[f"bn{i}.running_mean"], bn[f"bn{i}.running_var"]]) for i in range(1, 7)]).to(dev)
    nbt = torch.zeros(6, dtype=torch.int64, device=dev)
    feats = ctx.frontend(pcm.to(dev), O.mel_filterbank(40).to(dev), "time_major", zmuv=(-2.0166, 3.9955))
    ws = torch.zeros(ctx.res8_workspace_bytes(B, feats.shape[1], L), dtype=torch.uint8, device=dev)
    logits = ctx.res8_fwd(feats, flat, bnd, nbt, True, ws)
    grads, loss = torch.zeros_like(flat), torch.zeros(1, device=dev)
    ctx.res8_bwd(feats, labels.to(dev), flat, grads, loss, ws)
    torch.cuda.synchronize()
    return logits.cpu(), grads.cpu(), loss.item(), bnd.cpu()

for (B, T, L) in ([(4, 16000, 12), (7, 8000, 4), (300, 16000, 12)] if sel in ("all", "conv") else []):
    l0, g0, s0, b0 = run(0, B, T, L)
    l1, g1, s1, b1 = run(1, B, T, L)
    print(f"B={B} T={T}: logits tc-vs-fp32 max {(l1 - l0).abs().max().item():.3e} (scale {l0.abs().max().item():.3f}) loss {s0:.6f} {s1:.6f} "
          f"bn max {(b1 - b0).abs().max().item():.3e}", flush=True)
    off = 0
    for name, shape in O.res8_param_shapes(L):
        n = 1
        for d in shape: n *= d
        a, b = g0[off:off + n], g1[off:off + n]; off += n
        print(f"   grad {name:14s} rel-l2 {((a - b).norm() / a.norm()).item():.3e} max/scale {((a - b).abs().max() / a.abs().max()).item():.3e}", flush=True)
