"""GPU: the MobileNetV2 path (BASELINE.json configs[2]) -- tensor-core GEMMs over the tile-major operand format, the
forward / backward of the whole network against the fp32 oracle (oracle/howl_oracle.py:mobilenet_forward, pinned to the shipped GSC
checkpoint by tests/test_oracle_golden.py), the nn.Module mirror and the fused train step.

Parity bar (bf16 storage of activations / gradients / GEMM weights, fp32 accumulation, BatchNorm statistics and master weights; SURVEY §7
hard part 9).  The oracle restates the forward in fp32 AND with the same bf16 storage points (mobilenet_forward(bf16=True)):
  * shipped GSC checkpoint (trained weights; tests/golden/mobilenet_ckpt.npz): logits rel-L2 <= 2e-2 against the REFERENCE module's logits in
    eval and in train (batch-statistics) mode with the same argmax, loss within 2 %, gradients rel-L2 <= 5e-2 per tensor against the oracle's
    autograd (pinned to the reference's gradients on the CPU);
  * random initialisation: eval mode only (rel-L2 <= 2e-2).  With batch-statistics BatchNorm a RANDOM MobileNetV2 is chaotic -- perturbations
    grow exponentially with depth; the oracle's own bf16 restatement differs from its fp32 one by ~20 % there -- so train-mode parity is judged
    on trained weights."""
import ctypes as C

import numpy as np
import pytest
import torch

from oracle import howl_oracle as O

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0")


@pytest.fixture(scope="module")
def ctx():
    import howl_b200

    c = howl_b200.Context("cuda:0", n_mels=40)
    yield c
    c.close()


def _vp(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def _bf16(x):
    return x.to(torch.bfloat16).to(torch.float32)


@pytest.mark.parametrize("M,K,N,with_add", [(128, 16, 16, False), (300, 32, 96, True), (1000, 144, 24, False), (517, 96, 576, True), (260, 960, 320, False),
                                            (64, 320, 1280, False), (4096, 27, 32, False)])
def test_tensor_core_gemm(ctx, M, K, N, with_add):
    g = torch.Generator().manual_seed(M + K + N)
    A, W = _bf16(torch.randn(M, K, generator=g)).to(DEV), _bf16(torch.randn(N, K, generator=g) / K ** 0.5).to(DEV)
    add = _bf16(torch.randn(M, N, generator=g)).to(DEV) if with_add else None
    out = torch.empty(M, N, device=DEV)
    ws = torch.empty(int(ctx.lib.howl_b200_debug_mbn_workspace_bytes(M, K, N)), dtype=torch.uint8, device=DEV)
    rc = ctx.lib.howl_b200_debug_mbn_gemm(ctx.handle, ctx._stream(), _vp(A), _vp(W), _vp(add), _vp(out), M, K, N, _vp(ws), ws.numel())
    assert rc == 0, ctx.lib.howl_b200_last_error(ctx.handle)
    want = A.double() @ W.double().t() + (add.double() if with_add else 0)
    np.testing.assert_allclose(out.cpu().numpy(), _bf16(want.float()).cpu().numpy(), rtol=1e-2, atol=1e-2)       # one bf16 rounding of the output
    assert (out.double() - want).norm() / want.norm() < 4e-3


@pytest.mark.parametrize("M,N,K", [(128, 16, 16), (1000, 24, 144), (5000, 96, 16), (777, 576, 96), (300, 320, 960), (260, 1280, 320), (4000, 32, 27)])
def test_tensor_core_weight_gradient(ctx, M, N, K):
    g = torch.Generator().manual_seed(M + K + N)
    dC, A = _bf16(torch.randn(M, N, generator=g)).to(DEV), _bf16(torch.randn(M, K, generator=g)).to(DEV)
    dW = torch.empty(N, K, device=DEV)
    ws = torch.empty(int(ctx.lib.howl_b200_debug_mbn_workspace_bytes(M, K, N)), dtype=torch.uint8, device=DEV)
    rc = ctx.lib.howl_b200_debug_mbn_wgrad(ctx.handle, ctx._stream(), _vp(dC), _vp(A), _vp(dW), M, N, K, _vp(ws), ws.numel())
    assert rc == 0, ctx.lib.howl_b200_last_error(ctx.handle)
    want = dC.double().t() @ A.double()
    assert (dW.double() - want).norm() / want.norm() < 1e-5          # exact bf16 products, fp32 accumulation


def _random_state(L, seed):
    """Seeded MobileNetClassifier state dict (oracle key names) with non-trivial BatchNorm parameters and running statistics."""
    from howl_b200 import mobilenet as mb

    g = torch.Generator().manual_seed(seed)
    flat = mb.init_flat(L, seed)
    sd, off = {}, 0
    for name, shape in mb.param_shapes(L):
        n = int(np.prod(shape))
        sd[name] = flat[off:off + n].view(shape).clone()
        off += n
    for _, _, bn, shape, _ in mb.layer_plan(L):
        c = shape[0]
        sd[bn + ".weight"] = 0.7 + 0.6 * torch.rand(c, generator=g)
        sd[bn + ".bias"] = 0.2 * torch.randn(c, generator=g)
        sd[bn + ".running_mean"] = 0.1 * torch.randn(c, generator=g)
        sd[bn + ".running_var"] = 0.5 + torch.rand(c, generator=g)
    sd["model.classifier.1.weight"] = torch.randn(L, 1280, generator=g) * 0.05
    sd["model.classifier.1.bias"] = torch.randn(L, generator=g) * 0.1
    return sd


def _flat_of(sd, L):
    from howl_b200 import mobilenet as mb

    return torch.cat([sd[name].reshape(-1) for name, _ in mb.param_shapes(L)])


def _bn_of(sd, L):
    from howl_b200 import mobilenet as mb

    bns = [bn for _, _, bn, _, _ in mb.layer_plan(L)]
    return torch.stack([torch.cat([sd[b + ".running_mean"] for b in bns]), torch.cat([sd[b + ".running_var"] for b in bns])]).contiguous()


def _checkpoint(golden):
    """The shipped GSC MobileNetV2 checkpoint (tests/golden/mobilenet_ckpt.npz, written by oracle/make_golden_mobilenet.py): GEMM weights as
    bf16 bit patterns, the rest fp32, with the REFERENCE module's logits / loss for exactly these weights."""
    g = golden("mobilenet_ckpt")
    sd = {}
    for k, v in g.items():
        if k.endswith("::bf16"):
            sd[k[:-6]] = torch.from_numpy(v.view(np.int16).copy()).view(torch.bfloat16).to(torch.float32)
        elif k.startswith(("downsample.", "model.")):
            sd[k] = torch.from_numpy(v)
    return g, sd


@pytest.mark.parametrize("train", [False, True])
def test_mobilenet_shipped_checkpoint_logits(ctx, golden, train):
    """Forward parity on TRAINED weights against the reference module's own logits (eval: running statistics; train: batch statistics,
    dropout off).  bf16 bar: rel-L2 <= 2e-2 and the same argmax (the oracle's bf16 restatement measures 0.7-1.3 % for this checkpoint)."""
    from howl_b200 import mobilenet as mb

    g, sd = _checkpoint(golden)
    L = 30
    pcm = torch.from_numpy(g["pcm"])
    mean = float(g["zmuv_mean"][0])
    std = float(np.sqrt(g["zmuv_mean2"][0] - g["zmuv_mean"][0] ** 2))
    feats = ctx.frontend(pcm.to(DEV), O.mel_filterbank(40).to(DEV), "mels", zmuv=(mean, std))
    bn = _bn_of(sd, L).to(DEV)
    nbt = torch.zeros(mb.bn_layers(ctx), dtype=torch.int64, device=DEV)
    ws = torch.empty(mb.workspace_bytes(ctx, pcm.shape[0], feats.shape[2], L), dtype=torch.uint8, device=DEV)
    logits = mb.forward(ctx, feats, _flat_of(sd, L).to(DEV), bn, nbt, train, ws).cpu()
    want = torch.from_numpy(g["logits_train" if train else "logits_eval"])
    rel = ((logits - want).norm() / want.norm()).item()
    print(f"mobilenet checkpoint logits train={train}: rel-L2 {rel:.4f}")
    assert rel < 2e-2, rel
    assert torch.equal(logits.argmax(1), want.argmax(1))


def test_mobilenet_shipped_checkpoint_gradients(ctx, golden):
    """Backward parity on trained weights: loss against the reference's; every parameter gradient against the autograd of the oracle's
    bf16-storage restatement with the ACTIVATION DECISIONS FORCED to the ones the GPU took (howl_b200_mobilenet_debug_masks; the fp32,
    unforced twin is pinned to the reference's gradients by tests/test_oracle_golden.py).  As for res8, the gradient of this network is
    discontinuous across ReLU / ReLU6 / max-pool boundaries, and -- unlike res8 -- it stays ILL-CONDITIONED even with the decisions forced:
    measured on the oracle alone (fp32, masks forced), a 1e-3 relative input perturbation moves the gradients by 7 % (median over tensors,
    20 % max), and its bf16 and fp32 restatements differ by 20 % / 36 % (53 BatchNorm layers with batch statistics amplify every bf16
    rounding of a stored activation).  No bf16 implementation can sit closer to another than that, so the bars are: the head's tensors,
    which see no amplification, tight (classifier and last BatchNorm <= 4e-2); every tensor <= 0.30 and the median <= 0.12 against the
    mask-forced bf16 restatement (measured 0.15 / 0.07); the unforced fp32 gradients as a loose envelope (<= 0.6)."""
    from howl_b200 import mobilenet as mb

    g, sd = _checkpoint(golden)
    L = 30
    pcm, labels = torch.from_numpy(g["pcm"]), torch.from_numpy(g["labels"])
    mean, mean2 = torch.from_numpy(g["zmuv_mean"]), torch.from_numpy(g["zmuv_mean2"])
    fb = O.mel_filterbank(40)
    feats = ctx.frontend(pcm.to(DEV), fb.to(DEV), "mels", zmuv=(float(mean[0]), float((mean2 - mean ** 2).sqrt()[0])))
    flat = _flat_of(sd, L).to(DEV)
    bn = _bn_of(sd, L).to(DEV)
    nbt = torch.zeros(mb.bn_layers(ctx), dtype=torch.int64, device=DEV)
    ws = torch.empty(mb.workspace_bytes(ctx, pcm.shape[0], feats.shape[2], L), dtype=torch.uint8, device=DEV)
    mb.forward(ctx, feats, flat, bn, nbt, True, ws)
    grads, loss = torch.zeros_like(flat), torch.zeros(1, device=DEV)
    mb.backward(ctx, feats, labels.to(DEV), flat, grads, loss, ws)
    assert abs(loss.item() - float(g["loss_train"])) < 2e-2 * float(g["loss_train"])
    x = O.hot_path_features(pcm, fb, mean, mean2)
    _, _, fgrads = O.mobilenet_grads(x, labels, sd)
    np.testing.assert_allclose([float(fgrads[k].norm()) for k in O.mobilenet_param_names(sd)], g["grad_norms"], rtol=1e-2,
                               atol=1e-4 * float(g["grad_norms"].max()))   # oracle (fp32) == reference
    # the activation decisions the GPU took (stem ReLU + max-pool routing, every ReLU6), NHWC bytes -> NCHW masks for the oracle
    B, n_mels, frames = pcm.shape[0], feats.shape[1], feats.shape[2]
    raw_masks = torch.empty(int(ctx.lib.howl_b200_mobilenet_debug_mask_bytes(B, frames, n_mels)), dtype=torch.uint8, device=DEV)
    rc = ctx.lib.howl_b200_mobilenet_debug_masks(ctx.handle, ctx._stream(), _vp(feats), _vp(flat), B, frames, n_mels, L, _vp(ws), ws.numel(), _vp(raw_masks))
    assert rc == 0, ctx.lib.howl_b200_last_error(ctx.handle)
    raw_masks = raw_masks.cpu()
    n0 = B * 3 * n_mels * (frames + 4)
    masks, off = [raw_masks[:n0].view(B, 3, n_mels, frames + 4)], n0
    h, w = (n_mels - 1) // 2 + 1, ((frames + 4) // 2 - 1) // 2 + 1
    shapes, inp = [(h, w, 32)], 32
    for t, c, n, st in mb.SETTING:
        for i in range(n):
            hidden, stride = inp * t, (st if i == 0 else 1)
            if t != 1:
                shapes.append((h, w, hidden))
            if stride == 2:
                h, w = (h - 1) // 2 + 1, (w - 1) // 2 + 1
            shapes.append((h, w, hidden))
            inp = c
    shapes.append((h, w, mb.LAST))
    for hh, ww, cc in shapes:
        n = B * hh * ww * cc
        masks.append(raw_masks[off:off + n].view(B, hh, ww, cc).permute(0, 3, 1, 2))
        off += n
    assert off == raw_masks.numel()
    _, _, ograds = O.mobilenet_grads(x, labels, sd, bf16=True, masks=masks)
    got, off, errs, ferrs = grads.cpu(), 0, {}, {}
    total = torch.cat([ograds[k].reshape(-1) for k in O.mobilenet_param_names(sd)]).norm()
    for name, shape in mb.param_shapes(L):
        n = int(np.prod(shape))
        gg, w, wf = got[off:off + n], ograds[name].reshape(-1), fgrads[name].reshape(-1)
        off += n
        if wf.norm() < 1e-5 * total:
            # analytically zero: a per-channel shift in front of a (1x1 conv +) batch-statistics BatchNorm -- the stem's conv bias and the
            # projection BatchNorm biases of blocks whose output only feeds such a layer.  Both sides hold rounding noise only.
            assert gg.norm() < 2e-3 * total, (name, gg.norm().item())
            continue
        errs[name] = ((gg - w).norm() / w.norm()).item()
        ferrs[name] = ((gg - wf).norm() / wf.norm()).item()
    worst = max(errs, key=errs.get)
    print(f"mobilenet checkpoint gradients vs bf16 restatement: median rel-L2 {np.median(list(errs.values())):.4f}, max {errs[worst]:.4f} ({worst}); "
          f"vs fp32: median {np.median(list(ferrs.values())):.4f}, max {max(ferrs.values()):.4f}")
    for name in ("model.classifier.1.weight", "model.classifier.1.bias", "model.features.18.1.weight", "model.features.18.1.bias"):
        assert errs[name] < 4e-2, (name, errs[name])
    assert errs[worst] < 0.30, (worst, errs[worst])
    assert np.median(list(errs.values())) < 0.12
    assert max(ferrs.values()) < 0.6


@pytest.mark.parametrize("B,T", [(4, 16000), (37, 8000), (3, 12345)])
def test_mobilenet_random_init_eval_forward(ctx, B, T):
    """Ragged sizes (row counts that are not multiples of the 128-row tile, other clip lengths) in eval mode against the fp32 oracle.
    (Train mode is not compared at random initialisation: a random MobileNetV2 with batch-statistics BatchNorm is chaotic -- the oracle's
    own bf16 restatement differs from its fp32 one by ~20 % there.)"""
    from howl_b200 import mobilenet as mb

    L = 12
    sd = _random_state(L, seed=B)
    pcm, _ = O.synthetic_batch(B, T, L, seed=B + 1)
    fb = O.mel_filterbank(40)
    zm = (-2.0166, 3.9955)
    feats = ctx.frontend(pcm.to(DEV), fb.to(DEV), "mels", zmuv=zm)
    bn = _bn_of(sd, L).to(DEV)
    nbt = torch.zeros(mb.bn_layers(ctx), dtype=torch.int64, device=DEV)
    ws = torch.empty(mb.workspace_bytes(ctx, B, feats.shape[2], L), dtype=torch.uint8, device=DEV)
    bn_before = bn.clone()
    logits = mb.forward(ctx, feats, _flat_of(sd, L).to(DEV), bn, nbt, False, ws).cpu()
    x = O.hot_path_features(pcm, fb, torch.tensor([zm[0]]), torch.tensor([zm[0] ** 2 + zm[1] ** 2]))
    with torch.no_grad():
        want = O.mobilenet_forward(x, sd, train=False)
    rel = ((logits - want).norm() / want.norm()).item()
    assert rel < 2e-2, rel
    assert torch.equal(bn, bn_before) and int(nbt.sum()) == 0
    # train mode on the same input: finite, running statistics of the stem's BatchNorm (fp32 end to end) updated as nn.BatchNorm2d does
    logits = mb.forward(ctx, feats, _flat_of(sd, L).to(DEV), bn, nbt, True, ws)
    assert torch.isfinite(logits).all() and nbt.tolist() == [1] * mb.bn_layers(ctx)
    stem = torch.nn.functional.conv2d(x[:, :1], sd["downsample.0.weight"], sd["downsample.0.bias"], padding=(1, 3))
    np.testing.assert_allclose(bn[0, :3].cpu().numpy(), (0.9 * sd["downsample.1.running_mean"] + 0.1 * stem.mean((0, 2, 3))).numpy(), rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(bn[1, :3].cpu().numpy(), (0.9 * sd["downsample.1.running_var"] + 0.1 * stem.var((0, 2, 3), unbiased=True)).numpy(),
                               rtol=1e-4, atol=1e-5)


def test_mobilenet_module_and_train_step():
    """Registry name, the reference's state_dict keys, autograd through the module with torch's CrossEntropyLoss + AdamW, and the fused
    train step learning a 4-class toy problem."""
    from howl_b200.model import RegisteredModel
    from howl_b200.trainer import MobileNetTrainStep

    cls = RegisteredModel.find_registered_class("mobilenet")
    model = cls(30)
    assert sum(p.numel() for p in model.parameters()) == 2262338 and len(model.state_dict()) == 321       # BASELINE.md / reference keys
    torch.manual_seed(0)
    model = cls(4).to(DEV).train()
    model.dropout_p = 0.0
    opt = torch.optim.AdamW(model.parameters(), 1e-3)
    x = torch.randn(16, 3, 40, 81, device=DEV)
    y = torch.randint(0, 4, (16,), device=DEV)
    losses = []
    for _ in range(8):
        loss = torch.nn.functional.cross_entropy(model(x, None), y)
        opt.zero_grad()
        loss.backward()
        opt.step()
        losses.append(loss.item())
    assert losses[-1] < 0.6 * losses[0], losses
    model.eval()
    with torch.no_grad():
        assert model(x, None).shape == (16, 4)
    step = MobileNetTrainStep(DEV, num_labels=4, batch=32, samples=16000, lr=1e-3, zmuv=(-2.0166, 3.9955), dropout_p=0.2)
    g = torch.Generator().manual_seed(1)
    t = torch.arange(16000) / 16000.0
    labels = torch.randint(0, 4, (32,), generator=g)
    pcm = torch.stack([0.2 * torch.sin(2 * np.pi * (300 + 400 * int(l)) * t) for l in labels]) + 0.02 * torch.randn(32, 16000, generator=g)
    first = None
    for _ in range(12):
        loss = step.step(pcm.to(DEV), labels.to(DEV)).item()
        first = loss if first is None else first
    assert np.isfinite(loss) and loss < 0.5 * first, (first, loss)


def test_mobilenet_full_batch_properties():
    """BASELINE config 3 size (B = 8192 x 1 s): size-independent properties of one forward + backward through the C ABI --
    (1) eval-mode logits of utterance i do not depend on the batch around it (the first 64 rows equal a 64-utterance call bit for bit),
    (2) permuting the batch permutes the train-mode logits and leaves loss / BatchNorm running statistics unchanged up to summation order,
    (3) every gradient is finite and the padded rows of the last 128-row tile did not leak into the statistics (counts = real rows)."""
    import howl_b200
    from howl_b200 import mobilenet as mb

    B, T, L = 8192, 16000, 12
    ctx = howl_b200.Context(DEV, n_mels=40)
    g = torch.Generator().manual_seed(0)
    pcm = (torch.randn(B, T, generator=g) * 0.1).clamp_(-1, 1).to(DEV)
    labels = torch.randint(0, L, (B,), generator=g).to(DEV)
    fb = O.mel_filterbank(40).to(DEV)
    feats = ctx.frontend(pcm, fb, "mels", zmuv=(-2.0166, 3.9955))
    del pcm
    flat = mb.init_flat(L, 3).to(DEV)
    nc = mb.bn_channels(ctx)

    def fresh():
        return torch.cat([torch.zeros(1, nc), torch.ones(1, nc)]).contiguous().to(DEV), torch.zeros(mb.bn_layers(ctx), dtype=torch.int64, device=DEV)

    ws = torch.empty(mb.workspace_bytes(ctx, B, feats.shape[2], L), dtype=torch.uint8, device=DEV)
    bn, nbt = fresh()
    ev_full = mb.forward(ctx, feats, flat, bn, nbt, False, ws).clone()
    ev_head = mb.forward(ctx, feats[:64].contiguous(), flat, bn, nbt, False, ws).clone()
    assert torch.equal(ev_full[:64], ev_head)
    bn1, nbt1 = fresh()
    lg1 = mb.forward(ctx, feats, flat, bn1, nbt1, True, ws).clone()
    g1, l1 = torch.empty_like(flat), torch.zeros(1, device=DEV)
    mb.backward(ctx, feats, labels, flat, g1, l1, ws)
    assert torch.isfinite(g1).all() and torch.isfinite(lg1).all() and nbt1.tolist() == [1] * mb.bn_layers(ctx)
    perm = torch.randperm(B, generator=g).to(DEV)
    feats_p = feats[perm].contiguous()
    bn2, nbt2 = fresh()
    lg2 = mb.forward(ctx, feats_p, flat, bn2, nbt2, True, ws).clone()
    l2 = torch.zeros(1, device=DEV)
    g2 = torch.empty_like(flat)
    mb.backward(ctx, feats_p, labels[perm].contiguous(), flat, g2, l2, ws)
    # batch statistics are fp64 sums of fp32 partials: order-dependent only in the last bits for the first layers (stem + entry
    # convolution = the first 35 channels); deeper layers see those bits amplified by the bf16 activations between them
    np.testing.assert_allclose(bn2[:, :35].cpu().numpy(), bn1[:, :35].cpu().numpy(), rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(bn2.cpu().numpy(), bn1.cpu().numpy(), rtol=0, atol=5e-3)
    assert abs(l1.item() - l2.item()) <= 2e-3 * abs(l1.item())
    # bf16 activations amplify those last bits through 52 layers: compare in the aggregate, not element-wise
    rel = ((lg2 - lg1[perm]).norm() / lg1.norm()).item()
    assert rel < 0.15, rel            # measured 0.052 on B200: a random-initialised MobileNetV2 in bf16 is chaotic in train mode
    cos = torch.nn.functional.cosine_similarity(g1, g2, dim=0).item()
    assert cos > 0.7, cos             # measured 0.87: the conditioning of this network's gradient, far from an indexing bug (~0)
