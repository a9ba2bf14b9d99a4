"""GPU: LASClassifier (howl/model/rnn.py:133-215) through the nn.Module mirror and the C ABI: forward against the
reference's own logits for the shipped GSC checkpoint (tests/golden/las.npz, full and ragged lengths) and against the oracle on seeded
weights / other shapes, eval and train (batch statistics) mode, exact fp32: allclose(rtol, atol) = 1e-4; backward (CrossEntropy ->
every parameter gradient) against fp64 autograd over the oracle, per tensor max|diff| <= 2e-3 max|ref|; dropout through the mask the
gradients reveal; a few AdamW steps against the oracle loop."""
import numpy as np
import pytest
import torch

from oracle import howl_oracle as O

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0")
RTOL = ATOL = 1e-4


@pytest.fixture(autouse=True)
def _env(monkeypatch):
    from howl_b200.settings import SETTINGS

    monkeypatch.setenv("NUM_MELS", "40")
    SETTINGS.reset()
    yield
    SETTINGS.reset()


def test_las_shipped_checkpoint_full_and_ragged(golden):
    from howl_b200.model import RegisteredModel

    g = golden("las")
    model = RegisteredModel.find_registered_class("las")(30)
    keys = [k[3:] for k in g if k.startswith("sd.")]
    assert list(model.state_dict().keys()) == keys and sum(p.numel() for p in model.parameters()) == 477862
    model.load_state_dict({k: torch.from_numpy(g["sd." + k]) for k in keys})
    model = model.to(DEV).eval()
    feats = torch.from_numpy(g["feats"]).to(DEV)
    with torch.no_grad():
        full = model(feats, None).cpu().numpy()
        ragged = model(feats, torch.from_numpy(g["lengths"])).cpu().numpy()
    np.testing.assert_allclose(full, g["logits_full"], rtol=RTOL, atol=ATOL)
    np.testing.assert_allclose(ragged, g["logits_ragged"], rtol=RTOL, atol=ATOL)
    assert np.array_equal(full.argmax(1), g["logits_full"].argmax(1))


@pytest.mark.parametrize("B,T,train", [(3, 8000, False), (37, 16000, False), (20, 16000, True), (5, 12345, True)])
def test_las_forward_vs_oracle(B, T, train):
    import howl_b200
    from howl_b200 import las

    L = 12
    ctx = howl_b200.Context(DEV, n_mels=40)
    g = torch.Generator().manual_seed(B)
    sd = {}
    for name, shape in las.param_shapes(L):
        sd[name] = torch.randn(shape, generator=g) * (0.5 / np.sqrt(max(int(np.prod(shape[1:])), 1)) if len(shape) > 1 else 0.1)
    for idx in ("1", "5"):
        p = f"encoder.conv_encoder.{idx}"
        sd[p + ".weight"] = 0.7 + 0.6 * torch.rand(8, generator=g)
        sd[p + ".running_mean"], sd[p + ".running_var"] = 0.1 * torch.randn(8, generator=g), 0.5 + torch.rand(8, generator=g)
    pcm, _ = O.synthetic_batch(B, T, L, seed=B + 3)
    fb = O.mel_filterbank(40)
    zm = (-2.0166, 3.9955)
    F = O.num_frames(T)
    rng = np.random.default_rng(B)
    lengths = torch.from_numpy(np.sort(rng.integers(F // 3, F + 1, size=B))[::-1].copy())
    lengths[0] = F
    feats = ctx.frontend(pcm.to(DEV), fb.to(DEV), "stacked", zmuv=zm)
    flat = torch.cat([sd[n].reshape(-1) for n, _ in las.param_shapes(L)]).to(DEV)
    bn = torch.stack([torch.stack([sd[f"encoder.conv_encoder.{i}.running_mean"], sd[f"encoder.conv_encoder.{i}.running_var"]]) for i in ("1", "5")]).to(DEV)
    nbt = torch.zeros(2, dtype=torch.int64, device=DEV)
    bn_before = bn.clone()
    logits = las.forward(ctx, feats, lengths, flat, bn, nbt, train, num_labels=L).cpu().numpy()
    x = O.hot_path_features(pcm, fb, torch.tensor([zm[0]]), torch.tensor([zm[0] ** 2 + zm[1] ** 2]))
    with torch.no_grad():
        want = O.las_forward(x, sd, lengths, train=train).numpy()
    np.testing.assert_allclose(logits, want, rtol=RTOL, atol=ATOL)
    if train:
        assert nbt.tolist() == [1, 1] and not torch.equal(bn, bn_before)
        c1 = torch.nn.functional.conv2d(x, sd["encoder.conv1.weight"], sd["encoder.conv1.bias"], padding=2)
        np.testing.assert_allclose(bn[0, 0].cpu().numpy(), (0.9 * sd["encoder.conv_encoder.1.running_mean"] + 0.1 * c1.mean((0, 2, 3))).numpy(), rtol=1e-4, atol=1e-5)
        np.testing.assert_allclose(bn[0, 1].cpu().numpy(), (0.9 * sd["encoder.conv_encoder.1.running_var"] + 0.1 * c1.var((0, 2, 3), unbiased=True)).numpy(),
                                   rtol=1e-4, atol=1e-5)
    else:
        assert torch.equal(bn, bn_before) and int(nbt.sum()) == 0
    ctx.close()


def _seeded(B, T, L, seed):
    import howl_b200
    from howl_b200 import las

    ctx = howl_b200.Context(DEV, n_mels=40)
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for name, shape in las.param_shapes(L):
        sd[name] = torch.randn(shape, generator=g) * (0.7 / np.sqrt(max(int(np.prod(shape[1:])), 1)) if len(shape) > 1 else 0.1)
    for idx in ("1", "5"):
        p = f"encoder.conv_encoder.{idx}"
        sd[p + ".weight"] = 0.7 + 0.6 * torch.rand(8, generator=g)
        sd[p + ".running_mean"], sd[p + ".running_var"] = torch.zeros(8), torch.ones(8)
    pcm, labels = O.synthetic_batch(B, T, L, seed=seed + 3)
    fb = O.mel_filterbank(40)
    zm = (-2.0166, 3.9955)
    F = O.num_frames(T)
    rng = np.random.default_rng(seed)
    lengths = torch.from_numpy(np.sort(rng.integers(F // 3, F + 1, size=B))[::-1].copy())
    lengths[0] = F
    feats = ctx.frontend(pcm.to(DEV), fb.to(DEV), "stacked", zmuv=zm)
    x = O.hot_path_features(pcm, fb, torch.tensor([zm[0]]), torch.tensor([zm[0] ** 2 + zm[1] ** 2]))
    return ctx, sd, feats, x, labels, lengths


def _run_gpu(ctx, sd, feats, labels, lengths, L, p=0.0, seed=0):
    from howl_b200 import las

    B, _, M, Fr = feats.shape
    flat = torch.cat([sd[n].reshape(-1) for n, _ in las.param_shapes(L)]).to(DEV)
    bn = torch.stack([torch.stack([sd[f"encoder.conv_encoder.{i}.running_mean"], sd[f"encoder.conv_encoder.{i}.running_var"]]) for i in ("1", "5")]).to(DEV)
    nbt = torch.zeros(2, dtype=torch.int64, device=DEV)
    ws = torch.empty(las.workspace_bytes(ctx, B, Fr, M, L, True), dtype=torch.uint8, device=DEV)
    enc = las.encoder_lengths(ctx, lengths, B, Fr)
    logits = las.forward(ctx, feats, None, flat, bn, nbt, True, ws, L, p, seed, enc)
    grads = torch.full_like(flat, float("nan"))
    loss = torch.zeros(1, device=DEV)
    las.backward(ctx, feats, enc, flat, grads, ws, L, labels=labels.to(DEV), loss=loss, dropout_p=p)
    out, off = {}, 0
    for n, shape in las.param_shapes(L):
        k = int(np.prod(shape))
        out[n] = grads[off:off + k].view(shape).cpu()
        off += k
    return logits.cpu(), float(loss), out


def _compare(got, want, bar=2e-3):
    worst = {}
    for k, ref in want.items():
        ref = ref.float()
        err = float((got[k] - ref).abs().max())
        scale = float(ref.abs().max())
        worst[k] = err / max(scale, 1e-12)
        assert torch.isfinite(got[k]).all(), k
        assert err <= bar * scale + 1e-7, f"{k}: max|diff| {err:.3e} vs max|ref| {scale:.3e}"
    return worst


@pytest.mark.parametrize("B,T", [(1, 16000), (6, 8000), (19, 16000), (40, 12345), (72, 16000)])     # 72 x 21 steps >= 1024 rows: tensor-core weight gradients
def test_las_gradients_vs_oracle(B, T):
    L = 12
    ctx, sd, feats, x, labels, lengths = _seeded(B, T, L, seed=100 + B)
    logits, loss, grads = _run_gpu(ctx, sd, feats, labels, lengths, L)
    want_loss, want_logits, want = O.las_grads(x, labels, sd, lengths)
    np.testing.assert_allclose(logits.numpy(), want_logits.float().numpy(), rtol=RTOL, atol=ATOL)
    assert abs(loss - want_loss) <= 1e-5 * max(1.0, abs(want_loss))
    _compare(grads, want)


def test_las_dropout_mask_consistent():
    """Dropout uses the library's own counter-based mask: the same seed repeats, another seed differs, and an oracle given the mask that the
    fc.0 bias gradient reveals (kept and active units) reproduces logits and every gradient."""
    L, B, p = 12, 9, 0.5
    ctx, sd, feats, x, labels, lengths = _seeded(B, 16000, L, seed=7)
    a = _run_gpu(ctx, sd, feats, labels, lengths, L, p, seed=11)
    b = _run_gpu(ctx, sd, feats, labels, lengths, L, p, seed=11)
    c = _run_gpu(ctx, sd, feats, labels, lengths, L, p, seed=12)
    assert torch.equal(a[0], b[0]) and not torch.equal(a[0], c[0])
    base = _run_gpu(ctx, sd, feats, labels, lengths, L, 0.0)
    assert not torch.allclose(a[0], base[0])
    # per-utterance mask: a unit of utterance b was kept iff dropping it changes ... recover it from single-utterance runs
    masks = []
    for i in range(B):
        one = _run_gpu_single(ctx, sd, feats, labels, lengths, L, p, 11, i)
        masks.append(one)
    mask = torch.stack(masks)
    frac = float((mask > 0).float().mean())
    assert 0.35 < frac < 0.65
    want_loss, want_logits, want = O.las_grads(x, labels, sd, lengths, hid_mask=mask)
    np.testing.assert_allclose(a[0].numpy(), want_logits.float().numpy(), rtol=RTOL, atol=ATOL)
    _compare(a[2], want)


def _run_gpu_single(ctx, sd, feats, labels, lengths, L, p, seed, i):
    """Keep factors [256] of utterance i: with fc.3 = identity rows the logits ARE the dropped hidden units, so logits / undropped logits
    gives keep / (1 - p) wherever the unit is active."""
    from howl_b200 import las

    B, _, M, Fr = feats.shape
    sd2 = dict(sd)
    Lh = 256
    sd2["fc.3.weight"], sd2["fc.3.bias"] = torch.eye(256), torch.zeros(256)
    sd2["fc.0.bias"] = sd["fc.0.bias"] + 50.0          # every unit active
    shapes = las.param_shapes(Lh)
    flat = torch.cat([sd2[n].reshape(-1) for n, _ in shapes]).to(DEV)
    bn = torch.stack([torch.stack([sd[f"encoder.conv_encoder.{k}.running_mean"], sd[f"encoder.conv_encoder.{k}.running_var"]]) for k in ("1", "5")]).to(DEV)
    nbt = torch.zeros(2, dtype=torch.int64, device=DEV)
    ws = torch.empty(las.workspace_bytes(ctx, B, Fr, M, Lh, True), dtype=torch.uint8, device=DEV)
    enc = las.encoder_lengths(ctx, lengths, B, Fr)
    dropped = las.forward(ctx, feats, None, flat, bn.clone(), nbt, True, ws, Lh, p, seed, enc)[i]
    full = las.forward(ctx, feats, None, flat, bn.clone(), nbt, True, ws, Lh, 0.0, seed, enc)[i]
    ratio = (dropped / full).cpu()
    keep = ratio > 0.5
    assert torch.allclose(ratio[keep], torch.full_like(ratio[keep], 1.0 / (1.0 - p)), rtol=1e-5)
    assert float(ratio[~keep].abs().max()) == 0.0 if (~keep).any() else True
    return keep.float() / (1.0 - p)


def test_las_module_trains_like_oracle():
    """nn.Module under autograd + torch.optim.AdamW (train.py:255-256,299-302): three steps track the fp64 oracle loop."""
    from howl_b200.model import RegisteredModel

    L, B = 12, 10
    ctx, sd, feats, x, labels, lengths = _seeded(B, 16000, L, seed=21)
    torch.manual_seed(5)
    model = RegisteredModel.find_registered_class("las")(L)
    model.dropout_p = 0.0
    state = model.state_dict()
    for k in state:
        if k in sd:
            state[k] = sd[k].clone()
    state["encoder.conv_encoder.0.weight"], state["encoder.conv_encoder.0.bias"] = sd["encoder.conv1.weight"], sd["encoder.conv1.bias"]
    state["encoder.conv_encoder.4.weight"], state["encoder.conv_encoder.4.bias"] = sd["encoder.conv2.weight"], sd["encoder.conv2.bias"]
    model.load_state_dict(state)
    model = model.to(DEV).train()
    opt = torch.optim.AdamW(model.parameters(), 1e-3, weight_decay=1e-2)
    names = [n for n, _ in __import__("howl_b200").las.param_shapes(L)]
    ref = {k: v.double().clone() for k, v in sd.items()}
    ref_params = [ref[n].requires_grad_(True) for n in names]
    ref_opt = torch.optim.AdamW(ref_params, 1e-3, weight_decay=1e-2)
    crit = torch.nn.CrossEntropyLoss()
    for step in range(3):
        opt.zero_grad()
        loss = crit(model(feats, lengths), labels.to(DEV))
        loss.backward()
        opt.step()
        ref_opt.zero_grad()
        ref_loss = crit(O.las_forward(x.double(), ref, lengths, train=True), labels)
        ref_loss.backward()
        ref_opt.step()
        assert abs(float(loss.detach()) - float(ref_loss.detach())) <= 2e-4 * max(1.0, abs(float(ref_loss))), step
    got = dict(model.named_parameters())
    for n in names:
        # Three tensors have an exactly zero gradient: v_proj.bias shifts every attention logit of a head equally (softmax shift
        # invariance) and the convolution biases are removed by the batch-statistics BatchNorm that follows.  AdamW's normalised update
        # only sees rounding noise there (fp32 here, fp64 in the oracle) -- bounded by 3 steps x lr, not by parity
        atol = 3.1e-3 if n in ("attn.v_proj.bias", "encoder.conv1.bias", "encoder.conv2.bias") else 3e-4
        np.testing.assert_allclose(got[n].detach().cpu().numpy(), ref[n].detach().float().numpy(), rtol=0, atol=atol)
    bn = model.state_dict()["encoder.conv_encoder.1.num_batches_tracked"]
    assert int(bn) == 3


def test_las_full_batch_permutation():
    """Bench size (B = 2048 x 1 s, ragged lengths): permuting the batch permutes the logits and leaves loss and every parameter gradient
    unchanged (exact fp32 arithmetic; only the order of the batch sums differs)."""
    from howl_b200 import las

    L, B = 12, 2048
    ctx, sd, feats, x, labels, lengths = _seeded(B, 16000, L, seed=5)
    del x
    lg1, loss1, g1 = _run_gpu(ctx, sd, feats, labels, lengths, L)
    perm = torch.randperm(B, generator=torch.Generator().manual_seed(1))
    lg2, loss2, g2 = _run_gpu(ctx, sd, feats[perm.to(DEV)].contiguous(), labels[perm], lengths[perm], L)
    np.testing.assert_allclose(lg2.numpy(), lg1[perm].numpy(), rtol=1e-4, atol=1e-4)
    assert abs(loss1 - loss2) <= 1e-5 * abs(loss1)
    for k in g1:
        scale = float(g1[k].abs().max())
        assert float((g1[k] - g2[k]).abs().max()) <= 2e-3 * scale + 1e-7, k
