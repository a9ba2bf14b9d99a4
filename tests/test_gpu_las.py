"""GPU: LASClassifier forward (howl/model/rnn.py:133-215) through the nn.Module mirror, against the reference's own
logits for the shipped GSC checkpoint (tests/golden/las.npz, full and ragged lengths) and against the oracle on seeded weights / other
shapes, eval and train (batch statistics) mode.  Exact fp32: allclose(rtol, atol) = 1e-4."""
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
    model.train()
    with pytest.raises(NotImplementedError):
        model(feats, None)


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
