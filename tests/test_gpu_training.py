"""GPU: the fused train step LEARNS -- accuracy on held-out synthetic labels matches the oracle trained the same way.

north_star: "matching reference accuracy on held-out synthetic labels".  Data per SURVEY §8(d): class-conditional
formant pairs + noise, split by sha256(index) % 100 (mirrors howl/utils/hash_utils.py:20-40).
"""
import hashlib
import math

import numpy as np
import pytest
import torch

from oracle import howl_oracle as O

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0")
L, T, N_TRAIN_STEPS, BATCH = 4, 8000, 60, 64


def _clip(index: int, label: int, rng: np.random.Generator) -> np.ndarray:
    t = np.arange(T) / 16000.0
    f1, f2 = [(300, 2300), (500, 1500), (700, 1100), (400, 3000)][label]
    jitter = 1.0 + 0.03 * rng.standard_normal()
    x = 0.12 * np.sin(2 * math.pi * f1 * jitter * t + rng.uniform(0, 6.28)) + 0.08 * np.sin(2 * math.pi * f2 * jitter * t)
    return np.clip(x + 0.05 * rng.standard_normal(T), -1, 1).astype(np.float32)


def _bucket(index: int) -> str:
    h = int(hashlib.sha256(str(index).encode()).hexdigest(), 16) % 100
    return "train" if h < 80 else ("dev" if h < 90 else "test")


def _dataset(n: int):
    rng = np.random.default_rng(7)
    items = [(i, i % L) for i in range(n)]
    split = {"train": [], "dev": [], "test": []}
    for i, y in items:
        split[_bucket(i)].append((_clip(i, y, rng), y))
    return split


def _batches(items, steps):
    rng = np.random.default_rng(3)
    for _ in range(steps):
        idx = rng.choice(len(items), BATCH, replace=False)
        yield (torch.from_numpy(np.stack([items[i][0] for i in idx])), torch.tensor([items[i][1] for i in idx]))


def _accuracy_oracle(params, bn, items, fb, zm, zm2):
    pcm = torch.from_numpy(np.stack([c for c, _ in items]))
    y = torch.tensor([l for _, l in items])
    feats = O.hot_path_features(pcm, fb, zm, zm2)
    with torch.no_grad():
        logits = O.res8_forward(feats, params, bn, training=False)
    return (logits.argmax(1) == y).float().mean().item()


@pytest.mark.parametrize("engine", [0, 1], ids=["fp32", "tcgen05"])
def test_fused_step_learns_like_the_oracle(engine):
    from howl_b200.trainer import Res8TrainStep

    data = _dataset(1200)
    held_out = data["dev"] + data["test"]
    assert len(held_out) > 150 and len(data["train"]) > 800
    fb = O.mel_filterbank(40)
    # ZMUV fit as training/run/train.py:235-240 (on the first clips of the training split)
    total, mean, mean2 = torch.zeros(1), torch.zeros(1), torch.zeros(1)
    for clip, _ in data["train"][:64]:
        total, mean, mean2 = O.zmuv_update(total, mean, mean2, O.standard_audio_transform_f32(torch.from_numpy(clip)[None], fb))
    std = O.zmuv_std(mean, mean2)

    # ---- oracle training run (CPU)
    params, bn = O.res8_init(L, seed=11), O.res8_bn_init()
    m = {k: torch.zeros_like(p) for k, p in params.items()}
    v = {k: torch.zeros_like(p) for k, p in params.items()}
    o_losses = []
    for step, (pcm, y) in enumerate(_batches(data["train"], N_TRAIN_STEPS), 1):
        feats = O.hot_path_features(pcm, fb, mean, mean2)
        loss, _, _ = O.res8_train_step(feats, y, params, bn, m, v, step, 0.01, 1e-5)
        o_losses.append(loss.item())
    o_acc = _accuracy_oracle(params, bn, held_out, fb, mean, mean2)

    # ---- the same run through the fused CUDA step (same init, same batches)
    tr = Res8TrainStep(DEV, num_labels=L, batch=BATCH, samples=T, lr=0.01, weight_decay=1e-5,
                       zmuv=(mean.item(), std.item()), seed=0)
    tr.ctx.set_option("conv_engine", engine)
    tr.params.copy_(O.flatten(O.res8_init(L, seed=11), L).to(DEV))
    g_losses = []
    for pcm, y in _batches(data["train"], N_TRAIN_STEPS):
        g_losses.append(tr.step(pcm.to(DEV), y.to(DEV)).item())
    # evaluate the CUDA-trained weights with the CUDA forward (eval-mode BatchNorm)
    pcm = torch.from_numpy(np.stack([c for c, _ in held_out])).to(DEV)
    y = torch.tensor([l for _, l in held_out])
    feats = tr.ctx.frontend(pcm, tr.fb, "time_major", zmuv=tr.zmuv)
    ws = torch.empty(tr.ctx.res8_workspace_bytes(pcm.shape[0], feats.shape[1], L, False), dtype=torch.uint8, device=DEV)
    logits = tr.ctx.res8_fwd(feats, tr.params, tr.bn_running, tr.nbt, False, ws).cpu()
    g_acc = (logits.argmax(1) == y).float().mean().item()

    # same trajectory at the start (identical batches), same outcome at the end
    np.testing.assert_allclose(g_losses[0], o_losses[0], rtol=1e-4)        # identical first step
    np.testing.assert_allclose(g_losses[:3], o_losses[:3], rtol=3e-2)      # then AdamW's sign-like steps amplify rounding
    assert np.mean(g_losses[-10:]) < 0.5 * np.mean(g_losses[:3]), "training loss did not go down"
    assert abs(np.mean(g_losses[-10:]) - np.mean(o_losses[-10:])) < 0.15
    assert o_acc > 0.9 and g_acc > 0.9 and abs(g_acc - o_acc) <= 0.05, (o_acc, g_acc)
    # and the CUDA-trained weights score the same through the oracle's forward (state_dict interop)
    sd = {k: t.cpu() for k, t in tr.state_dict().items()}
    assert abs(_accuracy_oracle(sd, sd, held_out, fb, mean, mean2) - g_acc) <= 0.02
