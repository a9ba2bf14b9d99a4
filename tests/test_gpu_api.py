"""GPU: the howl-shaped Python surface (transforms, ZMUV, registry/Res8 module, inference engines) on top of the C ABI."""
import json
import os
import random

import numpy as np
import pytest
import torch

from oracle import howl_oracle as O

from conftest import GOLDEN

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0")
RTOL = ATOL = 1e-4


@pytest.fixture(autouse=True)
def _env(monkeypatch):
    from howl_b200.settings import SETTINGS

    monkeypatch.setenv("NUM_MELS", "40")
    monkeypatch.setenv("INFERENCE_SEQUENCE", "[0,1,2]")
    monkeypatch.setenv("INFERENCE_THRESHOLD", "0")
    SETTINGS.reset()
    yield
    SETTINGS.reset()


def test_standard_audio_transform_eval_and_train_draw_order(golden):
    from howl_b200.transform import StandardAudioTransform

    g, v = golden("frontend"), golden("vtlp")
    pcm = torch.from_numpy(g["t8000_pcm"]).to(DEV)
    std = StandardAudioTransform().to(DEV).eval()
    np.testing.assert_allclose(std(pcm).cpu().numpy(), g["t8000_out"], rtol=RTOL, atol=ATOL)
    np.testing.assert_allclose(std(pcm, mels_only=True).cpu().numpy(), g["t8000_mels_only"], rtol=RTOL, atol=ATOL)
    lens = torch.from_numpy(g["lengths_in"])
    assert np.array_equal(std.compute_lengths(lens).numpy(), g["lengths_out"])
    # train mode: the same global-`random` seed must reproduce the reference's coin flips and alphas
    std.train()
    random.seed(7)
    for k in range(4):
        np.testing.assert_allclose(std(pcm).cpu().numpy(), v["train_outs"][k], rtol=RTOL, atol=ATOL)


def test_standard_audio_transform_deltas_only(golden):
    """_execute_op(deltas_only=True) (transform.py:272-280): the reference's own log-mels in, its stacked [B,3,M,F] tensor out."""
    from howl_b200.transform import StandardAudioTransform

    g = golden("frontend")
    std = StandardAudioTransform().to(DEV).eval()
    for tag in ("t8000", "t16000", "t4567", "t1000"):
        mels = torch.from_numpy(g[tag + "_mels_only"]).to(DEV)
        out = std(mels, deltas_only=True)
        assert out.shape == g[tag + "_out"].shape
        assert torch.equal(out[:, 0], mels)
        np.testing.assert_allclose(out.cpu().numpy(), g[tag + "_out"], rtol=1e-5, atol=1e-5)
        assert torch.equal(std(mels, deltas_only=True, mels_only=True), mels)
    with pytest.raises(ValueError):
        std(torch.zeros(40, 41, device=DEV), deltas_only=True)


def test_zmuv_transform_update_and_state_dict(golden):
    from howl_b200.transform import ZmuvTransform

    z, g = golden("zmuv"), golden("frontend")
    zm = ZmuvTransform().to(DEV)
    for x in [g["t8000_out"][i:i + 1] for i in range(3)] + [g["speech_out"]]:
        zm.update(torch.from_numpy(x).to(DEV))
    np.testing.assert_allclose(zm.total.cpu().numpy(), z["total"])
    np.testing.assert_allclose(zm.mean.cpu().numpy(), z["mean"], rtol=1e-5)
    np.testing.assert_allclose(zm.mean2.cpu().numpy(), z["mean2"], rtol=1e-5)
    assert set(zm.state_dict()) == {"total", "mean", "mean2"}
    out = zm(torch.from_numpy(z["fwd_in"]).to(DEV)).cpu().numpy()
    np.testing.assert_allclose(out, z["fwd_out"], rtol=1e-5, atol=1e-5)


def test_spec_augment_replays_reference_draws(golden):
    from howl_b200.transform import SpecAugmentTransform

    sa = golden("specaugment")
    spec = SpecAugmentTransform().train()
    random.seed(11)
    out = spec(torch.from_numpy(sa["in"]).to(DEV)).cpu().numpy()
    assert np.array_equal(out, sa["out"])
    random.seed(11)
    rects = spec.draw_rects(sa["in"].shape[0], 40, sa["in"].shape[3])
    assert np.array_equal(rects.numpy(), sa["rects"])


def test_registry_and_res8_state_dict_keys(golden):
    from howl_b200.model import RegisteredModel

    assert "res8" in RegisteredModel.registered_names()
    model = RegisteredModel.find_registered_class("res8")(4)
    g = golden("res8_heyfirefox")
    want = sorted(k[3:] for k in g if k.startswith("sd."))
    assert sorted(model.state_dict().keys()) == want
    assert sum(p.numel() for p in model.parameters()) == 109939
    assert model.streaming() is model and model.static() is model and model.compute_length(5) == 5


def test_res8_module_trains_like_the_reference_loop(golden):
    """The loop body of training/run/train.py:288-302 with torch's CrossEntropyLoss and torch.optim.AdamW over
    model.parameters(), checked against the golden reference run."""
    from howl_b200.model import RegisteredModel
    from howl_b200.transform import StandardAudioTransform, ZmuvTransform

    g = golden("res8_train")
    L = 12
    model = RegisteredModel.find_registered_class("res8")(L)
    model.load_state_dict({k[5:]: torch.from_numpy(v) for k, v in g.items() if k.startswith("init.")})
    model = model.to(DEV).streaming()
    std = StandardAudioTransform().to(DEV).eval()
    zmuv = ZmuvTransform().to(DEV)
    zmuv.mean, zmuv.mean2 = torch.from_numpy(g["zmuv_mean"]).to(DEV), torch.from_numpy(g["zmuv_mean2"]).to(DEV)
    opt = torch.optim.AdamW(model.parameters(), float(g["lr"]), weight_decay=float(g["wd"]))
    crit = torch.nn.CrossEntropyLoss()
    pcm, labels = torch.from_numpy(g["pcm"]).to(DEV), torch.from_numpy(g["labels"]).to(DEV)
    model.train()
    for step in (1, 2):
        lengths = std.compute_lengths(torch.full((pcm.size(0),), pcm.size(1)))
        scores = model(zmuv(std(pcm)), lengths)
        loss = crit(scores, labels)
        opt.zero_grad()
        model.zero_grad()
        loss.backward()
        np.testing.assert_allclose(loss.item(), g[f"step{step}.loss"], rtol=RTOL, atol=ATOL)
        np.testing.assert_allclose(scores.detach().cpu().numpy(), g[f"step{step}.logits"], rtol=RTOL, atol=ATOL)
        for k, p in model.named_parameters():
            # default engine = tcgen05 bf16x3: conv gradients carry ReLU flip noise (tests/test_gpu_parity.py::_assert_grads)
            want = g[f"step{step}.grad.{k}"]
            rel = np.linalg.norm(p.grad.cpu().numpy() - want) / np.linalg.norm(want)
            assert rel <= (1e-3 if k.startswith("output") else 3e-2), (k, rel)
        opt.step()
        sd = model.state_dict()
        for i in range(1, 7):
            np.testing.assert_allclose(sd[f"bn{i}.running_var"].cpu().numpy(), g[f"step{step}.sd.bn{i}.running_var"], rtol=1e-4, atol=1e-5)
            assert int(sd[f"bn{i}.num_batches_tracked"]) == step
        # teacher-force the parameters (AdamW's first steps are ill-conditioned in g; see test_oracle_golden)
        model.load_state_dict({k: torch.from_numpy(g[f"step{step}.sd.{k}"]) for k in sd})


@pytest.mark.parametrize("batched", ["infer", "infer_batched", "ingest_frame"])
def test_frame_inference_engine_known_answers(golden, batched):
    """the shipped hey-fire-fox res8 + zmuv through FrameInferenceEngine(500, 63)."""
    from howl_b200.inference import FrameInferenceEngine, SimpleContext
    from howl_b200.model import RegisteredModel
    from howl_b200.transform import ZmuvTransform

    g = golden("res8_heyfirefox")
    meta = json.load(open(os.path.join(GOLDEN, "meta.json")))["traces"]
    model = RegisteredModel.find_registered_class("res8")(4)
    model.load_state_dict({k[3:]: torch.from_numpy(v) for k, v in g.items() if k.startswith("sd.")})
    model = model.to(DEV).eval()
    zmuv = ZmuvTransform()
    zmuv.load_state_dict({k: torch.from_numpy(g[f"zmuv.{k}"]) for k in ("total", "mean", "mean2")})
    zmuv = zmuv.to(DEV)
    ctx = SimpleContext.for_vocab(["hey", "fire", "fox"])
    assert (ctx.num_labels, ctx.negative_label, ctx.blank_label) == (4, 3, -1)
    for name in ("hey_fire_fox", "hello_world"):
        engine = FrameInferenceEngine(500, 63, model, zmuv, ctx)
        audio = torch.from_numpy(g[f"trace_{name}_pcm"]).to(DEV)
        if batched == "ingest_frame":      # the live-streaming path: one window at a time, as howl_client.py:94-105 drives it
            from howl_b200.inference import stride

            detected = False
            for window in stride(audio, 500, 63, 16000):
                engine.ingest_frame(window, engine.curr_time)
                engine.curr_time += 63
                if engine.sequence_present(engine.curr_time):
                    detected = True
                    break
        else:
            detected = getattr(engine, batched)(audio)
        assert detected == meta[name]["detected"]
        assert [int(l) for _, l in engine.label_history] == meta[name]["labels"]


def test_device_batchifier_matches_reference_batches():
    """clips resident in HBM, plan drawn on the host with the reference's draws, one gather kernel."""
    import howl_b200
    from howl_b200.batchifier import DeviceFrameBatchifier
    from test_host_logic import _batchifier_inputs

    g, clips = _batchifier_inputs()
    ctx = howl_b200.Context(DEV, n_mels=40)
    dev_clips = torch.from_numpy(g["clips"]).to(DEV)
    for trial in range(6):
        random.seed(100 + trial)
        b = DeviceFrameBatchifier(3, positive_sample_prob=[0.5, 0.9, 0.1][trial % 3])
        audio, labels, lengths = b(ctx, dev_clips, clips * 2)
        assert np.array_equal(audio.cpu().numpy(), g[f"t{trial}.audio"])
        assert np.array_equal(labels.cpu().numpy(), g[f"t{trial}.labels"])
        assert np.array_equal(lengths.cpu().numpy(), g[f"t{trial}.lengths"])
    ctx.close()


def test_autograd_two_forwards_before_backward():
    """The drop-in nn.Module under ordinary PyTorch usage: a second training forward (and an eval pass) between a forward and
    its backward must not disturb that backward -- each forward carries its own activation workspace on the autograd ctx."""
    from howl_b200.model import RegisteredModel

    torch.manual_seed(0)
    L = 4
    model = RegisteredModel.find_registered_class("res8")(L).to(DEV).train()
    xa = torch.randn(6, 3, 40, 41, device=DEV)
    xb = torch.randn(9, 3, 40, 81, device=DEV)       # larger batch, longer clip: would have reallocated a shared workspace
    ya, yb = torch.randint(0, L, (6,), device=DEV), torch.randint(0, L, (9,), device=DEV)

    def grad_of(x, y):
        model.zero_grad()
        torch.nn.functional.cross_entropy(model(x, None), y).backward()
        return torch.cat([p.grad.reshape(-1) for p in model.parameters()]).clone()

    ga, gb = grad_of(xa, ya), grad_of(xb, yb)
    model.zero_grad()
    la = torch.nn.functional.cross_entropy(model(xa, None), ya)
    lb = torch.nn.functional.cross_entropy(model(xb, None), yb)
    with torch.no_grad():
        model.eval()
        model(xb, None)
        model.train()
    la.backward()
    got_a = torch.cat([p.grad.reshape(-1) for p in model.parameters()]).clone()
    model.zero_grad()
    lb.backward()
    got_b = torch.cat([p.grad.reshape(-1) for p in model.parameters()]).clone()
    assert torch.allclose(got_a, ga, rtol=1e-3, atol=1e-4 * ga.abs().max().item())      # atomics reorder between runs
    assert torch.allclose(got_b, gb, rtol=1e-3, atol=1e-4 * gb.abs().max().item())
    # parameters changed between forward and backward -> refuse instead of differentiating against the wrong weights
    loss = torch.nn.functional.cross_entropy(model(xa, None), ya)
    with torch.no_grad():
        model.conv1.weight.add_(1e-3)
    with pytest.raises(RuntimeError, match="modified"):
        loss.backward()


def test_converted_static_model_batches_windows():
    from howl_b200.model import ConvertedStaticModel, RegisteredModel

    torch.manual_seed(1)
    inner = RegisteredModel.find_registered_class("res8")(4).to(DEV).eval()
    conv = ConvertedStaticModel(inner, 41, 10).eval()
    assert "converted" in RegisteredModel.registered_names() and conv.compute_length(81) == 4
    x = torch.randn(3, 3, 40, 101, device=DEV)
    with torch.no_grad():
        got = conv(x, None)
        # the reference loop (base.py:52-62), window by window
        wins, idx = [x[:, :, :, 41:]], 10
        while x[:, :, :, idx:idx + 41].size(3) == 41:
            wins.append(x[:, :, :, idx:idx + 41])
            idx += 10
        want = torch.stack([inner(w.contiguous(), None) for w in wins])
    assert got.shape == want.shape == (len(wins), 3, 4)
    assert torch.allclose(got, want, rtol=1e-5, atol=1e-6)
    for name in ("small-cnn", "seq-cnn", "gru"):
        with pytest.raises(NotImplementedError):
            RegisteredModel.find_registered_class(name)(4)


def test_trainer_epoch_applies_the_reference_augmentations(monkeypatch):
    """Trainer.train_epoch = the loop of training/run/train.py:280-307: per step the VTLP coin (+ alpha) and the two SpecAugment
    coins (+ per-clip rectangles) come from the global `random` in the reference's order; batches may shrink or grow."""
    from howl_b200.config import TrainingConfig
    from howl_b200.trainer import Trainer
    from howl_b200.transform import SpecAugmentTransform

    cfg = TrainingConfig(**{"num_epochs": 1, "learning_rate": 0.01, "lr_decay": 0.5, "weight_decay": 1e-5,
                            "context_config": {"vocab": ["hey", "fire", "fox"], "token_type": "word", "seed": 3},
                            "model_config": {"architecture": "res8"}})
    tr = Trainer(cfg, device="cuda:0")
    sizes = [(8, 8000), (5, 8000), (12, 8000)]          # smaller last batch, then a larger one: workspace follows
    batches = [O.synthetic_batch(b, t, 4, seed=i) for i, (b, t) in enumerate(sizes)]
    random.seed(11)
    loss = tr.train_epoch(batches, zmuv=(-1.8, 3.9))
    after = random.random()
    assert np.isfinite(loss) and tr.step_obj.step_count == 3 and abs(tr.step_obj.lr - 0.005) < 1e-12
    # replay the draws the reference loop makes for the same batches
    random.seed(11)
    spec = SpecAugmentTransform().train()
    for b, t in sizes:
        if random.random() < 0.75:
            random.random()
        spec.draw_rects(b, 40, 1 + t // 200)
    assert random.random() == after
    # augment=False consumes no draws
    random.seed(11)
    tr.train_epoch(batches[:1], zmuv=(-1.8, 3.9), augment=False)
    random.seed(11)
    first = random.random()
    random.seed(11)
    tr.train_epoch(batches[:1], zmuv=(-1.8, 3.9), augment=False)
    assert random.random() == first


def test_device_wave_augmentation_matches_reference_chain_and_noise_statistics():
    """SURVEY §8f row 3 on the device: (1) mixer + time shift + batchifier fused into one gather kernel reproduce the reference chain's
    batches bit for bit (tests/golden/augment.npz); (2) the in-kernel Philox noise has the distributions NoiseTransform draws from
    (N(0, sigma) clamped; +-1 impulses with probability p / 2 each) and respects the final clamp to [-1, 1]."""
    import howl_b200
    from howl_b200.batchifier import DeviceFrameBatchifier, DeviceWaveAugmenter
    from test_host_logic import _batchifier_inputs

    g, clips = _batchifier_inputs()
    a = dict(np.load(os.path.join(GOLDEN, "augment.npz")))
    ctx = howl_b200.Context(DEV, n_mels=40)
    dev_clips, dev_bg = torch.from_numpy(g["clips"]).to(DEV), torch.from_numpy(a["bg"]).to(DEV)
    for trial in range(8):
        random.seed(500 + trial)
        aug = DeviceWaveAugmenter(DeviceFrameBatchifier(3, positive_sample_prob=[0.5, 0.9, 0.1][trial % 3]), a["bg_lengths"].tolist(), noise=False)
        audio, labels, lengths = aug(ctx, dev_clips, clips * 2, dev_bg)
        assert np.array_equal(audio.cpu().numpy(), a[f"t{trial}.audio"])
        assert np.array_equal(labels.cpu().numpy(), a[f"t{trial}.labels"]) and np.array_equal(lengths.cpu().numpy(), a[f"t{trial}.lengths"])
    # ---- noise statistics on a constant signal
    B, T = 8, 200000
    base = torch.full((B * T,), 0.25, device=DEV)
    starts = torch.arange(B, device=DEV) * T
    counts, dst = torch.full((B,), T, dtype=torch.int64, device=DEV), torch.zeros(B, dtype=torch.int64, device=DEV)
    sigma = torch.tensor([0.0, 0.001, 0.01, 0.1, 0.0, 0.0, 0.5, 0.05], device=DEV)
    sp = torch.tensor([0.0, 0.0, 0.0, 0.0, 0.01, 0.2, 0.0, 0.1], device=DEV)
    out = ctx.batch_gather_aug(base, starts, counts, dst, T, sigma=sigma, sp_prob=sp, seed=7).cpu()
    other = ctx.batch_gather_aug(base, starts, counts, dst, T, sigma=sigma, sp_prob=sp, seed=8).cpu()
    assert torch.equal(out[0], torch.full((T,), 0.25)) and not torch.equal(out[1], other[1]) and out.abs().max() <= 1.0
    for r in (1, 2, 3):
        d = out[r] - 0.25
        assert abs(d.mean().item()) < 4 * sigma[r].item() / T ** 0.5 and abs(d.std().item() / sigma[r].item() - 1) < 0.02
    for r in (4, 5):
        d = out[r] - 0.25
        up, down = (d > 0.5).float().mean().item(), (d < -0.5).float().mean().item()
        p = sp[r].item() / 2
        assert abs(up - p * (1 - p)) < 5 * (p / T) ** 0.5 and abs(down - p * (1 - p)) < 5 * (p / T) ** 0.5
        assert out[r].max() <= 1.0 and out[r].min() >= -0.75 - 1e-6            # 0.25 + 1 clamps to 1, 0.25 - 1 = -0.75
    assert (out[6] == 1.0).any() and out[6].min() >= -0.75 - 1e-6               # the noise mask itself is clamped to [-1, 1] first
    ctx.close()


@pytest.mark.parametrize("arch", ["lstm", "mobilenet", "las"])
def test_trainer_runs_every_accelerated_architecture(arch):
    """Trainer.train_epoch drives the CUDA step of every accelerated frame-objective model with the reference's augmentation draws
    (VTLP bank for the step, SpecAugment rectangles); a smaller batch uses a prefix, a larger one rebuilds the step and keeps the state."""
    from howl_b200.config import TrainingConfig
    from howl_b200.trainer import Trainer

    cfg = TrainingConfig(**{"num_epochs": 1, "learning_rate": 0.002, "lr_decay": 0.5, "weight_decay": 1e-5,
                            "context_config": {"vocab": ["hey", "fire", "fox"], "token_type": "word", "seed": 3},
                            "model_config": {"architecture": arch}})
    tr = Trainer(cfg, device="cuda:0")
    sizes = [(12, 16000), (8, 16000), (16, 16000)]
    batches = [O.synthetic_batch(b, t, 4, seed=i) for i, (b, t) in enumerate(sizes)]
    random.seed(5)
    loss = tr.train_epoch(batches[:2], zmuv=(-1.8, 3.9))
    assert np.isfinite(loss) and tr.step_obj.step_count == 2 and tr.step_obj.batch == 12
    before = tr.step_obj.params.clone()
    loss2 = tr.train_epoch(batches[2:], zmuv=(-1.8, 3.9), augment=False)
    assert np.isfinite(loss2) and tr.step_obj.step_count == 3 and tr.step_obj.batch == 16
    delta = (tr.step_obj.params - before).abs().max().item()
    assert 0 < delta < 0.05                       # the rebuilt step continued from the trained parameters (one AdamW step away)
    with pytest.raises(NotImplementedError):
        Trainer(TrainingConfig(**{"context_config": {"vocab": ["a"], "token_type": "word"}, "model_config": {"architecture": "gru"}}))
