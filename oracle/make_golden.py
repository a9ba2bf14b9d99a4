"""TEST INFRASTRUCTURE ONLY -- generate ``tests/golden/*.npz`` from the REAL reference.

Run in the build container (``/root/reference`` mounted read-only):

    python oracle/make_golden.py

It imports castorini/howl unmodified (three in-process shims, SURVEY.md §8c: pydantic v1
``BaseSettings``, stub ``librosa``, stub ``coloredlogs``), drives the reference's own
``StandardAudioTransform`` / ``ZmuvTransform`` / ``SpecAugmentTransform`` / ``Res8`` /
``FrameInferenceEngine`` / ``torch.optim.AdamW`` on seeded inputs and real-weight checkpoints
(``howl-models``), and freezes inputs + outputs as small float32 fixtures.  The GPU box has no
``/root/reference``; tests there read only the committed fixtures.
"""
import json
import os
import random
import sys
import types

import numpy as np

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


def _install_shims():
    import pydantic
    import pydantic.v1

    pydantic.BaseSettings = pydantic.v1.BaseSettings
    for n in ("librosa", "librosa.effects", "librosa.filters", "librosa.util", "librosa.core"):
        sys.modules[n] = types.ModuleType(n)
    lib = sys.modules["librosa"]
    lib.effects, lib.filters = sys.modules["librosa.effects"], sys.modules["librosa.filters"]
    lib.util, lib.core = sys.modules["librosa.util"], sys.modules["librosa.core"]
    sys.modules["librosa.filters"].get_window = None
    cl = types.ModuleType("coloredlogs")
    cl.install = lambda **k: None
    sys.modules["coloredlogs"] = cl
    sys.path.insert(0, REF)


def _read_wav(path):
    from scipy.io import wavfile

    sr, data = wavfile.read(path)
    assert sr == 16000 and data.dtype == np.int16
    return (data.astype(np.float32) / 32768.0)


def main():
    os.makedirs(OUT, exist_ok=True)
    os.environ.update({
        "NUM_MELS": "40", "MAX_WINDOW_SIZE_SECONDS": "0.5", "INFERENCE_SEQUENCE": "[0,1,2]",
        "VOCAB": '["hey","fire","fox"]', "INFERENCE_THRESHOLD": "0",
    })
    _install_shims()
    import torch

    torch.set_num_threads(1)  # deterministic reductions for the fixtures
    from howl.context import InferenceContext
    from howl.data.transform.operator import ZmuvTransform
    from howl.data.transform.transform import SpecAugmentTransform, StandardAudioTransform, create_vtlp_fb_matrix
    from howl.model import RegisteredModel
    from howl.model.inference import FrameInferenceEngine
    from howl.settings import SETTINGS

    assert SETTINGS.audio_transform.num_mels == 40
    meta = {"reference_commit": "4ba5f42e", "torch": torch.__version__}

    # ------------------------------------------------------------------ frontend (eval)
    std = StandardAudioTransform().eval()
    fe = {}
    g = torch.Generator().manual_seed(1234)
    for tag, (b, t) in {"t8000": (3, 8000), "t16000": (2, 16000), "t4567": (2, 4567), "t1000": (1, 1000)}.items():
        pcm = (torch.randn(b, t, generator=g) * 0.1).clamp_(-1, 1)
        fe[f"{tag}_pcm"] = pcm.numpy()
        fe[f"{tag}_out"] = std(pcm).numpy()
        fe[f"{tag}_mels_only"] = std(pcm, mels_only=True).numpy()
    speech = _read_wav(f"{REF}/test/test_data/sphinx_keyword_detector/hey_fire_fox.wav")
    fe["speech_pcm"] = speech[None, 8000:24000].copy()
    fe["speech_out"] = std(torch.from_numpy(fe["speech_pcm"])).numpy()
    fe["zeros_pcm"] = np.zeros((1, 2000), np.float32)
    fe["zeros_out"] = std(torch.zeros(1, 2000)).numpy()
    fe["fb"] = std.spec_transform.mel_scale.fb.numpy()
    fe["window"] = std.spec_transform.spectrogram.window.numpy()
    lens = torch.tensor([0, 1, 511, 512, 513, 711, 712, 8000, 16000, 35774, 16001, 199], dtype=torch.int64)
    fe["lengths_in"] = lens.numpy()
    fe["lengths_out"] = std.compute_lengths(lens).numpy()
    np.savez_compressed(os.path.join(OUT, "frontend.npz"), **fe)

    # ------------------------------------------------------------------ VTLP filterbanks + train-mode draw order
    vt = {}
    alphas = [0.9, 0.95, 1.0, 1.05, 1.0999, 0.9731]
    vt["alphas"] = np.array(alphas, np.float64)
    for i, a in enumerate(alphas):
        vt[f"fb_{i}"] = create_vtlp_fb_matrix(257, 0.0, 8000.0, 40, 16000, a).numpy()
    # train-mode forward: replay the global-`random` draws (coin at transform.py:93, alpha at :441)
    std_train = StandardAudioTransform().train()
    pcm = torch.from_numpy(fe["t8000_pcm"])
    random.seed(7)
    outs, coins, drawn = [], [], []
    state = random.getstate()
    for _ in range(4):
        outs.append(std_train(pcm).numpy())
    random.setstate(state)
    for _ in range(4):
        c = random.random()
        coins.append(c)
        drawn.append(random.random() * 0.2 + 0.9 if c < 0.75 else -1.0)
    vt["train_outs"] = np.stack(outs)
    vt["train_coins"] = np.array(coins)
    vt["train_alphas"] = np.array(drawn)
    np.savez_compressed(os.path.join(OUT, "vtlp.npz"), **vt)

    # ------------------------------------------------------------------ ZMUV
    zm = {}
    z = ZmuvTransform()
    feats = [torch.from_numpy(fe["t8000_out"][i:i + 1]) for i in range(3)] + [torch.from_numpy(fe["speech_out"])]
    for f in feats:
        z.update(f)
    zm["total"], zm["mean"], zm["mean2"] = z.total.numpy(), z.mean.numpy(), z.mean2.numpy()
    zm["std"] = z.std.numpy()
    zm["fwd_in"] = fe["t8000_out"]
    zm["fwd_out"] = z(torch.from_numpy(fe["t8000_out"])).numpy()
    np.savez_compressed(os.path.join(OUT, "zmuv.npz"), **zm)

    # ------------------------------------------------------------------ SpecAugment (host draws replayed)
    sa = {}
    spec = SpecAugmentTransform().train()
    x = torch.from_numpy(fe["t8000_out"]).clone()
    random.seed(11)
    state = random.getstate()
    sa["in"] = x.numpy().copy()
    sa["out"] = spec(x.clone()).numpy()
    # replay: per param -> coin; if taken, per-sample (len, start) draws (transform.py:310-326)
    random.setstate(state)
    rects = np.zeros((x.size(0), 4), np.int64)  # f0, flen, t0, tlen
    if random.random() < 0.75:
        for b in range(x.size(0)):
            f = random.randrange(0, 10)
            f0 = random.randrange(0, x.size(2) - f)
            rects[b, 0], rects[b, 1] = f0, f
    if random.random() < 0.75:
        for b in range(x.size(0)):
            t = random.randrange(0, 75)
            if x.size(3) - t <= 0:
                continue
            t0 = random.randrange(0, x.size(3) - t)
            rects[b, 2], rects[b, 3] = t0, t
    sa["rects"] = rects
    np.savez_compressed(os.path.join(OUT, "specaugment.npz"), **sa)

    # ------------------------------------------------------------------ res8 with the shipped hey-fire-fox checkpoint
    ck = f"{REF}/howl-models/howl/hey-fire-fox"
    ctx = InferenceContext(SETTINGS.training.vocab, token_type="word", use_blank=False)
    assert ctx.num_labels == 4
    model = RegisteredModel.find_registered_class("res8")(ctx.num_labels).eval()
    sd = torch.load(f"{ck}/model-best.pt.bin", map_location="cpu")
    model.load_state_dict(sd)
    zmuv = ZmuvTransform()
    zmuv.load_state_dict(torch.load(f"{ck}/zmuv.pt.bin", map_location="cpu"))
    r8 = {f"sd.{k}": v.numpy() for k, v in sd.items()}
    r8.update({"zmuv.total": zmuv.total.numpy(), "zmuv.mean": zmuv.mean.numpy(), "zmuv.mean2": zmuv.mean2.numpy()})
    wins = torch.stack([torch.from_numpy(speech[s:s + 8000]) for s in range(0, 35774 - 8000, 1008)][:24])
    r8["pcm"] = wins.numpy()
    with torch.no_grad():
        feats = zmuv(std(wins))
        r8["feats"] = feats.numpy()
        r8["logits"] = model(feats, std.compute_lengths(torch.full((wins.size(0),), 8000))).numpy()
    # known-answer traces of SURVEY App. B.3 through the reference's FrameInferenceEngine
    traces = {}
    for name, path in {
        "hey_fire_fox": f"{REF}/test/test_data/sphinx_keyword_detector/hey_fire_fox.wav",
        "hello_world": f"{REF}/test/test_data/sphinx_keyword_detector/hello_world.wav",
    }.items():
        engine = FrameInferenceEngine(500, 63, model, zmuv, ctx)
        audio = torch.from_numpy(_read_wav(path))
        detected = engine.infer(audio)
        traces[name] = {"detected": bool(detected), "labels": [int(l) for _, l in engine.label_history]}
        r8[f"trace_{name}_pcm"] = audio.numpy()
    meta["traces"] = traces
    np.savez_compressed(os.path.join(OUT, "res8_heyfirefox.npz"), **r8)

    # ------------------------------------------------------------------ res8 train steps (reference module + torch AdamW)
    ts = {}
    torch.manual_seed(5)
    num_labels, bsz, t = 12, 8, 16000
    model = RegisteredModel.find_registered_class("res8")(num_labels).train().streaming()
    ts.update({f"init.{k}": v.detach().numpy().copy() for k, v in model.state_dict().items()})
    g = torch.Generator().manual_seed(99)
    pcm = (torch.randn(bsz, t, generator=g) * 0.1).clamp_(-1, 1)
    labels = torch.randint(0, num_labels, (bsz,), generator=g)
    zm_mean, zm_mean2 = torch.tensor([-2.0166]), torch.tensor([-2.0166 ** 2 + 3.9955 ** 2])
    zmuv = ZmuvTransform()
    zmuv.mean, zmuv.mean2 = zm_mean, zm_mean2
    lr, wd = 0.01, 1e-5
    opt = torch.optim.AdamW(model.parameters(), lr, weight_decay=wd)
    crit = torch.nn.CrossEntropyLoss()
    ts["pcm"], ts["labels"] = pcm.numpy(), labels.numpy()
    ts["zmuv_mean"], ts["zmuv_mean2"] = zm_mean.numpy(), zm_mean2.numpy()
    ts["lr"], ts["wd"] = np.float64(lr), np.float64(wd)
    std.eval()
    for step in range(1, 4):
        feats = zmuv(std(pcm))
        lengths = std.compute_lengths(torch.full((bsz,), t))
        scores = model(feats, lengths)
        loss = crit(scores, labels)
        opt.zero_grad()
        model.zero_grad()
        loss.backward()
        ts[f"step{step}.loss"] = loss.detach().numpy()
        ts[f"step{step}.logits"] = scores.detach().numpy()
        for k, p in model.named_parameters():
            ts[f"step{step}.grad.{k}"] = p.grad.numpy().copy()
        opt.step()
        for k, v in model.state_dict().items():
            ts[f"step{step}.sd.{k}"] = v.detach().numpy().copy()
    np.savez_compressed(os.path.join(OUT, "res8_train.npz"), **ts)

    # ------------------------------------------------------------------ lstm / seq-lstm (rnn.py:41-91)
    ls = {}
    gsc = f"{REF}/howl-models/howl/experiments/commands_recognition/lstm/0"
    lsd = torch.load(f"{gsc}/model-best.pt.bin", map_location="cpu")
    lz = torch.load(f"{gsc}/zmuv.pt.bin", map_location="cpu")
    lstm = RegisteredModel.find_registered_class("lstm")(30).eval()
    lstm.load_state_dict(lsd)
    ls.update({f"sd.{k}": v.numpy() for k, v in lsd.items()})
    ls.update({"zmuv.mean": lz["mean"].numpy(), "zmuv.mean2": lz["mean2"].numpy()})
    wins16 = torch.stack([torch.from_numpy(speech[s:s + 16000]) for s in (0, 4000, 9000, 15000, 19000)])
    zl = ZmuvTransform()
    zl.load_state_dict(lz)
    with torch.no_grad():
        f16 = zl(std(wins16))
        len16 = std.compute_lengths(torch.full((wins16.size(0),), 16000))
        ls["pcm"], ls["lengths"] = wins16.numpy(), len16.numpy()
        ls["logits"] = lstm(f16, len16).numpy()
        # ragged lengths (descending, as pack_padded_sequence demands)
        ragged = torch.tensor([78, 70, 41, 12, 1])
        ls["ragged_lengths"] = ragged.numpy()
        ls["ragged_logits"] = lstm(f16, ragged).numpy()
        # seq-lstm with the same weights: per-frame logits and the carried streaming state over two calls
        seq = RegisteredModel.find_registered_class("seq-lstm")(30).eval().streaming()
        seq.load_state_dict(lsd)
        out1 = seq(f16, len16)
        h1, c1 = seq.streaming_state
        out2 = seq(f16, len16)
        ls["seq_out1"], ls["seq_out2"] = out1.numpy(), out2.numpy()
        ls["seq_h1"], ls["seq_c1"] = h1.numpy(), c1.numpy()
    # two reference train steps (frame objective, torch AdamW), seeded init
    torch.manual_seed(21)
    lm = RegisteredModel.find_registered_class("lstm")(12).train()
    ls.update({f"init.{k}": v.detach().numpy().copy() for k, v in lm.state_dict().items()})
    g2 = torch.Generator().manual_seed(5)
    lp = (torch.randn(6, 8000, generator=g2) * 0.1).clamp_(-1, 1)
    ly = torch.randint(0, 12, (6,), generator=g2)
    ll = std.compute_lengths(torch.full((6,), 8000))
    ls["t_pcm"], ls["t_labels"], ls["t_lengths"] = lp.numpy(), ly.numpy(), ll.numpy()
    lopt = torch.optim.AdamW(lm.parameters(), 0.01, weight_decay=1e-5)
    for stp in (1, 2):
        sc = lm(zl(std(lp)), ll)
        lo = torch.nn.functional.cross_entropy(sc, ly)
        lopt.zero_grad()
        lo.backward()
        ls[f"step{stp}.loss"], ls[f"step{stp}.logits"] = lo.detach().numpy(), sc.detach().numpy()
        for k, prm in lm.named_parameters():
            ls[f"step{stp}.grad.{k}"] = prm.grad.numpy().copy()
        lopt.step()
        for k, v in lm.state_dict().items():
            ls[f"step{stp}.sd.{k}"] = v.detach().numpy().copy()
    # two reference CTC steps with the STREAMING seq-lstm (train.py:246,253,294-302): state carried between the steps
    torch.manual_seed(33)
    sm = RegisteredModel.find_registered_class("seq-lstm")(5).train().streaming()
    ls.update({f"ctc.init.{k}": v.detach().numpy().copy() for k, v in sm.state_dict().items()})
    ctc = torch.nn.CTCLoss(4)
    tg = torch.tensor([[0, 1, 2], [1, 1, 3], [2, 3, 3], [0, 3, 3], [3, 2, 1], [0, 0, 0]])     # padded with the negative label 3
    tl = torch.tensor([3, 2, 1, 1, 3, 3])
    ls["ctc.targets"], ls["ctc.target_lengths"] = tg.numpy(), tl.numpy()
    sopt = torch.optim.AdamW(sm.parameters(), 0.01, weight_decay=1e-5)
    for stp in (1, 2):
        sc = sm(zl(std(lp)), ll)
        lo = ctc(torch.nn.functional.log_softmax(sc, -1), tg, ll, tl)
        sopt.zero_grad()
        lo.backward()
        ls[f"ctc.step{stp}.loss"], ls[f"ctc.step{stp}.scores"] = lo.detach().numpy(), sc.detach().numpy()
        ls[f"ctc.step{stp}.h"], ls[f"ctc.step{stp}.c"] = sm.streaming_state[0].numpy(), sm.streaming_state[1].numpy()
        for k, prm in sm.named_parameters():
            ls[f"ctc.step{stp}.grad.{k}"] = prm.grad.numpy().copy()
        sopt.step()
        for k, v in sm.state_dict().items():
            ls[f"ctc.step{stp}.sd.{k}"] = v.detach().numpy().copy()
    np.savez_compressed(os.path.join(OUT, "lstm.npz"), **ls)

    # ------------------------------------------------------------------ WakeWordFrameBatchifier (batchifier.py:37-118)
    from howl.data.common.example import WakeWordClipExample
    from howl.data.common.label import FrameLabelData
    from howl.data.common.metadata import AudioClipMetadata
    from howl.data.transform.batchifier import WakeWordFrameBatchifier

    bt = {}
    rngb = np.random.default_rng(17)
    maps = [{}, {300.0: 0, 650.0: 1}, {420.5: 2}, {}, {100.0: 0, 480.0: 1, 900.0: 2}, {250.0: 1}, {}, {700.0: 0, 1500.0: 2}]
    lens = [16000, 16000, 9000, 3000, 20000, 5000, 8000, 30000]
    clips_np = [rngb.standard_normal(n).astype(np.float32) * 0.1 for n in lens]
    exs = [WakeWordClipExample(FrameLabelData(m, [], []), AudioClipMetadata(path=".", phone_strings=None, words=None,
                                                                            phone_end_timestamps=None, end_timestamps=None,
                                                                            transcription=""),
                               torch.from_numpy(c), 16000) for m, c in zip(maps, clips_np)]
    bt["clips"] = np.concatenate(clips_np)
    bt["lengths"] = np.array(lens)
    meta["batchifier_maps"] = [[[k, v] for k, v in m.items()] for m in maps]   # insertion order matters (random.choice)
    for trial in range(6):
        random.seed(100 + trial)
        fb_ = WakeWordFrameBatchifier(3, positive_sample_prob=[0.5, 0.9, 0.1][trial % 3])
        out = fb_(exs * 2)
        bt[f"t{trial}.audio"], bt[f"t{trial}.labels"] = out.audio_data.numpy(), out.labels.numpy()
        bt[f"t{trial}.lengths"] = out.lengths.numpy()
    np.savez_compressed(os.path.join(OUT, "batchifier.npz"), **bt)

    # ------------------------------------------------------------------ label-sequence FSM (host logic) cases
    from howl.model.inference import InferenceEngine

    rng = random.Random(2024)
    fsm = []
    dummy = types.SimpleNamespace(streaming_state=None)
    for case in range(300):
        eng = InferenceEngine(dummy, zmuv, ctx)
        eng.sequence = [[0, 1, 2], [0], [2, 1], [0, 0, 1]][case % 4]
        eng.tolerance_window_ms = rng.choice([100, 500, 900])
        eng.inference_window_ms = rng.choice([1000, 2000])
        t, hist = 0.0, []
        for _ in range(rng.randrange(1, 40)):
            t += rng.choice([63, 63, 63, 200, 700])
            hist.append((t, rng.choice([0, 1, 2, 3, 3])))
        eng.label_history = list(hist)
        now = t + rng.choice([0, 63, 1500])
        fsm.append({"sequence": eng.sequence, "tolerance": eng.tolerance_window_ms, "window": eng.inference_window_ms,
                    "history": hist, "now": now, "present": bool(eng.sequence_present(now)),
                    "kept": len(eng.label_history)})
    meta["fsm_cases"] = fsm

    with open(os.path.join(OUT, "meta.json"), "w") as fh:
        json.dump(meta, fh, indent=1, sort_keys=True)
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))
    print(json.dumps(traces), sum(c["present"] for c in fsm), "of", len(fsm), "fsm cases present")


if __name__ == "__main__":
    main()
