"""Pin the CPU oracle against outputs of the real reference (tests/golden, made by oracle/make_golden.py)."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import howl_oracle as O

from conftest import GOLDEN

RTOL = ATOL = 1e-4  # north_star tolerance (fp32), implemented as allclose per SURVEY §8(c)


def test_filterbank_and_window_bit_exact(golden):
    g = golden("frontend")
    fb = O.mel_filterbank(40).numpy()
    assert np.array_equal(fb, g["fb"])
    assert (fb != 0).sum() == 493
    w = torch.hann_window(512, periodic=True).numpy()
    assert np.array_equal(w, g["window"])


def test_compute_lengths_bit_exact(golden):
    g = golden("frontend")
    assert np.array_equal(O.compute_lengths(g["lengths_in"]), g["lengths_out"])
    assert O.num_frames(8000) == 41 and O.num_frames(16000) == 81
    assert O.compute_lengths([8000, 16000]).tolist() == [38, 78]


def test_frontend_f32_matches_reference(golden):
    g = golden("frontend")
    fb = torch.from_numpy(g["fb"])
    for tag in ("t8000", "t16000", "t4567", "t1000", "speech", "zeros"):
        out = O.standard_audio_transform_f32(torch.from_numpy(g[f"{tag}_pcm"]), fb).numpy()
        assert out.shape == g[f"{tag}_out"].shape
        np.testing.assert_allclose(out, g[f"{tag}_out"], rtol=RTOL, atol=ATOL)
        if f"{tag}_mels_only" in g:
            np.testing.assert_allclose(out[:, 0], g[f"{tag}_mels_only"], rtol=RTOL, atol=ATOL)


def test_frontend_f64_restatement_agrees(golden):
    g = golden("frontend")
    for tag in ("t8000", "t4567", "speech", "zeros"):
        out = O.standard_audio_transform_f64(g[f"{tag}_pcm"], g["fb"])
        np.testing.assert_allclose(out, g[f"{tag}_out"], rtol=RTOL, atol=ATOL)


def test_vtlp_filterbanks(golden):
    g = golden("vtlp")
    for i, a in enumerate(g["alphas"]):
        fb = O.vtlp_filterbank(float(a), 40).numpy()
        np.testing.assert_allclose(fb, g[f"fb_{i}"], rtol=1e-6, atol=1e-6)
    # train-mode forward of the reference == oracle fed with the replayed host draws
    fe = golden("frontend")
    pcm = torch.from_numpy(fe["t8000_pcm"])
    for k in range(4):
        a = g["train_alphas"][k]
        fb = O.vtlp_filterbank(float(a), 40) if a > 0 else O.mel_filterbank(40)
        out = O.standard_audio_transform_f32(pcm, fb).numpy()
        np.testing.assert_allclose(out, g["train_outs"][k], rtol=RTOL, atol=ATOL)
    assert (g["train_alphas"] > 0).sum() >= 1


def test_zmuv(golden):
    g = golden("zmuv")
    fe = golden("frontend")
    total, mean, mean2 = torch.zeros(1), torch.zeros(1), torch.zeros(1)
    for x in [fe["t8000_out"][i:i + 1] for i in range(3)] + [fe["speech_out"]]:
        total, mean, mean2 = O.zmuv_update(total, mean, mean2, torch.from_numpy(x))
    np.testing.assert_allclose(total.numpy(), g["total"])
    np.testing.assert_allclose(mean.numpy(), g["mean"], rtol=1e-6)
    np.testing.assert_allclose(mean2.numpy(), g["mean2"], rtol=1e-6)
    out = O.zmuv_forward(torch.from_numpy(g["fwd_in"]), torch.from_numpy(g["mean"]), torch.from_numpy(g["mean2"]))
    np.testing.assert_allclose(out.numpy(), g["fwd_out"], rtol=1e-6, atol=1e-6)


def test_spec_augment(golden):
    g = golden("specaugment")
    out = O.spec_augment_apply(torch.from_numpy(g["in"]).clone(), [tuple(r) for r in g["rects"].tolist()])
    assert np.array_equal(out.numpy(), g["out"])


def _sd(g, prefix):
    return {k[len(prefix):]: torch.from_numpy(v) for k, v in g.items() if k.startswith(prefix)}


def test_res8_eval_real_weights(golden):
    g = golden("res8_heyfirefox")
    sd = _sd(g, "sd.")
    feats = O.hot_path_features(torch.from_numpy(g["pcm"]), O.mel_filterbank(40),
                                torch.from_numpy(g["zmuv.mean"]), torch.from_numpy(g["zmuv.mean2"]))
    np.testing.assert_allclose(feats.numpy(), g["feats"], rtol=RTOL, atol=ATOL)
    logits = O.res8_forward(feats, sd, sd, training=False)
    np.testing.assert_allclose(logits.numpy(), g["logits"], rtol=RTOL, atol=ATOL)
    assert np.array_equal(logits.argmax(1).numpy(), g["logits"].argmax(1))


def test_res8_train_steps(golden):
    g = golden("res8_train")
    L = 12
    init = _sd(g, "init.")
    params = {k: init[k].clone() for k, _ in O.res8_param_shapes(L)}
    bn = {k: v.clone() for k, v in init.items() if k.startswith("bn")}
    m = {k: torch.zeros_like(v) for k, v in params.items()}
    v = {k: torch.zeros_like(p) for k, p in params.items()}
    pcm, labels = torch.from_numpy(g["pcm"]), torch.from_numpy(g["labels"])
    feats = O.hot_path_features(pcm, O.mel_filterbank(40), torch.from_numpy(g["zmuv_mean"]),
                                torch.from_numpy(g["zmuv_mean2"]))
    for step in (1, 2, 3):
        loss, logits, grads = O.res8_train_step(feats, labels, params, bn, m, v, step, float(g["lr"]), float(g["wd"]))
        np.testing.assert_allclose(loss.numpy(), g[f"step{step}.loss"], rtol=RTOL, atol=ATOL)
        np.testing.assert_allclose(logits.numpy(), g[f"step{step}.logits"], rtol=RTOL, atol=ATOL)
        for k in params:
            np.testing.assert_allclose(grads[k].numpy(), g[f"step{step}.grad.{k}"], rtol=1e-3, atol=1e-5)
            # Adam's first steps are sign-like (lr*g/(|g|+eps)): where |g| ~ eps=1e-8 the update is
            # ill-conditioned in g (worst case a sign flip = 2*lr), so end-to-end parameter parity is
            # statistical: >= 99.9 % of elements within 5 % of lr, none beyond 2.5*lr.  The AdamW arithmetic
            # itself is checked tightly with identical grads in test_adamw_restated.  Parameters are then
            # teacher-forced to the reference's so later steps compare like with like.
            want = g[f"step{step}.sd.{k}"]
            diff = np.abs(params[k].numpy() - want)
            assert (diff > 5e-4).mean() <= 1e-3 and diff.max() <= 2.5 * float(g["lr"])
            params[k].copy_(torch.from_numpy(want))
        for k in bn:
            np.testing.assert_allclose(bn[k].numpy(), g[f"step{step}.sd.{k}"], rtol=1e-4, atol=1e-5)


def test_adamw_restated():
    torch.manual_seed(0)
    p0 = {"a": torch.randn(1000), "b": torch.randn(7, 5)}
    ref = {k: torch.nn.Parameter(t.clone()) for k, t in p0.items()}
    opt = torch.optim.AdamW(ref.values(), lr=0.01, weight_decay=1e-2)
    mine = {k: t.clone() for k, t in p0.items()}
    m = {k: torch.zeros_like(t) for k, t in p0.items()}
    v = {k: torch.zeros_like(t) for k, t in p0.items()}
    for step in range(1, 6):
        grads = {k: torch.randn_like(t) * 10 ** (-step) for k, t in p0.items()}
        for k in ref:
            ref[k].grad = grads[k].clone()
        opt.step()
        O.adamw_step(mine, grads, m, v, step, 0.01, 1e-2)
        for k in ref:
            np.testing.assert_allclose(mine[k].numpy(), ref[k].detach().numpy(), rtol=1e-6, atol=1e-7)


def test_flat_layout_roundtrip():
    p = O.res8_init(12, seed=3)
    flat = O.flatten(p, 12)
    assert flat.numel() == 109755 + 45 * 12 + 12
    back = O.unflatten(flat, 12)
    assert all(torch.equal(back[k], p[k]) for k in p)


def test_known_answer_traces_recorded():
    meta = json.load(open(os.path.join(GOLDEN, "meta.json")))
    assert meta["traces"]["hey_fire_fox"]["detected"] is True
    assert meta["traces"]["hey_fire_fox"]["labels"] == [3, 3, 3, 3, 3, 0, 0, 0, 0, 0, 3, 1, 1, 1, 1, 1, 1, 1, 3, 3, 3, 2]
    assert meta["traces"]["hello_world"] == {"detected": False, "labels": [3] * 7}


# ------------------------------------------------------------------------------------------ lstm / seq-lstm
def test_lstm_restatement_equals_torch_nn_lstm():
    torch.manual_seed(0)
    p = O.lstm_init(7, seed=4)
    ref = torch.nn.LSTM(40, 128)
    ref.load_state_dict({k[5:]: v for k, v in p.items() if k.startswith("lstm.")})
    feats = torch.randn(5, 3, 40, 23)
    lengths = torch.tensor([23, 23, 17, 9, 1])
    h_seq, (h, c) = O.lstm_recurrence(feats, p, lengths)
    packed = torch.nn.utils.rnn.pack_padded_sequence(feats[:, 0].permute(2, 0, 1).contiguous(), lengths)
    out, (hn, cn) = ref(packed)
    out, _ = torch.nn.utils.rnn.pad_packed_sequence(out)
    np.testing.assert_allclose(h_seq.detach().numpy(), out.detach().numpy(), rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(h.detach().numpy(), hn[0].detach().numpy(), rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(c.detach().numpy(), cn[0].detach().numpy(), rtol=1e-5, atol=1e-6)


def test_lstm_real_weights_and_streaming_state(golden):
    g = golden("lstm")
    sd = _sd(g, "sd.")
    assert [k for k, _ in O.lstm_param_shapes(30)] == list(sd.keys())
    feats = O.hot_path_features(torch.from_numpy(g["pcm"]), O.mel_filterbank(40), torch.from_numpy(g["zmuv.mean"]),
                                torch.from_numpy(g["zmuv.mean2"]))
    lengths = torch.from_numpy(g["lengths"])
    assert lengths.tolist() == [78] * 5
    np.testing.assert_allclose(O.lstm_forward(feats, sd, lengths).numpy(), g["logits"], rtol=RTOL, atol=ATOL)
    ragged = torch.from_numpy(g["ragged_lengths"])
    np.testing.assert_allclose(O.lstm_forward(feats, sd, ragged).numpy(), g["ragged_logits"], rtol=RTOL, atol=ATOL)
    out1, st = O.lstm_forward(feats, sd, lengths, sequential=True)
    np.testing.assert_allclose(out1.numpy(), g["seq_out1"], rtol=RTOL, atol=ATOL)
    np.testing.assert_allclose(st[0].numpy(), g["seq_h1"][0], rtol=RTOL, atol=ATOL)
    np.testing.assert_allclose(st[1].numpy(), g["seq_c1"][0], rtol=RTOL, atol=ATOL)
    out2, _ = O.lstm_forward(feats, sd, lengths, sequential=True, state=st)
    np.testing.assert_allclose(out2.numpy(), g["seq_out2"], rtol=RTOL, atol=ATOL)


def test_lstm_train_steps(golden):
    g = golden("lstm")
    L = 12
    init = _sd(g, "init.")
    params = {k: init[k].clone() for k, _ in O.lstm_param_shapes(L)}
    assert O.lstm_flatten(params, L).numel() == 121092 + 8 * 257
    m = {k: torch.zeros_like(v) for k, v in params.items()}
    v = {k: torch.zeros_like(p) for k, p in params.items()}
    feats = O.hot_path_features(torch.from_numpy(g["t_pcm"]), O.mel_filterbank(40), torch.from_numpy(g["zmuv.mean"]),
                                torch.from_numpy(g["zmuv.mean2"]))
    labels, lengths = torch.from_numpy(g["t_labels"]), torch.from_numpy(g["t_lengths"])
    for step in (1, 2):
        loss, logits, grads = O.lstm_train_step(feats, labels, lengths, params, m, v, step, 0.01, 1e-5)
        np.testing.assert_allclose(loss.numpy(), g[f"step{step}.loss"], rtol=RTOL, atol=ATOL)
        np.testing.assert_allclose(logits.numpy(), g[f"step{step}.logits"], rtol=RTOL, atol=ATOL)
        for k in params:
            np.testing.assert_allclose(grads[k].numpy(), g[f"step{step}.grad.{k}"], rtol=1e-3, atol=1e-6)
            params[k].copy_(torch.from_numpy(g[f"step{step}.sd.{k}"]))   # teacher-force (AdamW's first steps are sign-like)


def test_seq_lstm_ctc_steps(golden):
    g = golden("lstm")
    L = 5
    init = _sd(g, "ctc.init.")
    params = {k: init[k].clone() for k, _ in O.lstm_param_shapes(L)}
    m = {k: torch.zeros_like(v) for k, v in params.items()}
    v = {k: torch.zeros_like(p) for k, p in params.items()}
    feats = O.hot_path_features(torch.from_numpy(g["t_pcm"]), O.mel_filterbank(40), torch.from_numpy(g["zmuv.mean"]),
                                torch.from_numpy(g["zmuv.mean2"]))
    lengths = torch.from_numpy(g["t_lengths"])
    tg, tl = torch.from_numpy(g["ctc.targets"]), torch.from_numpy(g["ctc.target_lengths"])
    state = None
    for step in (1, 2):
        loss, scores, grads, state = O.seq_lstm_ctc_step(feats, tg, tl, lengths, params, state, 4, m, v, step, 0.01, 1e-5)
        np.testing.assert_allclose(loss.numpy(), g[f"ctc.step{step}.loss"], rtol=RTOL, atol=ATOL)
        np.testing.assert_allclose(scores.numpy(), g[f"ctc.step{step}.scores"], rtol=RTOL, atol=ATOL)
        np.testing.assert_allclose(state[0].numpy(), g[f"ctc.step{step}.h"][0], rtol=RTOL, atol=ATOL)
        for k in params:
            want = g[f"ctc.step{step}.grad.{k}"]
            np.testing.assert_allclose(grads[k].numpy(), want, rtol=1e-3, atol=1e-5 * max(1.0, np.abs(want).max()))
            params[k].copy_(torch.from_numpy(g[f"ctc.step{step}.sd.{k}"]))


def test_mobilenet_restatement_matches_reference(golden):
    """§8 row a10 groundwork: the functional MobileNetClassifier restatement against the reference module with the shipped GSC
    checkpoint.  The fixture holds inputs + logits; the 9 MB of weights are read from the mounted reference (build container) --
    skipped where it is absent (the GPU box has neither the mount nor, yet, a CUDA path for this model)."""
    import hashlib
    import os

    ckpt = "/root/reference/howl-models/howl/experiments/commands_recognition/mobilenet/0/model-best.pt.bin"
    if not os.path.exists(ckpt):
        pytest.skip("reference checkpoint not mounted")
    g = golden("mobilenet")
    sd = torch.load(ckpt, map_location="cpu")
    h = hashlib.sha256()
    for k in sd:
        h.update(k.encode())
        h.update(np.ascontiguousarray(sd[k].numpy()).tobytes())
    assert h.hexdigest() == bytes(g["digest"]).decode()
    feats = torch.from_numpy(g["feats"])
    assert len(O.mobilenet_plan()) == 17
    with torch.no_grad():
        np.testing.assert_allclose(O.mobilenet_forward(feats, sd, train=False).numpy(), g["logits_eval"], rtol=1e-4, atol=1e-4)
        np.testing.assert_allclose(O.mobilenet_forward(feats, sd, train=True).numpy(), g["logits_train"], rtol=1e-4, atol=1e-4)
    # and the frontend of that workspace (zmuv.pt.bin of the run) reproduces the stored features
    f2 = O.zmuv_forward(O.standard_audio_transform_f32(torch.from_numpy(g["pcm"])), torch.from_numpy(g["zmuv_mean"]), torch.from_numpy(g["zmuv_mean2"]))
    np.testing.assert_allclose(f2.numpy(), g["feats"], rtol=1e-4, atol=1e-4)


def test_las_restatement_matches_reference(golden):
    """§8 row a12 groundwork: LASClassifier forward (convs + packed BiLSTM + 4-head fixed attention + MLP) against the reference
    module with the shipped GSC checkpoint, equal-length and ragged (length-sorted) batches; weights from the mounted reference."""
    import hashlib
    import os

    ckpt = "/root/reference/howl-models/howl/experiments/commands_recognition/las/0/model-best.pt.bin"
    if not os.path.exists(ckpt):
        pytest.skip("reference checkpoint not mounted")
    g = golden("las")
    sd = torch.load(ckpt, map_location="cpu")
    h = hashlib.sha256()
    for k in sd:
        h.update(k.encode())
        h.update(np.ascontiguousarray(sd[k].numpy()).tobytes())
    assert h.hexdigest() == bytes(g["digest"]).decode()
    feats = torch.from_numpy(g["feats"])
    assert O.las_lengths(torch.tensor([78, 58, 18])).tolist() == [21, 16, 6]      # (n + 2) // 2, twice
    with torch.no_grad():
        np.testing.assert_allclose(O.las_forward(feats, sd).numpy(), g["logits_full"], rtol=1e-4, atol=1e-4)
        np.testing.assert_allclose(O.las_forward(feats, sd, torch.from_numpy(g["lengths"])).numpy(), g["logits_ragged"], rtol=1e-4, atol=1e-4)


def test_gru_restatement_matches_reference(golden):
    """SimpleGru forward (conv encoder + packed GRU + MLP) against the reference module, seeded weights stored in the fixture."""
    g = golden("gru")
    sd = {k[3:]: torch.from_numpy(np.asarray(g[k])) for k in g if k.startswith("sd.")}
    feats, lengths = torch.from_numpy(g["feats"]), torch.from_numpy(g["lengths"])
    with torch.no_grad():
        np.testing.assert_allclose(O.gru_forward(feats, sd).numpy(), g["logits_eval_full"], rtol=1e-4, atol=1e-5)
        np.testing.assert_allclose(O.gru_forward(feats, sd, lengths).numpy(), g["logits_eval_ragged"], rtol=1e-4, atol=1e-5)
        np.testing.assert_allclose(O.gru_forward(feats, sd, lengths, train=True).numpy(), g["logits_train_ragged"], rtol=1e-4, atol=1e-5)
    assert lengths.tolist() == [78, 68, 43, 23]       # the restatement must not modify its argument (the reference does: `lengths += 4`)


def test_mobilenet_checkpoint_fixture_pins_oracle_forward_and_gradients(golden):
    """tests/golden/mobilenet_ckpt.npz (the shipped GSC checkpoint + the REFERENCE module's logits / loss / gradient norms for it): the oracle's
    fp32 restatement reproduces them, and its bf16-storage restatement stays within the bf16 bar the GPU path is held to."""
    g = golden("mobilenet_ckpt")
    sd = {}
    for k, v in g.items():
        if k.endswith("::bf16"):
            sd[k[:-6]] = torch.from_numpy(v.view(np.int16).copy()).view(torch.bfloat16).to(torch.float32)
        elif k.startswith(("downsample.", "model.")):
            sd[k] = torch.from_numpy(v)
    pcm, labels = torch.from_numpy(g["pcm"]), torch.from_numpy(g["labels"])
    x = O.hot_path_features(pcm, O.mel_filterbank(40), torch.from_numpy(g["zmuv_mean"]), torch.from_numpy(g["zmuv_mean2"]))
    with torch.no_grad():
        np.testing.assert_allclose(O.mobilenet_forward(x, sd, train=False).numpy(), g["logits_eval"], rtol=1e-3, atol=1e-3)
        bf = O.mobilenet_forward(x, sd, train=True, bf16=True)
    loss, logits, grads = O.mobilenet_grads(x, labels, sd)
    np.testing.assert_allclose(logits.numpy(), g["logits_train"], rtol=1e-3, atol=1e-3)
    np.testing.assert_allclose(loss.item(), float(g["loss_train"]), rtol=1e-4)
    names = O.mobilenet_param_names(sd)
    # The gradient of this network is ill conditioned (ReLU6 decisions within rounding of a threshold flip with the summation order, which
    # changes with the host's thread count / CPU): per-tensor norms to 3e-2, the element sample in rel-L2 to 1e-2 with >= 99 % of the
    # elements inside the tight per-element bar.
    scale = 1e-4 * float(g["grad_norms"].max())
    np.testing.assert_allclose([float(grads[k].norm()) for k in names], g["grad_norms"], rtol=3e-2, atol=scale)
    got, ref = torch.cat([grads[k].reshape(-1)[:16] for k in names]).numpy(), g["grad_sample"]
    assert np.linalg.norm(got - ref) / np.linalg.norm(ref) < 1e-2
    assert np.mean(np.abs(got - ref) <= scale + 2e-2 * np.abs(ref)) >= 0.99
    want = torch.from_numpy(g["logits_train"])
    assert ((bf - want).norm() / want.norm()).item() < 2e-2 and torch.equal(bf.argmax(1), want.argmax(1))
