"""GPU: lstm / seq-lstm (howl/model/rnn.py:41-91) through the C ABI and the nn.Module mirrors, against the committed
reference outputs (real GSC checkpoint, ragged lengths, streaming state, two reference train steps) and the oracle."""
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


def _sd(g, prefix):
    return {k[len(prefix):]: torch.from_numpy(v) for k, v in g.items() if k.startswith(prefix)}


def _feats(g, key="pcm"):
    return O.hot_path_features(torch.from_numpy(g[key]), O.mel_filterbank(40), torch.from_numpy(g["zmuv.mean"]),
                               torch.from_numpy(g["zmuv.mean2"]))


def test_registry_and_state_dict_keys(golden):
    from howl_b200.model import RegisteredModel

    g = golden("lstm")
    for name in ("lstm", "seq-lstm"):
        m = RegisteredModel.find_registered_class(name)(30)
        assert list(m.state_dict().keys()) == [k[3:] for k in g if k.startswith("sd.")]
        assert sum(p.numel() for p in m.parameters()) == 127774      # lstm @30 labels


def test_lstm_real_weights_ragged_and_streaming(golden):
    from howl_b200.model import RegisteredModel

    g = golden("lstm")
    sd = _sd(g, "sd.")
    feats = _feats(g).to(DEV)
    lstm = RegisteredModel.find_registered_class("lstm")(30)
    lstm.load_state_dict(sd)
    lstm = lstm.to(DEV).eval()
    with torch.no_grad():
        out = lstm(feats, torch.from_numpy(g["lengths"]))
        np.testing.assert_allclose(out.cpu().numpy(), g["logits"], rtol=RTOL, atol=ATOL)
        assert np.array_equal(out.argmax(1).cpu().numpy(), g["logits"].argmax(1))
        out = lstm(feats, torch.from_numpy(g["ragged_lengths"]))
        np.testing.assert_allclose(out.cpu().numpy(), g["ragged_logits"], rtol=RTOL, atol=ATOL)
        seq = RegisteredModel.find_registered_class("seq-lstm")(30)
        seq.load_state_dict(sd)
        seq = seq.to(DEV).eval().streaming()
        o1 = seq(feats, torch.from_numpy(g["lengths"]))
        np.testing.assert_allclose(o1.cpu().numpy(), g["seq_out1"], rtol=RTOL, atol=ATOL)
        h, c = seq.streaming_state
        np.testing.assert_allclose(h.cpu().numpy(), g["seq_h1"], rtol=RTOL, atol=ATOL)
        np.testing.assert_allclose(c.cpu().numpy(), g["seq_c1"], rtol=RTOL, atol=ATOL)
        o2 = seq(feats, torch.from_numpy(g["lengths"]))
        np.testing.assert_allclose(o2.cpu().numpy(), g["seq_out2"], rtol=RTOL, atol=ATOL)


def test_lstm_module_reference_train_steps(golden):
    """training/run/train.py:292-302 with --model lstm: torch CrossEntropyLoss + torch AdamW over model.parameters()."""
    from howl_b200.model import RegisteredModel

    g = golden("lstm")
    L = 12
    model = RegisteredModel.find_registered_class("lstm")(L)
    model.load_state_dict(_sd(g, "init."))
    model = model.to(DEV).train()
    opt = torch.optim.AdamW(model.parameters(), 0.01, weight_decay=1e-5)
    feats = _feats(g, "t_pcm").to(DEV)
    labels, lengths = torch.from_numpy(g["t_labels"]).to(DEV), torch.from_numpy(g["t_lengths"])
    for step in (1, 2):
        scores = model(feats, lengths)
        loss = torch.nn.functional.cross_entropy(scores, labels)
        opt.zero_grad()
        loss.backward()
        np.testing.assert_allclose(loss.item(), g[f"step{step}.loss"], rtol=RTOL, atol=ATOL)
        np.testing.assert_allclose(scores.detach().cpu().numpy(), g[f"step{step}.logits"], rtol=RTOL, atol=ATOL)
        for k, p in model.named_parameters():
            want = g[f"step{step}.grad.{k}"]
            np.testing.assert_allclose(p.grad.cpu().numpy(), want, rtol=2e-3, atol=1e-4 * np.abs(want).max(), err_msg=k)
        opt.step()
        model.load_state_dict({k: torch.from_numpy(g[f"step{step}.sd.{k}"]) for k in model.state_dict()})


@pytest.mark.parametrize("B,T", [(1, 8000), (33, 8000), (70, 16000), (2048, 8000)])   # last: BASELINE config 4's batch
def test_lstm_fused_train_step_vs_oracle(B, T):
    import howl_b200

    L = 5
    ctx = howl_b200.Context(DEV, n_mels=40)
    pcm, labels = O.synthetic_batch(B, T, L, seed=B)
    fb = O.mel_filterbank(40)
    zmean, zstd = -1.78896, 3.93389
    F = O.num_frames(T)
    full = int(O.compute_lengths([T])[0])
    rng = np.random.default_rng(B)
    lengths = torch.from_numpy(np.sort(rng.integers(1, full + 1, size=B))[::-1].copy()) if B > 1 else torch.tensor([full])
    lengths[0] = full
    params = O.lstm_init(L, seed=3)
    flat = O.lstm_flatten(params, L).to(DEV)
    assert flat.numel() == ctx.lstm_param_count(L)
    grads, m, v = torch.zeros_like(flat), torch.zeros_like(flat), torch.zeros_like(flat)
    loss, logits = torch.zeros(1, device=DEV), torch.zeros(B, L, device=DEV)
    ws = torch.empty(ctx.lstm_train_step_workspace_bytes(B, T, full, L), dtype=torch.uint8, device=DEV)
    ctx.lstm_train_step(pcm.to(DEV), labels.to(DEV), lengths.to(DEV), full, fb.to(DEV), (zmean, zstd), flat, grads, m, v, 1,
                        0.01, 1e-5, loss, logits, ws)
    feats = O.hot_path_features(pcm, fb, torch.tensor([zmean]), torch.tensor([zmean ** 2 + zstd ** 2]))
    om = {k: torch.zeros_like(p) for k, p in params.items()}
    ov = {k: torch.zeros_like(p) for k, p in params.items()}
    oloss, ologits, ograds = O.lstm_train_step(feats, labels, lengths, params, om, ov, 1, 0.01, 1e-5)
    np.testing.assert_allclose(logits.cpu().numpy(), ologits.numpy(), rtol=RTOL, atol=ATOL)
    np.testing.assert_allclose(loss.item(), oloss.item(), rtol=RTOL, atol=ATOL)
    got = O.lstm_unflatten(grads.cpu(), L)
    for k in got:
        want = ograds[k].numpy()
        np.testing.assert_allclose(got[k].numpy(), want, rtol=2e-3, atol=2e-4 * np.abs(want).max(), err_msg=k)
    np.testing.assert_allclose(flat.cpu().numpy(), O.lstm_flatten(params, L).numpy(), rtol=1e-3, atol=5e-4)
    ctx.close()


@pytest.mark.parametrize("B,T,M", [(37, 8000, 40), (530, 8000, 40), (21, 4600, 80), (19, 8000, 44)])   # 44 mels: K = 172 = 21 * 8 + 4 (k tail)
def test_lstm_pipelined_recurrences_match_the_plain_ones(B, T, M):
    """Option "lstm_engine": the software-pipelined forward / backward recurrences (1: same thread mapping; 2: 2 x 8 register tile) move
    data earlier / re-tile the product but keep the summation order per output of the plain kernels (0).  The forward is deterministic end to end, so the logits must agree BIT FOR BIT; the
    weight-gradient GEMMs and the loss accumulate with atomics (order varies run to run), so gradients are held to 1e-5 of the
    tensor's scale -- on ragged lengths, for batches that are not a multiple of the 16-sequence CTA tile, with a k tail."""
    import howl_b200

    L = 7
    pcm, labels = O.synthetic_batch(B, T, L, seed=B + 1)
    fb = O.mel_filterbank(M)
    full = int(O.compute_lengths([T])[0])
    rng = np.random.default_rng(B)
    lengths = torch.from_numpy(rng.integers(1, full + 1, size=B))
    lengths[B // 2] = full
    got = {}
    for engine in (0, 1, 2):
        ctx = howl_b200.Context(DEV, n_mels=M)
        ctx.set_option("lstm_engine", engine)
        gen = torch.Generator().manual_seed(5)
        flat = (torch.randn(ctx.lstm_param_count(L), generator=gen) * 0.08).to(DEV)
        grads, m, v = torch.zeros_like(flat), torch.zeros_like(flat), torch.zeros_like(flat)
        loss, logits = torch.zeros(1, device=DEV), torch.zeros(B, L, device=DEV)
        ws = torch.empty(ctx.lstm_train_step_workspace_bytes(B, T, full, L), dtype=torch.uint8, device=DEV)
        ctx.lstm_train_step(pcm.to(DEV), labels.to(DEV), lengths.to(DEV), full, fb.to(DEV), (-1.8, 3.9), flat, grads, m, v, 1,
                            0.01, 1e-5, loss, logits, ws)
        torch.cuda.synchronize()
        got[engine] = (loss.clone(), logits.clone(), grads.clone())
        ctx.close()
    l0, z0, g0 = got[0]
    for engine in (1, 2):
        l1, z1, g1 = got[engine]
        assert torch.isfinite(z1).all() and torch.isfinite(g1).all() and g1.abs().max() > 0
        assert torch.equal(z0, z1), f"engine {engine} logits: max |diff| = {(z0 - z1).abs().max().item():.3e}"
        np.testing.assert_allclose(l1.item(), l0.item(), rtol=1e-6)
        # per parameter tensor of nn.LSTM(M -> 128) + Linear(128 -> 256) + Linear(256 -> L), flat in state_dict order
        off = 0
        for name, n in (("w_ih", 512 * M), ("w_hh", 512 * 128), ("b_ih", 512), ("b_hh", 512), ("w1", 256 * 128), ("b1", 256), ("w2", L * 256), ("b2", L)):
            a, b = g0[off:off + n], g1[off:off + n]
            assert (a - b).abs().max().item() <= 1e-5 * a.abs().max().item() + 1e-12, (engine, name)
            off += n
        assert off == g0.numel()


def test_seq_lstm_ctc_reference_steps_module_and_fused(golden):
    """The CTC branch of training/run/train.py:294-302 with the streaming seq-lstm, (a) through the nn.Module mirror with
    torch's log_softmax + nn.CTCLoss + AdamW, (b) through the fused C-ABI step with the library's own CTC kernel."""
    import howl_b200
    from howl_b200.model import RegisteredModel

    g = golden("lstm")
    L, blank = 5, 4
    feats_cpu = _feats(g, "t_pcm")
    lengths = torch.from_numpy(g["t_lengths"])
    tg, tl = torch.from_numpy(g["ctc.targets"]), torch.from_numpy(g["ctc.target_lengths"])
    # ---- (a) module
    model = RegisteredModel.find_registered_class("seq-lstm")(L)
    model.load_state_dict(_sd(g, "ctc.init."))
    model = model.to(DEV).train().streaming()
    opt = torch.optim.AdamW(model.parameters(), 0.01, weight_decay=1e-5)
    ctc = torch.nn.CTCLoss(blank)
    for step in (1, 2):
        scores = model(feats_cpu.to(DEV), lengths)
        loss = ctc(torch.nn.functional.log_softmax(scores, -1), tg.to(DEV), lengths.to(DEV), tl.to(DEV))
        opt.zero_grad()
        loss.backward()
        np.testing.assert_allclose(loss.item(), g[f"ctc.step{step}.loss"], rtol=RTOL, atol=ATOL)
        np.testing.assert_allclose(scores.detach().cpu().numpy(), g[f"ctc.step{step}.scores"], rtol=RTOL, atol=ATOL)
        np.testing.assert_allclose(model.streaming_state[0].cpu().numpy(), g[f"ctc.step{step}.h"], rtol=RTOL, atol=ATOL)
        for k, p in model.named_parameters():
            want = g[f"ctc.step{step}.grad.{k}"]
            np.testing.assert_allclose(p.grad.cpu().numpy(), want, rtol=2e-3, atol=2e-4 * np.abs(want).max(), err_msg=k)
        opt.step()
        model.load_state_dict({k: torch.from_numpy(g[f"ctc.step{step}.sd.{k}"]) for k in model.state_dict()})
    # ---- (b) fused step, own CTC kernel
    ctx = howl_b200.Context(DEV, n_mels=40)
    pcm = torch.from_numpy(g["t_pcm"]).to(DEV)
    B, T = pcm.shape
    steps = int(lengths.max())
    init = _sd(g, "ctc.init.")
    flat = O.lstm_flatten({k: init[k] for k, _ in O.lstm_param_shapes(L)}, L).to(DEV)
    grads, m, v = torch.zeros_like(flat), torch.zeros_like(flat), torch.zeros_like(flat)
    state = torch.zeros(2, B, 128, device=DEV)
    loss, scores = torch.zeros(1, device=DEV), torch.zeros(steps, B, L, device=DEV)
    ws = torch.empty(ctx.seq_lstm_train_step_workspace_bytes(B, T, steps, L), dtype=torch.uint8, device=DEV)
    mean = float(g["zmuv.mean"][0])
    std = float(np.sqrt(g["zmuv.mean2"][0] - g["zmuv.mean"][0] ** 2))
    for step in (1, 2):
        ctx.seq_lstm_ctc_train_step(pcm, tg.to(DEV), tl.to(DEV), blank, lengths.to(DEV), steps, O.mel_filterbank(40).to(DEV),
                                    (mean, std), flat, state, grads, m, v, step, 0.01, 1e-5, loss, scores, ws)
        np.testing.assert_allclose(loss.item(), g[f"ctc.step{step}.loss"], rtol=RTOL, atol=ATOL)
        np.testing.assert_allclose(scores.cpu().numpy(), g[f"ctc.step{step}.scores"], rtol=RTOL, atol=ATOL)
        np.testing.assert_allclose(state[0].cpu().numpy(), g[f"ctc.step{step}.h"][0], rtol=RTOL, atol=ATOL)
        np.testing.assert_allclose(state[1].cpu().numpy(), g[f"ctc.step{step}.c"][0], rtol=RTOL, atol=ATOL)
        got = O.lstm_unflatten(grads.cpu(), L)
        for k in got:
            want = g[f"ctc.step{step}.grad.{k}"]
            np.testing.assert_allclose(got[k].numpy(), want, rtol=2e-3, atol=2e-4 * np.abs(want).max(), err_msg=k)
        flat.copy_(O.lstm_flatten({k: torch.from_numpy(g[f"ctc.step{step}.sd.{k}"]) for k, _ in O.lstm_param_shapes(L)}, L).to(DEV))
    ctx.close()


@pytest.mark.parametrize("B,T", [(3, 8000), (40, 16000), (2048, 8000)])   # last: BASELINE config 4 (seq-lstm + CTC, batch 2048)
def test_ctc_kernel_vs_torch_ragged(B, T):
    """Own CTC kernel against torch's F.ctc_loss on ragged input / target lengths (incl. repeated labels, length-1 targets)."""
    import howl_b200

    L, blank = 6, 5
    ctx = howl_b200.Context(DEV, n_mels=40)
    pcm, _ = O.synthetic_batch(B, T, L, seed=B + 1)
    fb = O.mel_filterbank(40)
    zmean, zstd = -1.78896, 3.93389
    full = int(O.compute_lengths([T])[0])
    rng = np.random.default_rng(B)
    lengths = torch.from_numpy(np.sort(rng.integers(max(8, full // 2), full + 1, size=B))[::-1].copy())
    lengths[0] = full
    tl = torch.from_numpy(rng.integers(1, 7, size=B))
    tg = torch.from_numpy(rng.integers(0, blank, size=(B, 6)))
    params = O.lstm_init(L, seed=9)
    flat = O.lstm_flatten(params, L).to(DEV)
    feats_cpu = O.hot_path_features(pcm, fb, torch.tensor([zmean]), torch.tensor([zmean ** 2 + zstd ** 2]))
    feats = ctx.frontend(pcm.to(DEV), fb.to(DEV), "time_major", zmuv=(zmean, zstd))
    ws = torch.empty(ctx.lstm_workspace_bytes(B, full, L, True, True), dtype=torch.uint8, device=DEV)
    scores = ctx.lstm_fwd(feats, lengths.to(DEV), full, flat, ws, sequential=True, train=True)
    grads, loss = torch.zeros_like(flat), torch.zeros(1, device=DEV)
    ctx.lstm_ctc_bwd(tuple(feats.shape), lengths.to(DEV), full, tg.to(DEV), tl.to(DEV), blank, flat, grads, loss, ws)
    m = {k: torch.zeros_like(p) for k, p in params.items()}
    v = {k: torch.zeros_like(p) for k, p in params.items()}
    oloss, oscores, ograds, _ = O.seq_lstm_ctc_step(feats_cpu, tg, tl, lengths, params, None, blank, m, v, 1, 0.01, 0.0)
    np.testing.assert_allclose(scores.cpu().numpy(), oscores.numpy(), rtol=RTOL, atol=ATOL)
    np.testing.assert_allclose(loss.item(), oloss.item(), rtol=RTOL, atol=ATOL)
    got = O.lstm_unflatten(grads.cpu(), L)
    for k in got:
        want = ograds[k].numpy()
        np.testing.assert_allclose(got[k].numpy(), want, rtol=2e-3, atol=2e-4 * np.abs(want).max(), err_msg=k)
    ctx.close()
