"""TEST INFRASTRUCTURE ONLY -- ``tests/golden/mobilenet.npz``: the reference's MobileNetClassifier (howl/model/cnn.py:15-29) with
the shipped GSC checkpoint (``howl-models/.../commands_recognition/mobilenet/0``, 30 labels) on seeded clips, eval mode, plus the
train-mode (batch-statistics) logits of the same weights.  torchvision's ``mobilenet_v2(pretrained=True)`` would download ImageNet
weights (no network here): it is called with ``pretrained=False`` -- every weight is overwritten by the checkpoint anyway.

The fixture keeps the inputs, the logits and the SHA-256 of the checkpoint's tensors, not the 9 MB of weights: the CPU test that
uses it runs where ``/root/reference`` is mounted and skips elsewhere; GPU-box tests of this model will get their weights from the
seeded initialiser.

    python oracle/make_golden_mobilenet.py
"""
import hashlib
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from make_golden import OUT, REF, _install_shims  # noqa: E402

CKPT = os.path.join(REF, "howl-models", "howl", "experiments", "commands_recognition", "mobilenet", "0")


def state_dict_digest(sd) -> str:
    h = hashlib.sha256()
    for k in sd:
        h.update(k.encode())
        h.update(np.ascontiguousarray(sd[k].detach().cpu().numpy()).tobytes())
    return h.hexdigest()


def main():
    os.environ.update({"NUM_MELS": "40", "MAX_WINDOW_SIZE_SECONDS": "1", "VOCAB": '["hey","fire","fox"]', "INFERENCE_SEQUENCE": "[0,1,2]"})
    _install_shims()
    import torch
    import torchvision.models

    torch.set_num_threads(1)
    real = torchvision.models.mobilenet_v2
    import howl.model.cnn as cnn

    cnn.mobilenet_v2 = lambda pretrained=True, **k: real(weights=None, **k)      # no network: see the module docstring
    from howl.data.transform.operator import ZmuvTransform
    from howl.data.transform.transform import StandardAudioTransform
    from howl.model import RegisteredModel

    model = RegisteredModel.find_registered_class("mobilenet")(30)
    sd = torch.load(os.path.join(CKPT, "model-best.pt.bin"), map_location="cpu")
    model.load_state_dict(sd)
    zmuv = ZmuvTransform()
    zmuv.load_state_dict(torch.load(os.path.join(CKPT, "zmuv.pt.bin"), map_location="cpu"))
    std = StandardAudioTransform().eval()
    g = torch.Generator().manual_seed(123)
    pcm = (torch.randn(4, 16000, generator=g) * 0.1).clamp_(-1, 1)
    with torch.no_grad():
        feats = zmuv(std(pcm))
        model.eval()
        logits_eval = model(feats, None)
        model.train()
        model.model.classifier[0].p = 0.0                # dropout off: the train-mode fixture pins the batch-statistics BatchNorm path
        logits_train = model(feats, None)
    np.savez_compressed(os.path.join(OUT, "mobilenet.npz"), pcm=pcm.numpy(), feats=feats.numpy(), logits_eval=logits_eval.numpy(),
                        logits_train=logits_train.numpy(), zmuv_mean=zmuv.mean.numpy(), zmuv_mean2=zmuv.mean2.numpy(),
                        digest=np.frombuffer(state_dict_digest(sd).encode(), dtype=np.uint8))
    print("mobilenet fixture:", feats.shape, logits_eval[0, :4], state_dict_digest(sd)[:16])

    # ---- the checkpoint itself, for the GPU box (where /root/reference does not exist).  A randomly initialised MobileNetV2 with
    # batch-statistics BatchNorm is chaotic (perturbations grow exponentially with depth), so bf16 parity can only be judged on trained
    # weights.  The GEMM convolutions' weights are stored as bf16 bit patterns (what the tensor-core path feeds the MMAs anyway), the
    # rest as fp32: 4.6 MB.  The logits below come from the REFERENCE module loaded with exactly these (dequantised) weights, on a
    # larger seeded batch, in eval and in train (batch statistics, dropout off) mode, plus the reference's own autograd gradients.
    q = {}
    sd_q = {}
    for k, v in sd.items():
        if v.dim() == 4 and (v.shape[2] == 1 or k == "model.features.0.0.weight") and v.shape[1] > 1:
            bits = v.to(torch.bfloat16).view(torch.int16).numpy().view(np.uint16)
            q[k + "::bf16"] = bits
            sd_q[k] = v.to(torch.bfloat16).to(torch.float32)
        else:
            q[k] = v.numpy()
            sd_q[k] = v.clone()
    model.load_state_dict(sd_q)
    g2 = torch.Generator().manual_seed(321)
    pcm2 = (torch.randn(24, 16000, generator=g2) * 0.1).clamp_(-1, 1)
    labels2 = torch.randint(0, 30, (24,), generator=g2)
    with torch.no_grad():
        feats2 = zmuv(std(pcm2))
        model.eval()
        q["logits_eval"] = model(feats2, None).numpy()
    model.train()
    model.model.classifier[0].p = 0.0
    model.zero_grad()
    out = model(feats2, None)
    loss = torch.nn.functional.cross_entropy(out, labels2)
    loss.backward()
    q["logits_train"], q["loss_train"] = out.detach().numpy(), np.float32(loss.item())
    # (the 9 MB of reference gradients are not shipped: tests/test_oracle_golden.py checks, where the reference is mounted, that the
    # oracle's autograd reproduces them; the GPU box compares against the oracle)
    gnorm = {k: float(p_.grad.norm()) for k, p_ in model.named_parameters()}
    q["grad_norms"] = np.array([gnorm[k] for k, _ in model.named_parameters()], dtype=np.float32)
    q["grad_sample"] = torch.cat([p_.grad.reshape(-1)[:16] for _, p_ in model.named_parameters()]).numpy()
    q["pcm"], q["labels"] = pcm2.numpy(), labels2.numpy()
    q["zmuv_mean"], q["zmuv_mean2"] = zmuv.mean.numpy(), zmuv.mean2.numpy()
    np.savez_compressed(os.path.join(OUT, "mobilenet_ckpt.npz"), **q)
    print("mobilenet checkpoint fixture:", os.path.getsize(os.path.join(OUT, "mobilenet_ckpt.npz")) / 1e6, "MB, train loss", loss.item())


if __name__ == "__main__":
    main()
