"""TEST INFRASTRUCTURE ONLY -- ``tests/golden/gru.npz``: the reference's SimpleGru (howl/model/rnn.py:94-130) with a seeded random
initialisation (no checkpoint of this model ships), eval and train (batch-statistics BatchNorm, dropout off) mode, equal-length and
ragged length-sorted batches.  The state dict is small (≈0.2 MB) and is stored in the fixture.

    python oracle/make_golden_gru.py
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from make_golden import OUT, _install_shims  # noqa: E402


def main():
    os.environ.update({"NUM_MELS": "40", "MAX_WINDOW_SIZE_SECONDS": "1", "VOCAB": '["hey","fire","fox"]', "INFERENCE_SEQUENCE": "[0,1,2]"})
    _install_shims()
    import torch

    torch.set_num_threads(1)
    from howl.data.transform.transform import StandardAudioTransform
    from howl.model import RegisteredModel

    torch.manual_seed(17)
    model = RegisteredModel.find_registered_class("gru")(12)
    with torch.no_grad():                       # make the BatchNorm running statistics non-trivial
        for m in model.modules():
            if isinstance(m, torch.nn.BatchNorm2d):
                m.running_mean.uniform_(-0.5, 0.5)
                m.running_var.uniform_(0.5, 2.0)
    std = StandardAudioTransform().eval()
    g = torch.Generator().manual_seed(55)
    pcm = (torch.randn(4, 16000, generator=g) * 0.1).clamp_(-1, 1)
    sample_lengths = torch.tensor([16000, 14000, 9000, 5000])
    for i, n in enumerate(sample_lengths.tolist()):
        pcm[i, n:] = 0
    out = {"sd." + k: v.detach().numpy().copy() for k, v in model.state_dict().items()}
    with torch.no_grad():
        feats = std(pcm)
        lengths = std.compute_lengths(sample_lengths)
        model.eval()
        out["logits_eval_full"] = model(feats, None).numpy()
        out["logits_eval_ragged"] = model(feats, lengths.clone()).numpy()
        model.train()
        model.dnn[2].p = 0.0
        out["logits_train_ragged"] = model(feats, lengths.clone()).numpy()
    out.update(pcm=pcm.numpy(), feats=feats.numpy(), lengths=lengths.numpy())
    np.savez_compressed(os.path.join(OUT, "gru.npz"), **out)
    print("gru fixture:", feats.shape, lengths.tolist(), out["logits_eval_ragged"][3, :3])


if __name__ == "__main__":
    main()
