"""TEST INFRASTRUCTURE ONLY -- ``tests/golden/las.npz``: the reference's LASClassifier (howl/model/rnn.py:194-215) with the shipped
GSC checkpoint (``howl-models/.../commands_recognition/las/0``, 30 labels) on seeded clips: an equal-length batch and a ragged,
length-sorted one (packed BiLSTM + attention mask), eval mode.  Inputs, logits and the SHA-256 of the checkpoint's tensors are
kept; the weights (1.9 MB) are read from the mounted reference by the CPU test.

    python oracle/make_golden_las.py
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from make_golden import OUT, REF, _install_shims  # noqa: E402
from make_golden_mobilenet import state_dict_digest  # noqa: E402

CKPT = os.path.join(REF, "howl-models", "howl", "experiments", "commands_recognition", "las", "0")


def main():
    os.environ.update({"NUM_MELS": "40", "MAX_WINDOW_SIZE_SECONDS": "1", "VOCAB": '["hey","fire","fox"]', "INFERENCE_SEQUENCE": "[0,1,2]"})
    _install_shims()
    import torch

    torch.set_num_threads(1)
    from howl.data.transform.operator import ZmuvTransform
    from howl.data.transform.transform import StandardAudioTransform
    from howl.model import RegisteredModel

    model = RegisteredModel.find_registered_class("las")(30)
    sd = torch.load(os.path.join(CKPT, "model-best.pt.bin"), map_location="cpu")
    model.load_state_dict(sd)
    model.eval()
    zmuv = ZmuvTransform()
    zmuv.load_state_dict(torch.load(os.path.join(CKPT, "zmuv.pt.bin"), map_location="cpu"))
    std = StandardAudioTransform().eval()
    g = torch.Generator().manual_seed(321)
    pcm = (torch.randn(5, 16000, generator=g) * 0.1).clamp_(-1, 1)
    sample_lengths = torch.tensor([16000, 16000, 12000, 9000, 4100])       # sorted descending, as tensorize_audio_data leaves them
    for i, n in enumerate(sample_lengths.tolist()):
        pcm[i, n:] = 0
    with torch.no_grad():
        feats = zmuv(std(pcm))
        full = model(feats, None)
        lengths = std.compute_lengths(sample_lengths)
        ragged = model(feats, lengths.clone())     # LASEncoder mutates nothing, SimpleGru would (`lengths += 4`): clone anyway
    np.savez_compressed(os.path.join(OUT, "las.npz"), pcm=pcm.numpy(), feats=feats.numpy(), lengths=lengths.numpy(),
                        logits_full=full.numpy(), logits_ragged=ragged.numpy(), zmuv_mean=zmuv.mean.numpy(), zmuv_mean2=zmuv.mean2.numpy(),
                        digest=np.frombuffer(state_dict_digest(sd).encode(), dtype=np.uint8),
                        **{"sd." + k: v.numpy() for k, v in sd.items()})      # the checkpoint itself (1.9 MB): the GPU box has no /root/reference
    print("las fixture:", feats.shape, lengths.tolist(), full[0, :3], ragged[4, :3])


if __name__ == "__main__":
    main()
