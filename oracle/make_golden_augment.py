"""TEST INFRASTRUCTURE ONLY -- ``tests/golden/augment.npz``: the reference's training augmentation chain on seeded clips,
    compose(DatasetMixer(noise).train(), TimeshiftTransform().train(), WakeWordFrameBatchifier)        (training/run/train.py:202-221)
driven by the global ``random`` (NoiseTransform is left out of the FIXTURE: its samples come from torch's host generator, which the
device kernel does not replay; its strength / probability draws are pinned by the statistics tests instead).

    python oracle/make_golden_augment.py
"""
import json
import os
import random
import sys
import types

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from make_golden import OUT, _install_shims  # noqa: E402


def main():
    os.environ.update({"NUM_MELS": "40", "MAX_WINDOW_SIZE_SECONDS": "0.5", "VOCAB": '["hey","fire","fox"]', "INFERENCE_SEQUENCE": "[0,1,2]"})
    _install_shims()
    import torch
    from howl.data.common.example import WakeWordClipExample
    from howl.data.common.label import FrameLabelData
    from howl.data.common.metadata import AudioClipMetadata
    from howl.data.transform.batchifier import WakeWordFrameBatchifier
    from howl.data.transform.operator import compose
    from howl.data.transform.transform import DatasetMixer, TimeshiftTransform

    g = dict(np.load(os.path.join(OUT, "batchifier.npz")))
    maps = json.load(open(os.path.join(OUT, "meta.json")))["batchifier_maps"]
    offs = np.concatenate([[0], np.cumsum(g["lengths"])])
    exs = [WakeWordClipExample(FrameLabelData({float(k): v for k, v in maps[i]}, [], []),
                               AudioClipMetadata(path=".", phone_strings=None, words=None, phone_end_timestamps=None, end_timestamps=None,
                                                 transcription=""),
                               torch.from_numpy(g["clips"][offs[i]:offs[i + 1]].copy()), 16000) for i in range(len(maps))]
    rng = np.random.default_rng(5)
    bg_lens = [40000, 31000, 52000]
    bg = [rng.standard_normal(n).astype(np.float32) * 0.05 for n in bg_lens]
    noise_ds = [types.SimpleNamespace(audio_data=torch.from_numpy(b)) for b in bg]
    out = {"bg": np.concatenate(bg), "bg_lengths": np.array(bg_lens)}
    for trial in range(8):
        random.seed(500 + trial)
        chain = compose(DatasetMixer(noise_ds).train(), TimeshiftTransform().train(),
                        WakeWordFrameBatchifier(3, positive_sample_prob=[0.5, 0.9, 0.1][trial % 3]))
        batch = chain(exs * 2)
        out[f"t{trial}.audio"], out[f"t{trial}.labels"], out[f"t{trial}.lengths"] = batch.audio_data.numpy(), batch.labels.numpy(), batch.lengths.numpy()
        out[f"t{trial}.next_draw"] = np.float64(random.random())          # the stream position after the batch
    np.savez_compressed(os.path.join(OUT, "augment.npz"), **out)
    print("augment fixture:", os.path.getsize(os.path.join(OUT, "augment.npz")) / 1e6, "MB")


if __name__ == "__main__":
    main()
