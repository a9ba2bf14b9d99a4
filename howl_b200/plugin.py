"""Drop-in installation into an importable castorini/howl checkout.

    import howl_b200.plugin; howl_b200.plugin.install()

replaces, inside the reference's own modules, the classes on the hot path by their CUDA-backed counterparts, so that
`python -m training.run.train` / `training.run.pretrain_gsc` / `howl.client` run unmodified on top of libhowl_b200.so:

  * howl.data.transform.transform.{StandardAudioTransform, SpecAugmentTransform}, howl.data.transform.operator.ZmuvTransform
  * the registry entries (and module attributes) of the models built here: res8, lstm, seq-lstm, mobilenet, las
  * howl.settings.SETTINGS becomes the single source of settings for the substituted classes (``howl_b200.settings.SETTINGS.bind``)

The reference's InferenceEngine / FrameInferenceEngine are NOT replaced: they run unmodified on the substituted transforms and
models; ``FrameInferenceEngine.infer_batched`` (all windows of a clip as one device batch) is added as an extra method.
Nothing outside those names is touched.  See INTEGRATION.md.
"""
from __future__ import annotations

import importlib


def install(strict: bool = True):
    from . import inference, model, settings, transform

    try:
        ref_settings = importlib.import_module("howl.settings")
        t = importlib.import_module("howl.data.transform.transform")
        op = importlib.import_module("howl.data.transform.operator")
        base = importlib.import_module("howl.model.base")
        cnn = importlib.import_module("howl.model.cnn")
        rnn = importlib.import_module("howl.model.rnn")
        inf = importlib.import_module("howl.model.inference")
        pkg = importlib.import_module("howl.data.transform")
    except ImportError as exc:
        if strict:
            raise RuntimeError("howl_b200.plugin.install(): the reference package `howl` is not importable") from exc
        return False
    settings.SETTINGS.bind(ref_settings.SETTINGS)
    t.StandardAudioTransform = transform.StandardAudioTransform
    t.SpecAugmentTransform = transform.SpecAugmentTransform
    op.ZmuvTransform = transform.ZmuvTransform
    pkg.ZmuvTransform = transform.ZmuvTransform
    inf.StandardAudioTransform = transform.StandardAudioTransform     # the reference engines build their own transform
    # the reference registry keeps name -> class; re-point the models built here and keep every other registered name
    for name, cls in model.ACCELERATED.items():
        base.RegisteredModel.registered_map[name] = cls
    cnn.Res8 = model.Res8
    rnn.SimpleLstm, rnn.SequentialLstm = model.SimpleLstm, model.SequentialLstm
    if "mobilenet" in model.ACCELERATED:
        cnn.MobileNetClassifier = model.ACCELERATED["mobilenet"]
    if "las" in model.ACCELERATED:
        rnn.LASClassifier = model.ACCELERATED["las"]
    inf.FrameInferenceEngine.infer_batched = inference.infer_batched
    return True
