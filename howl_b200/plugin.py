"""Drop-in installation into an importable castorini/howl checkout.

    import howl_b200.plugin; howl_b200.plugin.install()

replaces, inside the reference's own modules, the classes on the hot path by their CUDA-backed mirrors,
so that `python -m training.run.train` / `training.run.pretrain_gsc` / `howl.client` run unmodified on top of
libhowl_b200.so.  Nothing outside those names is touched.  See INTEGRATION.md.
"""
from __future__ import annotations

import importlib


def install(strict: bool = True):
    from . import inference, model, transform

    try:
        t = importlib.import_module("howl.data.transform.transform")
        op = importlib.import_module("howl.data.transform.operator")
        base = importlib.import_module("howl.model.base")
        cnn = importlib.import_module("howl.model.cnn")
        inf = importlib.import_module("howl.model.inference")
        pkg = importlib.import_module("howl.data.transform")
    except ImportError as exc:
        if strict:
            raise RuntimeError("howl_b200.plugin.install(): the reference package `howl` is not importable") from exc
        return False
    t.StandardAudioTransform = transform.StandardAudioTransform
    t.SpecAugmentTransform = transform.SpecAugmentTransform
    op.ZmuvTransform = transform.ZmuvTransform
    pkg.ZmuvTransform = transform.ZmuvTransform
    # the reference registry keeps name -> class; re-point "res8" and keep every other registered name
    base.RegisteredModel.registered_map["res8"] = model.Res8
    base.RegisteredModel.registered_map["lstm"] = model.SimpleLstm
    base.RegisteredModel.registered_map["seq-lstm"] = model.SequentialLstm
    cnn.Res8 = model.Res8
    inf.StandardAudioTransform = transform.StandardAudioTransform
    inf.InferenceEngine = inference.InferenceEngine
    inf.FrameInferenceEngine = inference.FrameInferenceEngine
    return True
