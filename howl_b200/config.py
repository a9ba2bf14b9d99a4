"""JSON training configuration with the field names, defaults and nesting of the reference's schema (``howl/config.py:9-93``), so a
``test_training_config.json`` written for howl parses unchanged: ``TrainingConfig.parse_file(path)``.

The schema is kept as ONE table (section -> field -> (type, default)) and the pydantic models are generated from it; sections that
other sections embed are listed first.  pydantic's v1 API is used on purpose: the reference's ``TrainingConfig`` has a field called
``model_config``, which pydantic v2 reserves.
"""
from typing import List

try:                                    # pydantic >= 2 ships the v1 API as a sub-package
    from pydantic.v1 import BaseModel, create_model
except ImportError:                     # pragma: no cover - pydantic 1.x
    from pydantic import BaseModel, create_model

# "@Name" = an embedded section (default: that section with its own defaults); "[@Name]" = a list of sections (default: empty)
_SCHEMA = {
    "CacheConfig": {"cache_size": (int, 128144)},
    "AudioConfig": {"sample_rate": (int, 16000), "use_mono": (bool, True)},
    "ContextConfig": {
        "seed": (int, 0), "vocab": (List[str], None), "sequence": (List[int], None),
        "token_type": (str, "word"),            # "word" | "phone" (phone-level contexts need the reference's pronunciation tooling)
        "phone_dictionary_path": (str, None),
    },
    "InferenceEngineConfig": {
        "per_frame": (bool, False), "inference_weights": (List[float], None), "inference_window_ms": (float, 2000),
        "smoothing_window_ms": (float, 50), "tolerance_window_ms": (float, 500), "inference_threshold": (float, 0),
    },
    "AudioTransformConfig": {"num_fft": (int, 512), "num_mels": (int, 40), "hop_length": (int, 200), "use_meyda_spectrogram": (bool, False)},
    "DatasetConfig": {"path": (str, None), "audio_config": "@AudioConfig", "audio_transform_config": "@AudioTransformConfig"},
    "ModelConfig": {"architecture": (str, "res8")},
    "TrainingConfig": {
        "batch_size": (int, 16), "learning_rate": (float, 0.01), "num_epochs": (int, 10), "lr_decay": (float, 0.955),
        "weight_decay": (float, 0.00001), "use_noise_dataset": (bool, False),
        "noise_datasets": "[@DatasetConfig]", "train_datasets": "[@DatasetConfig]", "val_datasets": "[@DatasetConfig]",
        "test_datasets": "[@DatasetConfig]",
        "inference_engine_config": "@InferenceEngineConfig", "cache_config": "@CacheConfig", "model_config": "@ModelConfig",
        "context_config": "@ContextConfig", "workspace_path": (str, None),
    },
}


def _build(schema):
    models = {}
    for section, fields in schema.items():
        spec = {}
        for name, desc in fields.items():
            if isinstance(desc, str) and desc.startswith("[@"):
                spec[name] = (List[models[desc[2:-1]]], [])
            elif isinstance(desc, str):
                spec[name] = (models[desc[1:]], models[desc[1:]]())
            else:
                spec[name] = desc
        models[section] = create_model(section, __base__=BaseModel, __module__=__name__, **spec)
    return models


_MODELS = _build(_SCHEMA)
CacheConfig = _MODELS["CacheConfig"]
AudioConfig = _MODELS["AudioConfig"]
ContextConfig = _MODELS["ContextConfig"]
InferenceEngineConfig = _MODELS["InferenceEngineConfig"]
AudioTransformConfig = _MODELS["AudioTransformConfig"]
DatasetConfig = _MODELS["DatasetConfig"]
ModelConfig = _MODELS["ModelConfig"]
TrainingConfig = _MODELS["TrainingConfig"]
