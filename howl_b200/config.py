"""JSON training configuration, field-for-field the reference's ``howl/config.py:9-93`` (names, defaults, nesting), so a
``test_training_config.json`` written for howl parses unchanged: ``TrainingConfig.parse_file(path)``.

pydantic's v1 API is used on purpose: the reference's ``TrainingConfig`` has a field called ``model_config``, which pydantic v2
reserves.
"""
from typing import List

try:                                    # pydantic >= 2 ships the v1 API as a sub-package
    from pydantic.v1 import BaseModel
except ImportError:                     # pragma: no cover - pydantic 1.x
    from pydantic import BaseModel


class CacheConfig(BaseModel):
    cache_size: int = 128144


class AudioConfig(BaseModel):
    sample_rate: int = 16000
    use_mono: bool = True


class ContextConfig(BaseModel):
    seed: int = 0
    vocab: List[str] = None
    sequence: List[int] = None
    token_type: str = "word"            # "word" | "phone" (phone-level contexts need the reference's pronunciation tooling)
    phone_dictionary_path: str = None


class InferenceEngineConfig(BaseModel):
    per_frame: bool = False
    inference_weights: List[float] = None
    inference_window_ms: float = 2000
    smoothing_window_ms: float = 50
    tolerance_window_ms: float = 500
    inference_threshold: float = 0


class AudioTransformConfig(BaseModel):
    num_fft: int = 512
    num_mels: int = 40
    hop_length: int = 200
    use_meyda_spectrogram: bool = False


class DatasetConfig(BaseModel):
    path: str = None
    audio_config: AudioConfig = AudioConfig()
    audio_transform_config: AudioTransformConfig = AudioTransformConfig()


class ModelConfig(BaseModel):
    architecture: str = "res8"


class TrainingConfig(BaseModel):
    batch_size: int = 16
    learning_rate: float = 0.01
    num_epochs: int = 10
    lr_decay: float = 0.955
    weight_decay: float = 0.00001
    use_noise_dataset: bool = False
    noise_datasets: List[DatasetConfig] = []
    train_datasets: List[DatasetConfig] = []
    val_datasets: List[DatasetConfig] = []
    test_datasets: List[DatasetConfig] = []
    inference_engine_config: InferenceEngineConfig = InferenceEngineConfig()
    cache_config: CacheConfig = CacheConfig()
    model_config: ModelConfig = ModelConfig()
    context_config: ContextConfig = ContextConfig()
    workspace_path: str = None
