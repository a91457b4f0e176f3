"""deepfake_vit_b200 -- B200 (sm_100a) implementation of the Deepfake-ViT reference's hot path.

A drop-in for `src.feature_extraction.DeepfakeDetectionModel` and
`src.training.losses.CombinedLoss` of Ji-Hyeon212/Deepfake-ViT: same constructors, same
forward signatures, same state_dict layout; all arithmetic runs in hand-written CUDA kernels
(libdfvit.so, C ABI in include/dfvit.h).  No CPU path, no stock-PyTorch compute path.
"""
from . import _lib, ops, parallel  # noqa: F401  (raises if libdfvit.so is missing)
from .losses import CombinedLoss
from .optim import FusedAdamW
from .graphed import GraphedInference, GraphedTrainStep
from .model import (ChannelAttention, DeepfakeDetectionModel, DeepfakeFeatureExtractor, EfficientNetB4Backbone,
                    HybridAttention, LandmarkAttention, SpatialAttention)

# The `model:` section of the reference's config/model_config.yaml:4-19 as constructor kwargs
# (`pretrained` is moot offline: the ImageNet file is absent and the reference falls back to
# random init, efficientnet.py:48-54).
DEFAULT_MODEL_CONFIG = {
    "num_classes": 2,
    "pretrained": False,
    "feature_extractor_config": {
        "pretrained": False, "freeze_bn": False, "dropout_rate": 0.4, "use_attention": True,
        "attention_config": {"use_landmark": True, "use_spatial": True, "use_channel": True},
    },
    "classifier_hidden_dims": [512, 128, 32],
    "dropout_rate": 0.4,
}

__all__ = ["FusedAdamW", "GraphedInference", "GraphedTrainStep", "DEFAULT_MODEL_CONFIG", "DeepfakeDetectionModel", "DeepfakeFeatureExtractor", "EfficientNetB4Backbone", "HybridAttention",
           "LandmarkAttention", "SpatialAttention", "ChannelAttention", "CombinedLoss", "ops"]
