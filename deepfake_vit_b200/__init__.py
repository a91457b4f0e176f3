"""deepfake_vit_b200 -- B200 (sm_100a) implementation of the Deepfake-ViT reference's hot path.

A drop-in for `src.feature_extraction.DeepfakeDetectionModel` and
`src.training.losses.CombinedLoss` of Ji-Hyeon212/Deepfake-ViT: same constructors, same
forward signatures, same state_dict layout; all arithmetic runs in hand-written CUDA kernels
(libdfvit.so, C ABI in include/dfvit.h).  No CPU path, no stock-PyTorch compute path.
"""
from . import _lib, ops  # noqa: F401  (raises if libdfvit.so is missing)
from .losses import CombinedLoss
from .model import (ChannelAttention, DeepfakeDetectionModel, DeepfakeFeatureExtractor, EfficientNetB4Backbone,
                    HybridAttention, LandmarkAttention, SpatialAttention)

__all__ = ["DeepfakeDetectionModel", "DeepfakeFeatureExtractor", "EfficientNetB4Backbone", "HybridAttention",
           "LandmarkAttention", "SpatialAttention", "ChannelAttention", "CombinedLoss", "ops"]
