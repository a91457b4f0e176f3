"""TEST INFRASTRUCTURE ONLY.  Restatement of the reference's wrapper code.

Used as the checker wherever /root/reference is not mounted (the GPU box).  Each
class cites the reference lines it follows; tests/test_oracle.py proves, in the
build container, that every class here is bit-identical (state_dict keys, seeded
init, forward values, loss values, gradients) to the real reference import.

Module / attribute names equal the reference's so that state_dict keys match
(SURVEY.md Appendix A.6).
"""
from typing import Dict, Optional, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ensure_shim_on_path

ensure_shim_on_path()
from efficientnet_pytorch import EfficientNet  # noqa: E402  (the restated shim)


class EfficientNetB4Backbone(nn.Module):
    """src/feature_extraction/efficientnet.py:13-170."""

    def __init__(self, pretrained=True, freeze_bn=False, dropout_rate=0.4, extract_features=True):
        super().__init__()
        # efficientnet.py:40-61 -- both branches build from_name(); the ImageNet file
        # (efficientnet.py:12) is never present offline, so init is always random.
        self.backbone = EfficientNet.from_name("efficientnet-b4", num_classes=1000)
        self.extract_features = extract_features
        self.freeze_bn = freeze_bn
        if extract_features:
            self.backbone._fc = nn.Identity()          # efficientnet.py:68
        if freeze_bn:
            self._freeze_bn_layers()
        self.feature_dim = 1792
        self.dropout = nn.Dropout(p=dropout_rate)      # efficientnet.py:78
        self.intermediate_features: Dict[str, torch.Tensor] = {}
        for idx, name in ((5, "reduction_2"), (10, "reduction_4"), (21, "reduction_5")):
            self.backbone._blocks[idx].register_forward_hook(self._stash(name))  # :110-118

    def _stash(self, name):
        def hook(_m, _i, out):
            self.intermediate_features[name] = out
        return hook

    def _freeze_bn_layers(self):                       # efficientnet.py:84-90
        for m in self.backbone.modules():
            if isinstance(m, (nn.BatchNorm2d, nn.BatchNorm1d)):
                m.eval()
                for p in m.parameters():
                    p.requires_grad = False

    def forward(self, x, return_intermediate=False):   # efficientnet.py:122-151
        f = self.backbone.extract_features(x)
        f = self.backbone._avg_pooling(f).flatten(1)
        f = self.dropout(f)
        return (f, self.intermediate_features) if return_intermediate else (f, None)

    def get_feature_maps(self, x):                     # efficientnet.py:153-163
        return self.backbone.extract_features(x)

    def train(self, mode=True):                        # efficientnet.py:165-170
        super().train(mode)
        if self.freeze_bn:
            self._freeze_bn_layers()
        return self


class LandmarkAttention(nn.Module):
    """src/feature_extraction/landmark_attention.py:13-150."""

    def __init__(self, feature_size=(7, 7), sigma=1.5, learnable=True):
        super().__init__()
        self.feature_size, self.sigma, self.learnable = feature_size, sigma, learnable
        if learnable:
            self.attention_weights = nn.Parameter(torch.ones(5))
        else:
            self.register_buffer("attention_weights", torch.ones(5))

    def forward(self, feature_maps, landmarks):        # :49-74
        _, _, H, W = feature_maps.shape
        return feature_maps * self._create_attention_map(landmarks, (H, W), feature_maps.device)

    def _create_attention_map(self, landmarks, feature_size, device):   # :76-130
        B = landmarks.shape[0]
        H, W = feature_size
        sx, sy = W / 224.0, H / 224.0                  # :97-98 (constant 224 regardless of input)
        lm = landmarks.clone()
        lm[:, :, 0] *= sx
        lm[:, :, 1] *= sy
        ys = torch.arange(H, device=device, dtype=torch.float32).view(1, 1, H, 1)
        xs = torch.arange(W, device=device, dtype=torch.float32).view(1, 1, 1, W)
        amap = torch.zeros(B, 1, H, W, device=device)
        for i in range(5):
            lx = lm[:, i:i + 1, 0:1].view(B, 1, 1, 1)
            ly = lm[:, i:i + 1, 1:2].view(B, 1, 1, 1)
            d2 = (xs - lx) ** 2 + (ys - ly) ** 2
            g = torch.exp(-d2 / (2 * self.sigma ** 2))
            amap += g * self.attention_weights[i]
        amap = amap / (amap.max() + 1e-8)              # :125 -- max over the whole batch
        return torch.clamp(amap, min=0.1, max=1.0)     # :128


class SpatialAttention(nn.Module):
    """landmark_attention.py:153-192."""

    def __init__(self, kernel_size=7):
        super().__init__()
        self.conv = nn.Conv2d(2, 1, kernel_size=kernel_size, padding=kernel_size // 2, bias=False)
        self.sigmoid = nn.Sigmoid()

    def forward(self, x):
        avg = torch.mean(x, dim=1, keepdim=True)
        mx, _ = torch.max(x, dim=1, keepdim=True)
        return x * self.sigmoid(self.conv(torch.cat([avg, mx], dim=1)))


class ChannelAttention(nn.Module):
    """landmark_attention.py:195-241."""

    def __init__(self, channels, reduction=16):
        super().__init__()
        self.avg_pool = nn.AdaptiveAvgPool2d(1)
        self.max_pool = nn.AdaptiveMaxPool2d(1)
        self.fc = nn.Sequential(
            nn.Linear(channels, channels // reduction, bias=False),
            nn.ReLU(inplace=True),
            nn.Linear(channels // reduction, channels, bias=False),
        )
        self.sigmoid = nn.Sigmoid()

    def forward(self, x):
        B, C, _, _ = x.shape
        a = self.fc(self.avg_pool(x).view(B, C))
        m = self.fc(self.max_pool(x).view(B, C))
        return x * self.sigmoid(a + m).view(B, C, 1, 1)


class HybridAttention(nn.Module):
    """landmark_attention.py:244-310: landmark -> channel -> spatial."""

    def __init__(self, channels, feature_size=(7, 7), use_landmark=True, use_spatial=True, use_channel=True):
        super().__init__()
        self.use_landmark, self.use_spatial, self.use_channel = use_landmark, use_spatial, use_channel
        if use_landmark:
            self.landmark_attn = LandmarkAttention(feature_size=feature_size, learnable=True)
        if use_spatial:
            self.spatial_attn = SpatialAttention()
        if use_channel:
            self.channel_attn = ChannelAttention(channels)

    def forward(self, feature_maps, landmarks=None):
        x = feature_maps
        if self.use_landmark and landmarks is not None:
            x = self.landmark_attn(x, landmarks)
        if self.use_channel:
            x = self.channel_attn(x)
        if self.use_spatial:
            x = self.spatial_attn(x)
        return x


class DeepfakeFeatureExtractor(nn.Module):
    """src/feature_extraction/feature_extractor.py:16-178."""

    def __init__(self, pretrained=True, freeze_bn=False, dropout_rate=0.4, use_attention=True,
                 attention_config: Optional[Dict] = None):
        super().__init__()
        self.backbone = EfficientNetB4Backbone(pretrained=pretrained, freeze_bn=freeze_bn,
                                               dropout_rate=dropout_rate, extract_features=True)
        self.use_attention = use_attention
        self.feature_dim = self.backbone.feature_dim
        if use_attention:
            if attention_config is None:
                attention_config = {"use_landmark": True, "use_spatial": True, "use_channel": True}
            self.attention = HybridAttention(channels=self.feature_dim, feature_size=(7, 7), **attention_config)
        else:
            self.attention = None

    def forward(self, images, landmarks=None, return_attention=False):   # :74-117
        fmap = self.backbone.get_feature_maps(images)
        amap = None
        if self.use_attention and self.attention is not None:
            if return_attention and landmarks is not None:
                amap = self.attention.landmark_attn._create_attention_map(landmarks, (7, 7), images.device)
            fmap = self.attention(fmap, landmarks)
        feats = F.adaptive_avg_pool2d(fmap, 1).flatten(1)
        feats = self.backbone.dropout(feats)
        return (feats, amap) if return_attention else (feats, None)

    def extract_multi_scale_features(self, images, landmarks=None):      # :119-154
        _, inter = self.backbone(images, return_intermediate=True)
        out = {}
        if inter:
            for name, feat in inter.items():
                out[name] = F.adaptive_avg_pool2d(feat, 1).flatten(1)
        out["final"], _ = self.forward(images, landmarks)
        return out

    def get_embedding(self, images, landmarks=None, normalize=True):     # :156-178
        feats, _ = self.forward(images, landmarks)
        return F.normalize(feats, p=2, dim=1) if normalize else feats


class DeepfakeDetectionModel(nn.Module):
    """feature_extractor.py:184-299."""

    def __init__(self, num_classes=2, pretrained=True, feature_extractor_config: Optional[Dict] = None,
                 classifier_hidden_dims=(512, 128, 32), dropout_rate=0.4):
        super().__init__()
        if feature_extractor_config is None:
            feature_extractor_config = {"pretrained": pretrained, "use_attention": True}
        self.feature_extractor = DeepfakeFeatureExtractor(**feature_extractor_config)
        layers, d = [], self.feature_extractor.feature_dim
        for h in classifier_hidden_dims:
            layers += [nn.Linear(d, h), nn.BatchNorm1d(h), nn.ReLU(inplace=True), nn.Dropout(dropout_rate)]
            d = h
        layers.append(nn.Linear(d, num_classes))
        self.classifier = nn.Sequential(*layers)
        self.num_classes = num_classes

    def forward(self, images, landmarks=None, return_features=False) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
        feats, _ = self.feature_extractor(images, landmarks)
        logits = self.classifier(feats)
        return (logits, feats) if return_features else (logits, None)

    def predict(self, images, landmarks=None, return_probs=True):
        with torch.no_grad():
            logits, _ = self.forward(images, landmarks)
            return torch.softmax(logits, dim=1) if return_probs else logits


class FocalLoss(nn.Module):
    """src/training/losses.py:12-62."""

    def __init__(self, alpha=None, gamma=2.0, reduction="mean"):
        super().__init__()
        self.alpha, self.gamma, self.reduction = alpha, gamma, reduction

    def forward(self, inputs, targets):
        ce = F.cross_entropy(inputs, targets, reduction="none")
        pt = torch.exp(-ce)
        fl = (1 - pt) ** self.gamma * ce
        if self.alpha is not None:
            fl = self.alpha[targets] * fl
        if self.reduction == "mean":
            return fl.mean()
        return fl.sum() if self.reduction == "sum" else fl


class ContrastiveLoss(nn.Module):
    """losses.py:65-113 (label 1 == same class, as the code -- not the docstring -- has it)."""

    def __init__(self, margin=1.0, distance="euclidean"):
        super().__init__()
        self.margin, self.distance = margin, distance

    def forward(self, e1, e2, labels):
        if self.distance == "euclidean":
            d = F.pairwise_distance(e1, e2)
        else:
            d = 1 - F.cosine_similarity(e1, e2)
        same = labels * d.pow(2)
        diff = (1 - labels) * F.relu(self.margin - d).pow(2)
        return (same + diff).mean()


class CombinedLoss(nn.Module):
    """losses.py:164-247."""

    def __init__(self, weights: dict, class_weights: Optional[torch.Tensor] = None):
        super().__init__()
        self.weights = weights
        self.ce_loss = nn.CrossEntropyLoss(weight=class_weights)
        self.focal_loss = FocalLoss(alpha=class_weights, gamma=2.0)
        self.contrastive_loss = ContrastiveLoss(margin=1.0)

    def forward(self, logits, targets, features=None) -> dict:
        out, total = {}, 0.0
        w = self.weights
        if "ce" in w and w["ce"] > 0:
            out["ce"] = self.ce_loss(logits, targets)
            total += w["ce"] * out["ce"]
        if "focal" in w and w["focal"] > 0:
            out["focal"] = self.focal_loss(logits, targets)
            total += w["focal"] * out["focal"]
        if features is not None and "contrastive" in w and w["contrastive"] > 0:
            if features.size(0) >= 2:
                f1, f2 = features[:-1:2], features[1::2]
                pair = (targets[:-1:2] == targets[1::2]).float()
                if len(f1) > 0:
                    out["contrastive"] = self.contrastive_loss(f1, f2, pair)
                    total += w["contrastive"] * out["contrastive"]
        out["total"] = total
        return out


# --------------------------------------------------------------------------------------
def get_oracle():
    """The checker: the real reference when mounted, else this restatement."""
    import types
    from .load_reference import load_reference, reference_available
    if reference_available():
        return load_reference()
    return types.SimpleNamespace(
        DeepfakeDetectionModel=DeepfakeDetectionModel,
        DeepfakeFeatureExtractor=DeepfakeFeatureExtractor,
        EfficientNetB4Backbone=EfficientNetB4Backbone,
        HybridAttention=HybridAttention,
        LandmarkAttention=LandmarkAttention,
        ChannelAttention=ChannelAttention,
        SpatialAttention=SpatialAttention,
        CombinedLoss=CombinedLoss,
        FocalLoss=FocalLoss,
        ContrastiveLoss=ContrastiveLoss,
        kind="port",
    )


MODEL_CONFIG = {   # config/model_config.yaml:4-19 (the `model:` section), pretrained off (no weights offline)
    "num_classes": 2,
    "pretrained": False,
    "feature_extractor_config": {
        "pretrained": False,
        "freeze_bn": False,
        "dropout_rate": 0.4,
        "use_attention": True,
        "attention_config": {"use_landmark": True, "use_spatial": True, "use_channel": True},
    },
    "classifier_hidden_dims": [512, 128, 32],
    "dropout_rate": 0.4,
}
