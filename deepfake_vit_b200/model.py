"""Drop-in for the reference's model classes, computed by libdfvit (sm_100a CUDA kernels).

Mirrors the reference's nn.Module API for the hot path:
  DeepfakeDetectionModel      src/feature_extraction/feature_extractor.py:184-299
  DeepfakeFeatureExtractor    src/feature_extraction/feature_extractor.py:16-178
  EfficientNetB4Backbone      src/feature_extraction/efficientnet.py:13-170
  HybridAttention & friends   src/feature_extraction/landmark_attention.py:13-310
Same constructor arguments (the YAML `model:` mapping, config/model_config.yaml:4-19), same
`forward(images, landmarks=None, return_features=False) -> (logits, features|None)`, same
module tree and therefore the same state_dict keys, so
`best_model.pth`-style checkpoints load with strict=True in either direction.

The sub-modules below only HOLD parameters (they are torch modules so that registration
order, default initialisation and state_dict layout are exactly the reference's); none of
their `forward`s is ever used.  All arithmetic happens in libdfvit through one C call.
"""
import ctypes as C
import weakref
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn as nn

from . import _lib, ops
from .parallel import BucketedAllReduce, flat_offsets
from ._lib import check, lib

BN_EPS, BN_MOM = 1e-3, 0.01   # efficientnet-pytorch global params for B4
IMAGENET_MEAN, IMAGENET_STD = (0.485, 0.456, 0.406), (0.229, 0.224, 0.225)   # src/data/dataset.py:59-60, task.ipynb:364-365


def _holder(cls):
    class Holder(cls):
        def forward(self, *a, **k):   # pragma: no cover
            raise RuntimeError(f"{cls.__name__} here is a parameter holder; libdfvit computes this layer")
    Holder.__name__ = Holder.__qualname__ = cls.__name__
    return Holder


_Conv2d, _BN2d, _BN1d, _Linear = (_holder(c) for c in (nn.Conv2d, nn.BatchNorm2d, nn.BatchNorm1d, nn.Linear))


class _MBConvParams(nn.Module):
    """Parameters of one MBConvBlock of efficientnet-pytorch 0.7.1 (names as upstream)."""

    def __init__(self, info: _lib.BlockInfo):
        super().__init__()
        self.info = {f[0]: getattr(info, f[0]) for f in info._fields_}
        cin, cmid, cout, k, sq = info.c_in, info.c_mid, info.c_out, info.kernel, info.se_squeeze
        if info.has_expand:
            self._expand_conv = _Conv2d(cin, cmid, 1, bias=False)
            self._bn0 = _BN2d(cmid, eps=BN_EPS, momentum=BN_MOM)
        self._depthwise_conv = _Conv2d(cmid, cmid, k, stride=info.stride, groups=cmid, bias=False)
        self._bn1 = _BN2d(cmid, eps=BN_EPS, momentum=BN_MOM)
        self._se_reduce = _Conv2d(cmid, sq, 1)
        self._se_expand = _Conv2d(sq, cmid, 1)
        self._project_conv = _Conv2d(cmid, cout, 1, bias=False)
        self._bn2 = _BN2d(cout, eps=BN_EPS, momentum=BN_MOM)


class _EfficientNetB4Params(nn.Module):
    """Parameter tree of `EfficientNet.from_name('efficientnet-b4')` with `_fc = Identity`."""

    def __init__(self):
        super().__init__()
        stem_c, head_c = lib.dfv_b4_stem_channels(), lib.dfv_b4_head_channels()
        self._conv_stem = _Conv2d(3, stem_c, 3, stride=2, bias=False)
        self._bn0 = _BN2d(stem_c, eps=BN_EPS, momentum=BN_MOM)
        blocks = _lib.b4_blocks()
        self._blocks = nn.ModuleList(_MBConvParams(b) for b in blocks)
        self._conv_head = _Conv2d(blocks[-1].c_out, head_c, 1, bias=False)
        self._bn1 = _BN2d(head_c, eps=BN_EPS, momentum=BN_MOM)
        # upstream builds Linear(1792, 1000) here and the reference replaces it with Identity
        # (efficientnet.py:68); build-and-drop keeps seeded initialisation identical.
        nn.Linear(head_c, 1000)
        self._fc = nn.Identity()
        self.head_channels = head_c
        self.drop_connect_rate = 0.2      # efficientnet-pytorch global params for B4


class EfficientNetB4Backbone(nn.Module):
    def __init__(self, pretrained=True, freeze_bn=False, dropout_rate=0.4, extract_features=True):
        super().__init__()
        # `pretrained` only matters if ./model/efficientnet-b4-6ed6700e.pth exists in the
        # reference (efficientnet.py:48-54); offline it never does -> random init, never raises.
        self.backbone = _EfficientNetB4Params()
        self.extract_features = extract_features
        self.freeze_bn = freeze_bn
        self.feature_dim = self.backbone.head_channels
        self.dropout = nn.Dropout(p=dropout_rate)
        self.intermediate_features: Dict[str, torch.Tensor] = {}
        if freeze_bn:
            self._freeze_bn_layers()

    def _freeze_bn_layers(self):
        for m in self.backbone.modules():
            if isinstance(m, (nn.BatchNorm2d, nn.BatchNorm1d)):
                m.eval()
                for p in m.parameters():
                    p.requires_grad = False

    def train(self, mode: bool = True):
        super().train(mode)
        if self.freeze_bn:
            self._freeze_bn_layers()
        return self

    def _model(self):
        owner = getattr(self, "_owner", None)
        m = owner() if owner is not None else None
        if m is None:
            raise RuntimeError("EfficientNetB4Backbone is a parameter holder; use it through DeepfakeDetectionModel")
        return m

    def forward(self, x, return_intermediate: bool = False):
        """-> (features (B, 1792), {'reduction_2' | 'reduction_4' | 'reduction_5': NCHW maps of blocks 5 / 10 / 21} | None)
        (efficientnet.py:122-151): backbone maps -> global average pool -> dropout, no attention.  Computed by the owning
        model's libdfvit forward with the attention stage switched off."""
        m = self._model()
        if m.training:
            feats, taps = m._train_features(x, None, attention=False, want_taps=return_intermediate)
        else:
            _, feats, _, taps = m._infer(x, None, taps=return_intermediate, attention=False)
        inter = None
        if return_intermediate:
            inter = {name: taps[1 + blk].permute(0, 3, 1, 2).float()
                     for name, blk in (("reduction_2", 5), ("reduction_4", 10), ("reduction_5", 21))}
            self.intermediate_features = inter
        return feats, inter

    def get_feature_maps(self, x):
        """(B, 1792, h, w) final feature maps (efficientnet.py:153-163), eval mode."""
        m = self._model()
        assert not m.training, "get_feature_maps is the eval-mode side API"
        _, _, _, taps = m._infer(x, None, taps=True, attention=False)
        return taps[-1].permute(0, 3, 1, 2).float()


class LandmarkAttention(nn.Module):
    def __init__(self, feature_size=(7, 7), sigma=1.5, learnable=True):
        super().__init__()
        self.feature_size, self.sigma, self.learnable = feature_size, sigma, learnable
        if learnable:
            self.attention_weights = nn.Parameter(torch.ones(5))
        else:
            self.register_buffer("attention_weights", torch.ones(5))

    def _create_attention_map(self, landmarks, feature_size, device=None, group=0):
        """(B,1,H,W) heat-map, landmark_attention.py:76-130, computed by dfv_landmark_heatmap_fwd."""
        H, W = feature_size
        lm = landmarks.detach().to(self.attention_weights.device, torch.float32).contiguous()
        heat = ops.landmark_heatmap(lm, self.attention_weights.detach().float().contiguous(), H, W, 224.0,
                                    self.sigma, group)
        return heat.unsqueeze(1)


class SpatialAttention(nn.Module):
    def __init__(self, kernel_size=7):
        super().__init__()
        self.conv = _Conv2d(2, 1, kernel_size, padding=kernel_size // 2, bias=False)
        self.sigmoid = nn.Sigmoid()


class ChannelAttention(nn.Module):
    def __init__(self, channels, reduction=16):
        super().__init__()
        self.avg_pool = nn.AdaptiveAvgPool2d(1)
        self.max_pool = nn.AdaptiveMaxPool2d(1)
        self.fc = nn.Sequential(_Linear(channels, channels // reduction, bias=False), nn.ReLU(inplace=True),
                                _Linear(channels // reduction, channels, bias=False))
        self.sigmoid = nn.Sigmoid()


class HybridAttention(nn.Module):
    def __init__(self, channels, feature_size=(7, 7), use_landmark=True, use_spatial=True, use_channel=True):
        super().__init__()
        self.use_landmark, self.use_spatial, self.use_channel = use_landmark, use_spatial, use_channel
        if use_landmark:
            self.landmark_attn = LandmarkAttention(feature_size=feature_size, learnable=True)
        if use_spatial:
            self.spatial_attn = SpatialAttention()
        if use_channel:
            self.channel_attn = ChannelAttention(channels)


class DeepfakeFeatureExtractor(nn.Module):
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

    # ---- the reference's public side APIs (feature_extractor.py:74-178), eval mode, computed by the owning model's
    # ---- libdfvit forward (this module itself only holds parameters)
    def _model(self):
        owner = getattr(self, "_owner", None)
        m = owner() if owner is not None else None
        if m is None:
            raise RuntimeError("DeepfakeFeatureExtractor is a parameter holder; use it through DeepfakeDetectionModel")
        return m

    def forward(self, images, landmarks=None, return_attention=False):
        """-> (features (B, 1792), attention_map (B, 1, 7, 7) | None)  (feature_extractor.py:74-117; the reference
        renders the returned map on a hard-coded 7x7 grid whatever the real feature size is, :100-103)."""
        m = self._model()
        if m.training:      # train mode: the same autograd node as model.forward, without the classifier's logits
            feats, _ = m._train_features(images, landmarks)
        else:
            _, feats, _, _ = m._infer(images, landmarks)
        amap = None
        if return_attention and landmarks is not None and self.use_attention and self.attention is not None \
                and getattr(self.attention, "use_landmark", False):
            with torch.no_grad():
                amap = self.attention.landmark_attn._create_attention_map(landmarks, (7, 7), images.device,
                                                                          group=int(m.landmark_max_group))
        return feats, amap

    def extract_multi_scale_features(self, images, landmarks=None):
        """{'reduction_2', 'reduction_4', 'reduction_5'}: pooled outputs of blocks 5 / 10 / 21 (efficientnet.py:110-118),
        'final': the attended features (feature_extractor.py:119-154)."""
        m = self._model()
        assert not m.training
        _, feats, _, taps = m._infer(images, landmarks, taps=True)
        out = {name: ops.global_avg_pool(taps[1 + blk]) for name, blk in (("reduction_2", 5), ("reduction_4", 10), ("reduction_5", 21))}
        out["final"] = feats
        return out

    def get_embedding(self, images, landmarks=None, normalize=True):
        """L2-normalised features (feature_extractor.py:156-178)."""
        feats, _ = self.forward(images, landmarks)
        return ops.l2_normalize(feats.contiguous()) if normalize else feats


class _TrainFn(torch.autograd.Function):
    """forward = dfv_train_fwd, backward = dfv_train_bwd (+ the data-parallel all-reduce)."""

    @staticmethod
    def forward(ctx, model, images, landmarks, opts, *params):
        opts = dict(opts)
        ctx.detached = opts.pop("detached_params", False)
        logits, feats, run = model._train_fwd(images, landmarks, **opts)
        ctx.model, ctx.run, ctx.n_inputs = model, run, len(params)
        return logits, feats

    @staticmethod
    def backward(ctx, dlogits, dfeats):
        grads = ctx.model._train_bwd(ctx.run, dlogits, dfeats)
        if ctx.detached:      # captured step: the caller attaches the gradient views itself (no AccumulateGrad nodes in the graph)
            ctx.model._detached_grads = grads
            return (None, None, None, None) + (None,) * ctx.n_inputs
        return (None, None, None, None) + tuple(grads)


class _Packed:
    """Folded, device-resident weights for one (dtype, device) pair (written by dfv_pack_weights)."""

    def __init__(self):
        self.key = None
        self.blob = None
        self.head: Optional[ops.HeadPack] = None
        self.ca_w2_t = None


class DeepfakeDetectionModel(nn.Module):
    """Reference-compatible module whose forward is one libdfvit call."""

    def __init__(self, num_classes: int = 2, pretrained: bool = True,
                 feature_extractor_config: Optional[Dict] = None,
                 classifier_hidden_dims: List[int] = [512, 128, 32], dropout_rate: float = 0.4):
        super().__init__()
        if feature_extractor_config is None:   # top-level `pretrained` ignored otherwise (:210-218)
            feature_extractor_config = {"pretrained": pretrained, "use_attention": True}
        self.feature_extractor = DeepfakeFeatureExtractor(**feature_extractor_config)
        # the side APIs of the parameter-holder sub-modules run this model's forward
        object.__setattr__(self.feature_extractor, "_owner", weakref.ref(self))
        object.__setattr__(self.feature_extractor.backbone, "_owner", weakref.ref(self))
        layers, d = [], self.feature_extractor.feature_dim
        for h in classifier_hidden_dims:
            layers += [_Linear(d, h), _BN1d(h), nn.ReLU(inplace=True), nn.Dropout(dropout_rate)]
            d = h
        layers.append(_Linear(d, num_classes))
        self.classifier = nn.Sequential(*layers)
        self.num_classes = num_classes
        # ---- knobs that are not part of the reference API
        self.compute_dtype = torch.bfloat16     # or torch.float32 (parity mode)
        self.landmark_max_group = 0             # images per heat-map max group; 0 = whole call (reference)
        self.landmark_max_scope = "rank"        # "rank": the maximum over this call's batch (= the reference run on this shard);
                                                # "global" (eval, torch.distributed): the maximum over ALL ranks' batches, one 4-byte
                                                # all-reduce(MAX) -- bit-equal to one GPU scoring the concatenated batch
        self.ddp_allreduce = True               # average gradients over torch.distributed ranks inside backward
        self.ddp_bucket_floats = 4 << 20        # ~16 MB gradient buckets, all-reduced while the backward still runs
        # uint8 (B, H, W, 3) RGB crops are normalised inside the stem: (u8 / 255 - mean) / std (dataset.py:95-98)
        self.input_mean, self.input_std = IMAGENET_MEAN, IMAGENET_STD
        self._packed: Dict[Tuple, _Packed] = {}
        self._workspace: Dict[Tuple, torch.Tensor] = {}
        self._pinned_ws = set()                 # workspace keys a live CUDA graph replays from: never evicted
        self._arena_pool: Dict[Tuple, List[torch.Tensor]] = {}
        self._train_scratch = None
        self._flat_grad = None
        self._reducer: Optional[BucketedAllReduce] = None
        self._epoch = 0                         # bumped by everything that rewrites weights behind autograd's back
        self._state_list = None
        self._bn_list = None
        self._seed_dev = None                   # GraphedTrainStep: device word holding the dropout seed of a captured step
        self._detached_anchor = None            # GraphedTrainStep: the only differentiable input of a captured step
        self._detached_grads = None
        self._want_taps = False                 # debug: keep NHWC copies of every stage output of a train forward
        self._last_taps = None
        self._last_flat_grad = None

    # ------------------------------------------------------------------ packing
    def invalidate_packed(self):
        """Tell the model its weights changed through a path torch's version counters do not see (`p.data` writes, raw
        pointer writes by a fused optimizer, a broadcast into `.data`).  FusedAdamW.step, parallel.broadcast_parameters
        and the train-mode forward (running statistics) call it."""
        self._epoch += 1

    def _version_key(self):
        if self._state_list is None:
            self._state_list = list(self.parameters()) + list(self.buffers())
        return (self._epoch, sum(t._version for t in self._state_list), sum(t.data_ptr() & 0xFFFFFF for t in self._state_list[:4]))

    def _apply(self, fn, *a, **k):      # .to() / .cuda() / .float(): new storages
        self._state_list = None
        self._epoch += 1
        return super()._apply(fn, *a, **k)

    def _head_spec(self):
        dims, p_drop, i = [self.feature_extractor.feature_dim], 0.0, 0
        mods = list(self.classifier)
        while i < len(mods):
            dims.append(mods[i].out_features)
            if i + 1 < len(mods) and isinstance(mods[i + 1], nn.BatchNorm1d):
                p_drop = mods[i + 3].p
                i += 4
            else:
                i += 1
        return dims, p_drop

    @torch.no_grad()
    def _pack(self, dtype: torch.dtype, device) -> _Packed:
        """Fold BatchNorm and lay the weights out for the inference kernels -- dfv_pack_weights, a handful of launches
        over the module's own parameter storage (no host-framework arithmetic)."""
        key = (dtype, str(device))
        pk = self._packed.setdefault(key, _Packed())
        ver = self._version_key()
        if pk.key == ver:
            return pk
        code = ops.dtype_code(dtype)
        T, dims, _ = self._train_tensors()
        for t in T:
            if t is not None:
                assert t.is_cuda and t.dtype == torch.float32 and t.is_contiguous(), "parameters must be contiguous fp32 CUDA tensors"
        f32 = dict(device=device, dtype=torch.float32)
        blob = torch.empty(lib.dfv_blob_bytes(code), dtype=torch.uint8, device=device)
        w_t = [torch.empty(dims[l], dims[l + 1], **f32) for l in range(len(dims) - 1)]
        bs = [torch.empty(dims[l + 1], **f32) for l in range(len(dims) - 1)]
        head = ops.HeadPack(w_t, bs)
        att = self.feature_extractor.attention
        use_c = bool(self.feature_extractor.use_attention and att is not None and att.use_channel)
        ca_hidden = att.channel_attn.fc[0].out_features if use_c else 0
        ca_w2_t = torch.empty(ca_hidden, self.feature_extractor.feature_dim, **f32) if use_c else None
        bn1d = [m for m in self.classifier if isinstance(m, nn.BatchNorm1d)]
        a = _lib.PackArgs()
        a.dtype, a.bn_eps = code, self.feature_extractor.backbone.backbone._bn0.eps
        a.cls_bn_eps = bn1d[0].eps if bn1d else 1e-5
        a.params = (C.c_void_p * len(T))(*[t.data_ptr() if t is not None else None for t in T])
        a.blob = blob.data_ptr()
        a.head_layers, a.head_dims = head.n, C.cast(head.dims, C.POINTER(C.c_int32))
        a.head_w_t, a.head_b = head.wp, head.bp
        a.ca_hidden, a.ca_w2_t = ca_hidden, (ca_w2_t.data_ptr() if use_c else None)
        check(lib.dfv_pack_weights(C.byref(a), torch.cuda.current_stream().cuda_stream))
        pk.blob, pk.head, pk.ca_w2_t, pk.key = blob, head, ca_w2_t, ver
        return pk

    def _ws(self, code, B, H, W, device):
        key = (code, B, H, W, str(device))
        ws = self._workspace.get(key)
        if ws is None:
            n = lib.dfv_infer_workspace_bytes(code, B, H, W)
            if n == 0:
                check(-1)
            # keep one eager shape resident, plus every shape a live CUDA graph was captured on
            self._workspace = {k: v for k, v in self._workspace.items() if k in self._pinned_ws}
            ws = self._workspace[key] = torch.empty(n, dtype=torch.uint8, device=device)
        return ws

    def _images(self, images):
        """-> (fp32 NCHW tensor | None, uint8 HWC tensor | None, B, H, W)."""
        if not images.is_cuda:
            raise RuntimeError("deepfake_vit_b200 runs on sm_100 CUDA devices only (no CPU path); move the "
                               "model and inputs to cuda")
        check(lib.dfv_device_check())
        if images.dtype == torch.uint8:
            assert images.dim() == 4 and images.shape[3] == 3, "uint8 images must be (B, H, W, 3) RGB crops"
            u8 = images.detach().contiguous()
            return None, u8, u8.shape[0], u8.shape[1], u8.shape[2]
        x = images.detach().to(torch.float32).contiguous()
        assert x.dim() == 4 and x.shape[1] == 3, "images must be (B, 3, H, W)"
        return x, None, x.shape[0], x.shape[2], x.shape[3]

    def _norm6(self):
        return (C.c_float * 6)(*self.input_mean, *self.input_std)

    def _check_modes(self):
        """Only the top-level mode selects the path; a sub-module left in the other mode would be silently ignored."""
        if self._bn_list is None:
            inside = {id(x) for x in self.feature_extractor.backbone.backbone.modules()}
            self._bn_list = [(n, m, id(m) in inside) for n, m in self.named_modules() if isinstance(m, (nn.BatchNorm2d, nn.BatchNorm1d))]
        frozen = self.feature_extractor.backbone.freeze_bn
        for name, m, in_backbone in self._bn_list:
            if m.training != self.training:
                if frozen and self.training and in_backbone:
                    continue        # freeze_bn: backbone BatchNorm stays in eval mode by construction
                raise RuntimeError(f"{name}.training = {m.training} but the model is in {'train' if self.training else 'eval'} "
                                   "mode: per-sub-module modes are not supported (use freeze_bn=True to freeze the backbone BatchNorm)")

    # ------------------------------------------------------------------ inference
    @torch.no_grad()
    def _infer(self, images, landmarks, want_heat=False, taps=False, attention=True):
        x, u8, B, H, W = self._images(images)
        dev = images.device
        dtype = self.compute_dtype
        code = ops.dtype_code(dtype)
        pk = self._pack(dtype, dev)
        ws = self._ws(code, B, H, W, dev)
        att = self.feature_extractor.attention
        use_att = bool(attention and self.feature_extractor.use_attention and att is not None)
        ho, wo = C.c_int(), C.c_int()
        check(lib.dfv_b4_output_hw(H, W, C.byref(ho), C.byref(wo)))
        Hf, Wf = ho.value, wo.value
        logits = torch.empty(B, self.num_classes, device=dev, dtype=torch.float32)
        feats = torch.empty(B, self.feature_extractor.feature_dim, device=dev, dtype=torch.float32)
        lm = None
        if landmarks is not None and use_att and att.use_landmark:
            lm = landmarks.detach().to(device=dev, dtype=torch.float32).contiguous()
            assert lm.shape == (B, 5, 2), "landmarks must be (B, 5, 2)"
        heat = torch.empty(B, Hf, Wf, device=dev, dtype=torch.float32) if (want_heat and lm is not None) else None

        tap_tensors, tap_ptrs = None, None
        if taps:
            tap_tensors = self._tap_tensors(B, H, W, dev, dtype)
            tap_ptrs = (C.c_void_p * len(tap_tensors))(*[t.data_ptr() for t in tap_tensors])

        a = _lib.InferArgs()
        a.dtype, a.B, a.H, a.W = code, B, H, W
        a.use_attention = int(use_att)
        a.use_landmark = int(use_att and att.use_landmark)
        a.use_channel = int(use_att and att.use_channel)
        a.use_spatial = int(use_att and att.use_spatial)
        a.heat_group = int(self.landmark_max_group)
        a.landmark_ref_size = 224.0
        floor_key = None
        if lm is not None and self.landmark_max_scope == "global":
            from . import parallel
            assert int(self.landmark_max_group) in (0, B), "the global normaliser is the whole-call maximum of every rank"
            floor_key = parallel.allreduce_max_key(ops.landmark_max_key(lm, att.landmark_attn.attention_weights.detach(), Hf, Wf,
                                                                        224.0, att.landmark_attn.sigma))
            a.heat_max_floor = floor_key.data_ptr()
        a.blob = pk.blob.data_ptr()
        a.images_nchw = x.data_ptr() if x is not None else None
        if u8 is not None:
            a.images_u8, a.u8_norm = u8.data_ptr(), self._norm6()
        a.landmarks = lm.data_ptr() if lm is not None else None
        # small attention tensors are read from the parameters themselves (fp32, contiguous, on the device)
        if use_att and att.use_landmark:
            a.lm_weights = att.landmark_attn.attention_weights.data_ptr()
        if use_att and att.use_channel:
            a.ca_w1 = att.channel_attn.fc[0].weight.data_ptr()
            a.ca_w2_t = pk.ca_w2_t.data_ptr()
            a.ca_hidden = att.channel_attn.fc[0].out_features
        if use_att and att.use_spatial:
            a.sa_w = att.spatial_attn.conv.weight.data_ptr()
        a.head_w_t, a.head_b = pk.head.wp, pk.head.bp
        a.head_dims, a.head_layers = C.cast(pk.head.dims, C.POINTER(C.c_int32)), pk.head.n
        a.workspace, a.workspace_bytes = ws.data_ptr(), ws.numel()
        a.logits, a.features = logits.data_ptr(), feats.data_ptr()
        a.heat = heat.data_ptr() if heat is not None else None
        a.taps = tap_ptrs if tap_ptrs is not None else None
        check(lib.dfv_infer_fwd(C.byref(a), torch.cuda.current_stream().cuda_stream))
        return logits, feats, heat, tap_tensors

    # ------------------------------------------------------------------ training
    def _train_tensors(self):
        """The module's own parameter / buffer tensors in the C ABI's table order (dfv_train_index)."""
        T = [None] * lib.dfv_train_table_size()
        ix, cx = lib.dfv_train_index, lib.dfv_train_cls_index

        def bn(block, base, m):
            T[ix(block, base)], T[ix(block, base + 1)] = m.weight, m.bias
            T[ix(block, base + 2)], T[ix(block, base + 3)] = m.running_mean, m.running_var

        bb = self.feature_extractor.backbone.backbone
        T[ix(-1, _lib.TG_STEM_W)] = bb._conv_stem.weight
        bn(-1, _lib.TG_STEM_G, bb._bn0)
        for i, blk in enumerate(bb._blocks):
            if blk.info["has_expand"]:
                T[ix(i, _lib.T_EXPAND_W)] = blk._expand_conv.weight
                bn(i, _lib.T_BN0_G, blk._bn0)
            T[ix(i, _lib.T_DW_W)] = blk._depthwise_conv.weight
            bn(i, _lib.T_BN1_G, blk._bn1)
            T[ix(i, _lib.T_SE_R_W)], T[ix(i, _lib.T_SE_R_B)] = blk._se_reduce.weight, blk._se_reduce.bias
            T[ix(i, _lib.T_SE_E_W)], T[ix(i, _lib.T_SE_E_B)] = blk._se_expand.weight, blk._se_expand.bias
            T[ix(i, _lib.T_PROJ_W)] = blk._project_conv.weight
            bn(i, _lib.T_BN2_G, blk._bn2)
        T[ix(-1, _lib.TG_HEAD_W)] = bb._conv_head.weight
        bn(-1, _lib.TG_HEAD_G, bb._bn1)
        att = self.feature_extractor.attention
        if att is not None and self.feature_extractor.use_attention:
            if att.use_landmark:
                T[ix(-1, _lib.TG_LM_W)] = att.landmark_attn.attention_weights
            if att.use_spatial:
                T[ix(-1, _lib.TG_SA_W)] = att.spatial_attn.conv.weight
            if att.use_channel:
                T[ix(-1, _lib.TG_CA_W1)] = att.channel_attn.fc[0].weight
                T[ix(-1, _lib.TG_CA_W2)] = att.channel_attn.fc[2].weight
        mods, layer, i = list(self.classifier), 0, 0
        while i < len(mods):
            lin = mods[i]
            T[cx(layer, 0)], T[cx(layer, 1)] = lin.weight, lin.bias
            if i + 1 < len(mods) and isinstance(mods[i + 1], nn.BatchNorm1d):
                b1 = mods[i + 1]
                T[cx(layer, 2)], T[cx(layer, 3)], T[cx(layer, 4)], T[cx(layer, 5)] = b1.weight, b1.bias, b1.running_mean, b1.running_var
                i += 4
            else:
                i += 1
            layer += 1
        dims, p_drop = self._head_spec()
        return T, dims, p_drop

    def _train_args(self, images, landmarks, T, dims, p_drop, attention=True):
        att = self.feature_extractor.attention
        use_att = bool(attention and self.feature_extractor.use_attention and att is not None)
        a = _lib.TrainArgs()
        a.dtype = ops.dtype_code(self.compute_dtype)
        a.B, _, a.H, a.W = images.shape
        a.use_attention = int(use_att)
        a.use_landmark = int(use_att and att.use_landmark)
        a.use_channel = int(use_att and att.use_channel)
        a.use_spatial = int(use_att and att.use_spatial)
        a.heat_group = int(self.landmark_max_group)
        if self.landmark_max_scope != "rank":
            raise RuntimeError("landmark_max_scope='global' is an inference mode: in training the maximum's gradient would have to "
                               "cross ranks; train with the per-rank normaliser (what DDP over the reference does)")
        a.landmark_ref_size = 224.0
        bb = self.feature_extractor.backbone.backbone
        a.bn_eps, a.bn_momentum = bb._bn0.eps, bb._bn0.momentum
        bn1d = [m for m in self.classifier if isinstance(m, nn.BatchNorm1d)]
        a.cls_bn_eps, a.cls_bn_momentum = (bn1d[0].eps, bn1d[0].momentum) if bn1d else (1e-5, 0.1)
        a.drop_connect_rate = float(bb.drop_connect_rate)
        a.feat_dropout = float(self.feature_extractor.backbone.dropout.p)
        a.cls_dropout = float(p_drop)
        a.freeze_bn = int(bool(self.feature_extractor.backbone.freeze_bn))
        for t in T:
            if t is not None:
                assert t.is_cuda and t.dtype == torch.float32 and t.is_contiguous(), "parameters must be contiguous fp32 CUDA tensors"
        a.params = (C.c_void_p * len(T))(*[t.data_ptr() if t is not None else None for t in T])
        a.images_nchw = images.data_ptr()
        a.landmarks = landmarks.data_ptr() if landmarks is not None else None
        a.ca_hidden = att.channel_attn.fc[0].out_features if (use_att and att.use_channel) else 0
        hd = (C.c_int32 * len(dims))(*dims)
        a.head_dims, a.head_layers = hd, len(dims) - 1
        return a, hd

    def _train_fwd(self, images, landmarks, needs_bwd=True, attention=True, want_taps=False):
        x, u8, B, H, W = self._images(images)
        dev = images.device
        att = self.feature_extractor.attention
        lm = None
        if landmarks is not None and attention and self.feature_extractor.use_attention and att is not None and att.use_landmark:
            lm = landmarks.detach().to(device=dev, dtype=torch.float32).contiguous()
            assert lm.shape == (B, 5, 2), "landmarks must be (B, 5, 2)"
        T, dims, p_drop = self._train_tensors()
        a, hd = self._train_args(images if x is None else x, lm, T, dims, p_drop, attention)
        a.B, a.H, a.W = B, H, W
        # uint8 crops: the stem's weight gradient reads the fp32 image, so it is normalised once into the tail of this
        # forward's own arena (same arithmetic as the fused stem; no per-step allocation)
        stage_bytes = B * 3 * H * W * 4 if u8 is not None else 0
        key = (a.dtype, B, H, W, str(dev), a.ca_hidden, stage_bytes)
        # every forward that will be differentiated OWNS its arena until its backward has run (two forwards before a
        # backward, siamese losses); finished arenas go back to a per-shape pool
        pool = self._arena_pool.setdefault(key, [])
        if pool:
            arena = pool.pop()
        else:
            n = lib.dfv_train_arena_bytes(a.dtype, B, H, W, hd, a.head_layers, a.ca_hidden)
            if n == 0:
                check(-1)
            for k in [k for k in self._arena_pool if k != key]:      # another shape: drop its idle arenas
                del self._arena_pool[k]
            pool = self._arena_pool.setdefault(key, [])
            arena = torch.empty((n + 255) // 256 * 256 + stage_bytes, dtype=torch.uint8, device=dev)
        a.arena, a.arena_bytes = arena.data_ptr(), arena.numel() - stage_bytes
        if u8 is not None:
            x = arena[arena.numel() - stage_bytes:].view(torch.float32).view(B, 3, H, W)
            check(lib.dfv_u8_to_nchw_f32(u8.data_ptr(), self._norm6(), x.data_ptr(), B, H, W, torch.cuda.current_stream().cuda_stream))
            a.images_nchw = x.data_ptr()
        logits = torch.empty(B, dims[-1], device=dev, dtype=torch.float32)
        feats = torch.empty(B, dims[0], device=dev, dtype=torch.float32)
        a.logits, a.features = logits.data_ptr(), feats.data_ptr()
        # dropout / drop-connect masks are a function of (seed, position); the seed comes from torch's CPU generator
        # (so torch.manual_seed reproduces a run) mixed with the data-parallel rank (ranks draw different masks)
        if self._seed_dev is not None:     # a captured step (GraphedTrainStep): the seed lives in a device word set before each replay
            a.seed, a.seed_dev = 0, self._seed_dev.data_ptr()
        else:
            a.seed = _mix_seed(int(torch.randint(0, 2 ** 62, (1,)).item()))
        taps = None
        if self._want_taps or want_taps:
            taps = self._tap_tensors(B, H, W, dev, self.compute_dtype)
            a.taps = (C.c_void_p * len(taps))(*[t.data_ptr() for t in taps])
        check(lib.dfv_train_fwd(C.byref(a), torch.cuda.current_stream().cuda_stream))
        frozen = self.feature_extractor.backbone.freeze_bn
        bb = self.feature_extractor.backbone.backbone
        nbt = [m.num_batches_tracked for m in self.modules() if isinstance(m, (nn.BatchNorm2d, nn.BatchNorm1d)) and m.training]
        if not frozen:
            assert len(nbt) > 0
        torch._foreach_add_(nbt, 1)
        self.invalidate_packed()          # running statistics were updated through raw pointers
        run = dict(args=a, keep=(hd, x, lm, T, feats), taps=taps, arena=arena, pool_key=key)
        if not needs_bwd:                 # nothing will come back for this arena: stream order makes it reusable at once
            pool.append(arena)
            run["arena"] = None
        self._last_taps = taps
        return logits, feats, run

    def _flat_layout(self):
        params = [p for _, p in self.named_parameters()]
        starts, total = flat_offsets([p.numel() for p in params])
        return params, starts, total

    def _train_bwd(self, run, dlogits, dfeats):
        a = run["args"]
        if run["arena"] is None:
            raise RuntimeError("backward through a train-mode forward that ran under torch.no_grad()")
        hd, images, lm, T, feats = run["keep"]
        dev = images.device
        params, starts, total = self._flat_layout()
        offs = {id(p): o for p, o in zip(params, starts)}
        # ONE flat fp32 gradient buffer per step (the all-reduce and the fused optimizer read it in place).  A fresh
        # zeroed allocation per backward: autograd hands these views out as .grad, so the previous step's buffer
        # may still be referenced by the caller (gradient accumulation adds into it).
        flat = torch.zeros(total, dtype=torch.float32, device=dev)
        gp = [flat.data_ptr() + 4 * offs[id(t)] if (t is not None and id(t) in offs) else None for t in T]
        a.grads = (C.c_void_p * len(T))(*gp)
        key = (a.dtype, a.B, a.H, a.W, str(dev))
        if self._train_scratch is None or self._train_scratch[0] != key:
            n = lib.dfv_train_scratch_bytes(a.dtype, a.B, a.H, a.W, hd, a.head_layers, a.ca_hidden)
            if n == 0:
                check(-1)
            self._train_scratch = (key, torch.empty(n, dtype=torch.uint8, device=dev))
        sc = self._train_scratch[1]
        a.scratch, a.scratch_bytes = sc.data_ptr(), sc.numel()
        dl = dlogits.detach().to(torch.float32).contiguous() if dlogits is not None else torch.zeros(a.B, self.num_classes, device=dev)
        df = dfeats.detach().to(torch.float32).contiguous() if dfeats is not None else None
        a.dlogits = dl.data_ptr()
        a.dfeatures = df.data_ptr() if df is not None else None
        reducer = None
        if self.ddp_allreduce and BucketedAllReduce.active():
            if self._reducer is None or self._reducer.total != total or self._reducer.bucket_floats != self.ddp_bucket_floats:
                self._reducer = BucketedAllReduce(self._unit_ranges(params, starts, total), total, self.ddp_bucket_floats, dev)
            reducer = self._reducer
            a.grad_events = reducer.event_table()
        check(lib.dfv_train_bwd(C.byref(a), torch.cuda.current_stream().cuda_stream))
        if reducer is not None:
            reducer.reduce(flat)          # bucket by bucket on a side stream, behind the events dfv_train_bwd recorded
        self._arena_pool.setdefault(run["pool_key"], []).append(run["arena"])
        run["arena"] = None
        for p, o in zip(params, starts):      # parameters the caller froze: no gradient leaves this function for them
            if not p.requires_grad:
                flat[o:o + p.numel()].zero_()
        self._last_flat_grad = flat
        return [flat[offs[id(p)]:offs[id(p)] + p.numel()].view_as(p) if p.requires_grad else None for p in params]

    def _unit_ranges(self, params, starts, total):
        """Flat-buffer range [lo, hi) of every gradient unit of dfv_train_bwd, in completion order (include/dfvit.h
        grad_events): unit 0 = head conv + attention + classifier, unit 1 + j = block 31 - j, unit 33 = stem."""
        names = [n for n, _ in self.named_parameters()]
        first = {}
        for n, o in zip(names, starts):
            if "._blocks." in n:
                blk = int(n.split("._blocks.")[1].split(".")[0])
                first.setdefault(("block", blk), o)
            elif "_conv_head" in n:
                first.setdefault(("head",), o)
        nblk = lib.dfv_b4_num_blocks()
        bounds = [first[("block", i)] for i in range(nblk)] + [first[("head",)]]
        units = [(bounds[nblk], total)]
        for j in range(nblk):
            i = nblk - 1 - j
            units.append((bounds[i], bounds[i + 1]))
        units.append((0, bounds[0]))
        assert len(units) == _lib.GRAD_UNITS
        return units

    def _train_features(self, images, landmarks, attention=True, want_taps=False):
        """Train-mode features through the same autograd node as forward() (used by the sub-modules' side APIs)."""
        self._check_modes()
        needs_bwd = torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters())
        params = [p for _, p in self.named_parameters()]
        opts = dict(needs_bwd=needs_bwd, attention=attention, want_taps=want_taps)
        _, feats = _TrainFn.apply(self, images, landmarks, opts, *params)
        return feats, self._last_taps

    def _tap_tensors(self, B, H, W, dev, dtype):
        bb = self.feature_extractor.backbone.backbone
        shapes = [(B, (H - 2) // 2 + 1, (W - 2) // 2 + 1, lib.dfv_b4_stem_channels())]
        h, w = shapes[0][1], shapes[0][2]
        for blk in bb._blocks:
            i = blk.info
            h = (h + i["pad_lo"] + i["pad_hi"] - i["kernel"]) // i["stride"] + 1
            w = (w + i["pad_lo"] + i["pad_hi"] - i["kernel"]) // i["stride"] + 1
            shapes.append((B, h, w, i["c_out"]))
        shapes.append((B, h, w, bb.head_channels))
        return [torch.empty(s, device=dev, dtype=dtype) for s in shapes]

    # ------------------------------------------------------------------ reference API
    def forward(self, images: torch.Tensor, landmarks: Optional[torch.Tensor] = None,
                return_features: bool = False) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
        """images: (B, 3, H, W) fp32 normalised crops (the reference's contract, src/data/dataset.py:82-116) or
        (B, H, W, 3) uint8 RGB crops, normalised inside the stem kernel.

        Mixed precision (trainer.py:137-141 runs the model under `torch.cuda.amp.autocast()`): inside an autocast region
        the call computes on the bf16 tensor-core path whatever `set_compute_dtype` says -- bf16 is this library's reduced
        precision (fp16's range is why the reference needs its GradScaler; the scaler protocol is honoured all the same:
        the scale arrives through the loss gradient, `unscale_` / `clip_grad_norm_` / `scaler.step` see ordinary fp32
        `.grad` tensors).  Logits and features are fp32 either way."""
        if torch.is_autocast_enabled("cuda") and self.compute_dtype != torch.bfloat16:
            prev, self.compute_dtype = self.compute_dtype, torch.bfloat16
            try:
                return self.forward(images, landmarks, return_features)
            finally:
                self.compute_dtype = prev
        self._check_modes()
        if self.training:
            # grad mode is read HERE: inside autograd.Function.forward it is always off
            needs_bwd = torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters())
            if self._detached_anchor is not None and needs_bwd:
                # inside GraphedTrainStep: the parameters are NOT autograd inputs (their AccumulateGrad nodes may live on another
                # stream, which breaks stream capture); the step attaches the gradient views to p.grad itself
                logits, feats = _TrainFn.apply(self, images, landmarks, dict(needs_bwd=True, detached_params=True), self._detached_anchor)
            else:
                params = [p for _, p in self.named_parameters()]
                logits, feats = _TrainFn.apply(self, images, landmarks, dict(needs_bwd=needs_bwd), *params)
            return (logits, feats) if return_features else (logits, None)
        logits, feats, _, _ = self._infer(images, landmarks)
        return (logits, feats) if return_features else (logits, None)

    def predict(self, images, landmarks=None, return_probs=True):
        with torch.no_grad():       # feature_extractor.py:287
            logits, _ = self.forward(images, landmarks)
            return torch.softmax(logits, dim=1) if return_probs else logits

    @torch.no_grad()
    def score_clips(self, images, landmarks=None, frames_per_clip: int = 32, threshold: float = 0.5):
        """Video-frame scoring (BASELINE.json configs[3]; task.ipynb:434-442 calls the model once per file): the batch
        holds whole clips of `frames_per_clip` consecutive frames; the landmark heat-map normaliser group is the clip
        (== one reference call per clip).  Returns {'mean_logits' (n_clips, 2), 'fake_prob' (n_clips,) = mean
        softmax[:, 1], 'labels' (n_clips,) = fake_prob >= threshold}."""
        assert not self.training and images.shape[0] % frames_per_clip == 0
        prev = self.landmark_max_group
        self.landmark_max_group = frames_per_clip
        try:
            logits, _, _, _ = self._infer(images, landmarks)
        finally:
            self.landmark_max_group = prev
        mean_logits, prob, labels = ops.clip_aggregate(logits.contiguous(), frames_per_clip, threshold)
        return {"mean_logits": mean_logits, "fake_prob": prob, "labels": labels}

    @torch.no_grad()
    def forward_with_taps(self, images, landmarks=None):
        """Debug/parity: (logits, features, heat (B,1,h,w)|None, [stem, block0..31, head] NHWC tensors)."""
        logits, feats, heat, taps = self._infer(images, landmarks, want_heat=True, taps=True)
        return logits, feats, (heat.unsqueeze(1) if heat is not None else None), taps

    def set_compute_dtype(self, dtype):
        dtype = {"bf16": torch.bfloat16, "fp32": torch.float32}.get(dtype, dtype)
        ops.dtype_code(dtype)
        self.compute_dtype = dtype
        return self


def _mix_seed(seed: int) -> int:
    return (seed ^ (_rank() * 0x9E3779B97F4A7C15)) & (2 ** 63 - 1)


def _rank() -> int:
    import torch.distributed as dist
    return dist.get_rank() if (dist.is_available() and dist.is_initialized()) else 0
