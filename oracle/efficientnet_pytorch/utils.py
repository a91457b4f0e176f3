"""TEST INFRASTRUCTURE ONLY (see package docstring).

Helpers of efficientnet-pytorch 0.7.1 ``utils.py`` restated from its published
behaviour.  Only what the B4 ``from_name`` path
touches is present.
"""
import math
from dataclasses import dataclass, replace
from typing import List, Optional, Tuple

import torch
from torch import nn
from torch.nn import functional as F


@dataclass(frozen=True)
class GlobalParams:
    width_coefficient: float
    depth_coefficient: float
    image_size: int
    dropout_rate: float
    num_classes: int = 1000
    batch_norm_momentum: float = 0.99  # TF convention; torch momentum = 1 - this
    batch_norm_epsilon: float = 1e-3
    drop_connect_rate: float = 0.2
    depth_divisor: int = 8
    min_depth: Optional[int] = None
    include_top: bool = True

    def _replace(self, **kw):   # namedtuple-style, as upstream's GlobalParams has it
        return replace(self, **kw)


@dataclass(frozen=True)
class BlockArgs:
    num_repeat: int
    kernel_size: int
    stride: int
    expand_ratio: int
    input_filters: int
    output_filters: int
    se_ratio: Optional[float]
    id_skip: bool = True

    def _replace(self, **kw):
        return replace(self, **kw)


# (width, depth, resolution, dropout) -- upstream ``efficientnet_params``.
_COMPOUND = {
    "efficientnet-b0": (1.0, 1.0, 224, 0.2),
    "efficientnet-b1": (1.0, 1.1, 240, 0.2),
    "efficientnet-b2": (1.1, 1.2, 260, 0.3),
    "efficientnet-b3": (1.2, 1.4, 300, 0.3),
    "efficientnet-b4": (1.4, 1.8, 380, 0.4),
    "efficientnet-b5": (1.6, 2.2, 456, 0.4),
    "efficientnet-b6": (1.8, 2.6, 528, 0.5),
    "efficientnet-b7": (2.0, 3.1, 600, 0.5),
}

_BASE_BLOCKS = [
    "r1_k3_s11_e1_i32_o16_se0.25",
    "r2_k3_s22_e6_i16_o24_se0.25",
    "r2_k5_s22_e6_i24_o40_se0.25",
    "r3_k3_s22_e6_i40_o80_se0.25",
    "r3_k5_s11_e6_i80_o112_se0.25",
    "r4_k5_s22_e6_i112_o192_se0.25",
    "r1_k3_s11_e6_i192_o320_se0.25",
]


def decode_block_string(s: str) -> BlockArgs:
    fields = {}
    for tok in s.split("_"):
        if tok == "noskip":
            continue
        key = tok[0] if not tok.startswith("se") else "se"
        fields[key] = tok[len(key):]
    return BlockArgs(
        num_repeat=int(fields["r"]),
        kernel_size=int(fields["k"]),
        stride=int(fields["s"][0]),
        expand_ratio=int(fields["e"]),
        input_filters=int(fields["i"]),
        output_filters=int(fields["o"]),
        se_ratio=float(fields["se"]) if "se" in fields else None,
        id_skip="noskip" not in s,
    )


def get_model_params(model_name: str, override: dict) -> Tuple[List[BlockArgs], GlobalParams]:
    w, d, res, p = _COMPOUND[model_name]
    gp = GlobalParams(width_coefficient=w, depth_coefficient=d, image_size=res, dropout_rate=p)
    if override:
        gp = replace(gp, **override)
    return [decode_block_string(s) for s in _BASE_BLOCKS], gp


def round_filters(filters: int, gp: GlobalParams) -> int:
    """Width scaling rounded to ``depth_divisor`` (never dropping below 90 %)."""
    if not gp.width_coefficient:
        return filters
    div = gp.depth_divisor
    scaled = filters * gp.width_coefficient
    floor_ = gp.min_depth or div
    new = max(floor_, int(scaled + div / 2) // div * div)
    if new < 0.9 * scaled:
        new += div
    return int(new)


def round_repeats(repeats: int, gp: GlobalParams) -> int:
    if not gp.depth_coefficient:
        return repeats
    return int(math.ceil(gp.depth_coefficient * repeats))


def drop_connect(inputs: torch.Tensor, p: float, training: bool) -> torch.Tensor:
    """Per-sample stochastic depth: x / keep * floor(keep + U[0,1))."""
    assert 0 <= p <= 1
    if not training:
        return inputs
    keep = 1 - p
    rnd = keep
    rnd = rnd + torch.rand([inputs.shape[0], 1, 1, 1], dtype=inputs.dtype, device=inputs.device)
    return inputs / keep * torch.floor(rnd)


def _as_hw(x) -> Tuple[int, int]:
    if isinstance(x, int):
        return x, x
    if isinstance(x, (list, tuple)):
        return (x[0], x[0]) if len(x) == 1 else (x[0], x[1])
    raise TypeError(type(x))


def calculate_output_image_size(image_size, stride):
    if image_size is None:
        return None
    h, w = _as_hw(image_size)
    s = stride if isinstance(stride, int) else stride[0]
    return [int(math.ceil(h / s)), int(math.ceil(w / s))]


class Conv2dStaticSamePadding(nn.Conv2d):
    """Conv2d whose TF-"SAME" padding is fixed at construction from ``image_size``.

    The pad is a function of the *construction-time* image size (the 380 chain for
    B4), never of the live input.  It is applied as an
    explicit ``ZeroPad2d`` followed by an unpadded convolution.
    """

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, image_size=None, **kwargs):
        super().__init__(in_channels, out_channels, kernel_size, stride, **kwargs)
        self.stride = self.stride if len(self.stride) == 2 else [self.stride[0]] * 2
        assert image_size is not None
        ih, iw = _as_hw(image_size)
        kh, kw = self.weight.size()[-2:]
        sh, sw = self.stride
        oh, ow = math.ceil(ih / sh), math.ceil(iw / sw)
        pad_h = max((oh - 1) * sh + (kh - 1) * self.dilation[0] + 1 - ih, 0)
        pad_w = max((ow - 1) * sw + (kw - 1) * self.dilation[1] + 1 - iw, 0)
        if pad_h > 0 or pad_w > 0:
            self.static_padding = nn.ZeroPad2d(
                (pad_w // 2, pad_w - pad_w // 2, pad_h // 2, pad_h - pad_h // 2)
            )
        else:
            self.static_padding = nn.Identity()

    def forward(self, x):
        x = self.static_padding(x)
        return F.conv2d(x, self.weight, self.bias, self.stride, self.padding, self.dilation, self.groups)


class _SwishFn(torch.autograd.Function):
    """x*sigmoid(x) saving only the input; backward recomputes sigmoid."""

    @staticmethod
    def forward(ctx, i):
        ctx.save_for_backward(i)
        return i * torch.sigmoid(i)

    @staticmethod
    def backward(ctx, g):
        (i,) = ctx.saved_tensors
        s = torch.sigmoid(i)
        return g * (s * (1 + i * (1 - s)))


class MemoryEfficientSwish(nn.Module):
    def forward(self, x):
        return _SwishFn.apply(x)


class Swish(nn.Module):
    def forward(self, x):
        return x * torch.sigmoid(x)
