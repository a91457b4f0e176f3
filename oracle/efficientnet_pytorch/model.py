"""TEST INFRASTRUCTURE ONLY (see package docstring).

``EfficientNet`` / ``MBConvBlock`` of efficientnet-pytorch 0.7.1 restated from the
published algorithm.  Construction order of the
sub-modules follows upstream so that a seeded default init consumes the torch RNG
in the same order and the ``state_dict`` key order is identical.
"""
from torch import nn
from torch.nn import functional as F

from .utils import (
    Conv2dStaticSamePadding,
    MemoryEfficientSwish,
    Swish,
    calculate_output_image_size,
    drop_connect,
    get_model_params,
    round_filters,
    round_repeats,
)

VALID_MODELS = tuple(f"efficientnet-b{i}" for i in range(8))


def _conv(image_size):
    def make(*a, **kw):
        return Conv2dStaticSamePadding(*a, image_size=image_size, **kw)
    return make


class MBConvBlock(nn.Module):
    """expand 1x1 -> depthwise kxk -> squeeze-excite -> project 1x1 (+ skip)."""

    def __init__(self, block_args, global_params, image_size=None):
        super().__init__()
        self._block_args = block_args
        mom = 1 - global_params.batch_norm_momentum
        eps = global_params.batch_norm_epsilon
        self._bn_mom, self._bn_eps = mom, eps
        self.has_se = block_args.se_ratio is not None and 0 < block_args.se_ratio <= 1
        self.id_skip = block_args.id_skip

        cin = block_args.input_filters
        cmid = cin * block_args.expand_ratio
        if block_args.expand_ratio != 1:
            self._expand_conv = _conv(image_size)(cin, cmid, kernel_size=1, bias=False)
            self._bn0 = nn.BatchNorm2d(cmid, momentum=mom, eps=eps)

        k, s = block_args.kernel_size, block_args.stride
        self._depthwise_conv = _conv(image_size)(cmid, cmid, groups=cmid, kernel_size=k, stride=s, bias=False)
        self._bn1 = nn.BatchNorm2d(cmid, momentum=mom, eps=eps)
        image_size = calculate_output_image_size(image_size, s)

        if self.has_se:
            sq = max(1, int(cin * block_args.se_ratio))
            self._se_reduce = _conv((1, 1))(cmid, sq, kernel_size=1)
            self._se_expand = _conv((1, 1))(sq, cmid, kernel_size=1)

        cout = block_args.output_filters
        self._project_conv = _conv(image_size)(cmid, cout, kernel_size=1, bias=False)
        self._bn2 = nn.BatchNorm2d(cout, momentum=mom, eps=eps)
        self._swish = MemoryEfficientSwish()

    def forward(self, inputs, drop_connect_rate=None):
        x = inputs
        if self._block_args.expand_ratio != 1:
            x = self._swish(self._bn0(self._expand_conv(x)))
        x = self._swish(self._bn1(self._depthwise_conv(x)))
        if self.has_se:
            g = F.adaptive_avg_pool2d(x, 1)
            g = self._se_expand(self._swish(self._se_reduce(g)))
            x = g.sigmoid() * x
        x = self._bn2(self._project_conv(x))
        a = self._block_args
        if self.id_skip and a.stride == 1 and a.input_filters == a.output_filters:
            if drop_connect_rate:
                x = drop_connect(x, p=drop_connect_rate, training=self.training)
            x = x + inputs
        return x

    def set_swish(self, memory_efficient=True):
        self._swish = MemoryEfficientSwish() if memory_efficient else Swish()


class EfficientNet(nn.Module):
    def __init__(self, blocks_args=None, global_params=None):
        super().__init__()
        assert blocks_args, "blocks_args must be a non-empty list"
        self._global_params = gp = global_params
        self._blocks_args = blocks_args
        mom = 1 - gp.batch_norm_momentum
        eps = gp.batch_norm_epsilon

        size = gp.image_size
        c_stem = round_filters(32, gp)
        self._conv_stem = _conv(size)(3, c_stem, kernel_size=3, stride=2, bias=False)
        self._bn0 = nn.BatchNorm2d(c_stem, momentum=mom, eps=eps)
        size = calculate_output_image_size(size, 2)

        self._blocks = nn.ModuleList()
        for args in blocks_args:
            args = args._replace(
                input_filters=round_filters(args.input_filters, gp),
                output_filters=round_filters(args.output_filters, gp),
                num_repeat=round_repeats(args.num_repeat, gp),
            )
            self._blocks.append(MBConvBlock(args, gp, image_size=size))
            size = calculate_output_image_size(size, args.stride)
            if args.num_repeat > 1:
                args = args._replace(input_filters=args.output_filters, stride=1)
            for _ in range(args.num_repeat - 1):
                self._blocks.append(MBConvBlock(args, gp, image_size=size))

        c_head = round_filters(1280, gp)
        self._conv_head = _conv(size)(args.output_filters, c_head, kernel_size=1, bias=False)
        self._bn1 = nn.BatchNorm2d(c_head, momentum=mom, eps=eps)
        self._avg_pooling = nn.AdaptiveAvgPool2d(1)
        if gp.include_top:
            self._dropout = nn.Dropout(gp.dropout_rate)
            self._fc = nn.Linear(c_head, gp.num_classes)
        self._swish = MemoryEfficientSwish()

    def set_swish(self, memory_efficient=True):
        self._swish = MemoryEfficientSwish() if memory_efficient else Swish()
        for b in self._blocks:
            b.set_swish(memory_efficient)

    def extract_features(self, inputs):
        x = self._swish(self._bn0(self._conv_stem(inputs)))
        n = len(self._blocks)
        for idx, block in enumerate(self._blocks):
            rate = self._global_params.drop_connect_rate
            if rate:
                rate *= float(idx) / n
            x = block(x, drop_connect_rate=rate)
        return self._swish(self._bn1(self._conv_head(x)))

    def forward(self, inputs):
        x = self._avg_pooling(self.extract_features(inputs))
        if self._global_params.include_top:
            x = self._fc(self._dropout(x.flatten(start_dim=1)))
        return x

    @classmethod
    def from_name(cls, model_name, in_channels=3, **override_params):
        if model_name not in VALID_MODELS:
            raise ValueError("model_name should be one of: " + ", ".join(VALID_MODELS))
        blocks_args, gp = get_model_params(model_name, override_params)
        model = cls(blocks_args, gp)
        if in_channels != 3:
            c = round_filters(32, gp)
            model._conv_stem = _conv(gp.image_size)(in_channels, c, kernel_size=3, stride=2, bias=False)
        return model
