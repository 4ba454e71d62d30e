"""MobileNetV2 building blocks of the SSDLite backbone with the reference's names, module trees and state_dict keys
(Object_Detection/ssd_qmv2.py:40-110): ``ConvBNReLU`` - an nn.Sequential of Conv2d, BatchNorm2d, ReLU - and
``InvertedResidual``, whose linear tail is a bare Conv2d + BatchNorm2d inside ``self.conv``.  ``fuse_model`` applies the rule of
the reference's ``MobileNetV2.fuse_model`` (:178-185) with this package's ``fuse_modules``; after ``attach_fake_quant`` every
fused conv is a prepared ``FrostConvBn2d`` that runs, called by the nn.Sequential around it, on the per-module executor.

The dilated depthwise convs of the backbone's last two stages (d = 2) run on the gather kernels of csrc/dw_dilated.cu.
``MobileNetV2`` is the backbone as the reference builds it (:112-185: ``features`` only, no classifier).  The SSD heads, extras
and MultiBox loss are not built.
"""
from torch import nn

from . import qat as Q


class ConvBNReLU(nn.Sequential):
    """ssd_qmv2.py:40-52."""

    def __init__(self, in_planes, out_planes, kernel_size=3, stride=1, groups=1, dilation=1):
        padding = dilation if dilation > 1 else (kernel_size - 1) // 2
        super().__init__(nn.Conv2d(in_planes, out_planes, kernel_size, stride, padding, dilation=dilation, groups=groups, bias=False),
                         nn.BatchNorm2d(out_planes, momentum=0.1), nn.ReLU(inplace=False))


class InvertedResidual(nn.Module):
    """ssd_qmv2.py:80-110."""

    def __init__(self, inp, oup, stride, dilation, expand_ratio):
        super().__init__()
        self.stride = stride
        assert stride in [1, 2]
        hidden_dim = int(round(inp * expand_ratio))
        self.use_res_connect = self.stride == 1 and inp == oup
        layers = []
        if expand_ratio != 1:
            layers.append(ConvBNReLU(inp, hidden_dim, kernel_size=1))                                   # pw
        layers.extend([ConvBNReLU(hidden_dim, hidden_dim, stride=stride, dilation=dilation, groups=hidden_dim),    # dw
                       nn.Conv2d(hidden_dim, oup, 1, 1, 0, bias=False), nn.BatchNorm2d(oup, momentum=0.1)])          # pw-linear
        self.conv = nn.Sequential(*layers)
        if self.use_res_connect:
            self.skip_add = Q.FloatFunctional()

    def forward(self, x):
        if self.use_res_connect:
            return self.skip_add.add(x, self.conv(x))
        return self.conv(x)


def _make_divisible(v, divisor, min_value=None):
    """ssd_qmv2.py:19-37."""
    if min_value is None:
        min_value = divisor
    new_v = max(min_value, int(v + divisor / 2) // divisor * divisor)
    if new_v < 0.9 * v:
        new_v += divisor
    return new_v


# t (expand), c, n (repeats), s (stride of the first), d (dilation) - ssd_qmv2.py:129-138
_SETTING = [[1, 16, 1, 1, 1], [6, 24, 2, 2, 1], [6, 32, 3, 2, 1], [6, 64, 4, 2, 1], [6, 96, 3, 1, 1], [6, 160, 3, 1, 2], [6, 320, 1, 1, 2]]


class MobileNetV2(nn.Module):
    """ssd_qmv2.py:112-185: the SSDLite backbone (``features``; output stride 16, the last two stages dilated)."""

    def __init__(self, num_classes=1000, width_mult=1.0, inverted_residual_setting=None, round_nearest=8):
        super().__init__()
        setting = _SETTING if inverted_residual_setting is None else inverted_residual_setting
        if len(setting) == 0 or len(setting[0]) != 5:
            raise ValueError("inverted_residual_setting should be non-empty or a 5-element list, got {}".format(setting))
        input_channel = _make_divisible(32 * width_mult, round_nearest)
        self.last_channel = _make_divisible(1280 * max(1.0, width_mult), round_nearest)
        features = [ConvBNReLU(3, input_channel, stride=2)]
        for t, c, n, s, d in setting:
            output_channel = _make_divisible(c * width_mult, round_nearest)
            for i in range(n):
                features.append(InvertedResidual(input_channel, output_channel, s if i == 0 else 1, dilation=d, expand_ratio=t))
                input_channel = output_channel
        features.append(ConvBNReLU(input_channel, self.last_channel, kernel_size=1))
        self.features = nn.Sequential(*features)
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode='fan_out')
            elif isinstance(m, nn.BatchNorm2d):
                nn.init.ones_(m.weight)
                nn.init.zeros_(m.bias)

    def forward(self, x):
        return self.features(x)

    def fuse_model(self):
        fuse_model(self)


def fuse_model(root):
    """MobileNetV2.fuse_model (ssd_qmv2.py:178-185) for any tree of these blocks."""
    for m in list(root.modules()):
        if type(m) == ConvBNReLU:
            Q.fuse_modules(m, ['0', '1', '2'], inplace=True)
        if type(m) == InvertedResidual:
            for idx in range(len(m.conv)):
                if type(m.conv[idx]) == nn.Conv2d:
                    Q.fuse_modules(m.conv, [str(idx), str(idx + 1)], inplace=True)
    return root
