"""MobileNetV2 building blocks of the SSDLite backbone with the reference's names, module trees and state_dict keys
(Object_Detection/ssd_qmv2.py:40-110): ``ConvBNReLU`` - an nn.Sequential of Conv2d, BatchNorm2d, ReLU - and
``InvertedResidual``, whose linear tail is a bare Conv2d + BatchNorm2d inside ``self.conv``.  ``fuse_model`` applies the rule of
the reference's ``MobileNetV2.fuse_model`` (:178-185) with this package's ``fuse_modules``; after ``attach_fake_quant`` every
fused conv is a prepared ``FrostConvBn2d`` that runs, called by the nn.Sequential around it, on the per-module executor.

Dilation 1 only: the dilated depthwise convs of the backbone's last two stages (d = 2) have no kernel yet - the constructor
accepts the argument (float model), the QAT path refuses it.  The SSD heads, extras and MultiBox loss are not built.
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
