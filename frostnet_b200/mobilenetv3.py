"""MobileNetV3 building blocks with the reference's names, module trees and state_dict keys
(Classification/models/imagenet/mobilenetv3.py:6-160): ``_ConvBNReLU`` (``cbr``), ``_ConvBN`` (``cb``),
``_Hswish``, ``_Hsigmoid``, ``_ConvBNHswish``, ``SEModule``, ``Identity`` and the inverted-residual ``Bottleneck``.

Float modules until ``fuse_model()`` + ``attach_fake_quant`` (this package's prepare_qat for arbitrary trees); afterwards every
member runs on the device through the per-module executor (block_engine.py): fused ConvBn(ReLU)2d on the tensor-core / depthwise
kernels, hard-swish as table passes, SE as in se.py, the residual through FloatFunctional.add - each boundary hands its
quantisation grid on with the tensor.  The plain ``nn.ReLU`` of the 'RE' blocks gets no observer from prepare_qat
(it is not in torch's propagation list): it is a ReLU on the incoming grid, and keeps that grid.

``MobileNetV3`` (mobilenetv3.py:162-383) wires them with the stem (stand-alone dense 3x3 on the direct-convolution kernels), the
dropout on the last feature map and the pooled, biased 1x1 head convs (se.QATConv1x1).  ``dilated=True`` (layer4 at dilation 2, no
classifier - a segmentation backbone, like the reference's) runs its depthwise convs on csrc/dw_dilated.cu.
"""
import torch
import torch.nn.functional as F
from torch import nn

from . import _lib as L
from . import qat as Q
from .frostnet import _ConvBlock
from .hswish import Hsigmoid as _Hsigmoid, Hswish as _Hswish
from .se import AvgPool, QATConv1x1, SEModule, dropout


class _ConvBNReLU(_ConvBlock):
    """mobilenetv3.py:6-25 (the ``relu6=True`` variant is never instantiated by the reference's networks)."""
    _relu = True
    _seq_name = "cbr"

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1, relu6=False,
                 norm_layer=nn.BatchNorm2d, **kwargs):
        if relu6:
            raise ValueError("frostnet_b200: _ConvBNReLU(relu6=True) is not supported (unused by the reference's MobileNetV3)")
        super().__init__(in_channels, out_channels, kernel_size, stride, padding, dilation, groups)
        self.relu6 = relu6


class _ConvBN(_ConvBlock):
    """mobilenetv3.py:27-41."""
    _relu = False
    _seq_name = "cb"

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1,
                 norm_layer=nn.BatchNorm2d, **kwargs):
        super().__init__(in_channels, out_channels, kernel_size, stride, padding, dilation, groups)


class _ConvBNHswish(nn.Module):
    """mobilenetv3.py:72-83."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1,
                 norm_layer=nn.BatchNorm2d, **kwargs):
        super().__init__()
        self.cb = _ConvBN(in_channels, out_channels, kernel_size, stride, padding, dilation, groups)
        self.act = _Hswish(True)

    def forward(self, x):
        return self.act(self.cb(x))

    def fuse_model(self):
        self.cb.fuse_model()


class Identity(nn.Module):
    """mobilenetv3.py:104-110."""

    def __init__(self, in_channels):
        super().__init__()

    def forward(self, x):
        return x


class _ReluFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        x = x.detach()
        if x.dtype != torch.float32 or not x.is_contiguous():
            x = x.float().contiguous()
        y = torch.empty_like(x)
        mask = torch.empty(x.shape, dtype=torch.uint8, device=x.device)
        with torch.cuda.device(x.device):
            L.call("frost_relu_forward", x.data_ptr(), x.numel(), y.data_ptr(), mask.data_ptr(), L.stream(x.device))
        ctx.mask = mask
        return y

    @staticmethod
    def backward(ctx, dy):
        dy = dy.contiguous().float()
        dx = torch.empty_like(dy)
        with torch.cuda.device(dy.device):
            L.call("frost_fq_backward", dy.data_ptr(), ctx.mask.data_ptr(), dy.numel(), dx.data_ptr(), L.stream(dy.device))
        return dx


class ReLU(nn.ReLU):
    """The 'RE' blocks' ``act(True)`` = nn.ReLU(inplace=True) (mobilenetv3.py:119-120, 134): no parameters, no observer.  On a
    tensor that carries a quantisation grid it runs on the device (frost_relu_forward) and hands the same grid on."""

    def forward(self, x):
        qp = getattr(x, "_frost_qparams", None)
        if qp is None or not x.is_cuda:
            return F.relu(x)
        from .block_engine import attach_qparams
        return attach_qparams(_ReluFunction.apply(x), *qp)


class Bottleneck(nn.Module):
    """mobilenetv3.py:113-160."""

    def __init__(self, in_channels, out_channels, exp_size, kernel_size, stride, dilation=1, se=False, nl='RE',
                 norm_layer=nn.BatchNorm2d, **kwargs):
        super().__init__()
        assert stride in [1, 2]
        self.use_res_connect = stride == 1 and in_channels == out_channels
        act = _Hswish if nl == 'HS' else ReLU
        SELayer = SEModule if se else Identity
        self.conv = nn.Sequential(
            # pw
            _ConvBNHswish(in_channels, exp_size, 1) if nl == 'HS' else _ConvBNReLU(in_channels, exp_size, 1),
            # dw
            _ConvBN(exp_size, exp_size, kernel_size, stride, (kernel_size - 1) // 2 * dilation, dilation, groups=exp_size),
            SELayer(exp_size),
            act(True),
            # pw-linear
            _ConvBN(exp_size, out_channels, 1)
        )
        self.se = se
        if self.use_res_connect:
            self.skip_add = Q.FloatFunctional()

    def forward(self, x):
        if self.use_res_connect:
            return self.skip_add.add(x, self.conv(x))
        return self.conv(x)

    def fuse_model(self):
        self.conv[0].fuse_model()
        self.conv[1].fuse_model()
        if self.se:
            self.conv[2].fuse_model()
        self.conv[4].fuse_model()


# (kernel, exp_size, out_channels, se, non-linearity, stride) per block - mobilenetv3.py:166-288; RE=True swaps every 'HS' for 'RE'
_LARGE = ([(3, 16, 16, False, 'RE', 1), (3, 64, 24, False, 'RE', 2), (3, 72, 24, False, 'RE', 1)],
          [(5, 72, 40, True, 'RE', 2), (5, 120, 40, True, 'RE', 1), (5, 120, 40, True, 'RE', 1)],
          [(3, 240, 80, False, 'HS', 2), (3, 200, 80, False, 'HS', 1), (3, 184, 80, False, 'HS', 1), (3, 184, 80, False, 'HS', 1),
           (3, 480, 112, True, 'HS', 1), (3, 672, 112, True, 'HS', 1)],
          [(5, 672, 160, True, 'HS', 2), (5, 960, 160, True, 'HS', 1), (5, 960, 160, True, 'HS', 1)])
_SMALL = ([(3, 16, 16, True, 'RE', 2)],
          [(3, 72, 24, False, 'RE', 2), (3, 88, 24, False, 'RE', 1)],
          [(5, 96, 40, True, 'HS', 2), (5, 240, 40, True, 'HS', 1), (5, 240, 40, True, 'HS', 1), (5, 120, 48, True, 'HS', 1),
           (5, 144, 48, True, 'HS', 1)],
          [(5, 288, 96, True, 'HS', 2), (5, 576, 96, True, 'HS', 1), (5, 576, 96, True, 'HS', 1)])


class MobileNetV3(nn.Module):
    """mobilenetv3.py:162-383: same constructor, attribute names, state_dict keys, forward and fuse_model()."""

    def __init__(self, nclass=1000, mode='large', width_mult=1.0, dilated=False, norm_layer=nn.BatchNorm2d, RE=False, **kwargs):
        super().__init__()
        if mode not in ('large', 'small'):
            raise ValueError('Unknown mode.')
        settings = [list(map(list, stage)) for stage in (_LARGE if mode == 'large' else _SMALL)]
        if dilated:                                   # the last block of layer4 at half width (mobilenetv3.py:183-187, 212-216)
            settings[3][2][1] //= 2
            settings[3][2][2] //= 2
        if RE:
            for stage in settings:
                for blk in stage:
                    blk[4] = 'RE'
        self.in_channels = int(16 * width_mult) if width_mult > 1.0 else 16
        first = _ConvBNReLU if RE else _ConvBNHswish
        self.conv1 = first(3, self.in_channels, 3, 2, 1, norm_layer=norm_layer)
        self.layer1 = self._make_layer(Bottleneck, settings[0], width_mult, norm_layer=norm_layer)
        self.layer2 = self._make_layer(Bottleneck, settings[1], width_mult, norm_layer=norm_layer)
        self.layer3 = self._make_layer(Bottleneck, settings[2], width_mult, norm_layer=norm_layer)
        self.layer4 = self._make_layer(Bottleneck, settings[3], width_mult, dilation=2 if dilated else 1, norm_layer=norm_layer)
        base = 960 if mode == 'large' else 576
        if dilated:
            base //= 2
        last_bneck_channels = int(base * width_mult) if width_mult > 1.0 else base
        self.layer5 = first(self.in_channels, last_bneck_channels, 1, norm_layer=norm_layer)
        if not dilated:
            hidden = 1280 if mode == 'large' else 1024
            head = [SEModule(last_bneck_channels)] if mode == 'small' else []
            head += [AvgPool(1), QATConv1x1(last_bneck_channels, hidden), _Hswish(True), QATConv1x1(hidden, nclass)]
            self.classifier = nn.Sequential(*head)
        self.mode = mode
        self.dilated = dilated
        self.drop_rate = 0.8                          # F.dropout(x, p=0.8) hard-wired in the reference's forward (:353)
        self.quant = Q.QuantStub()
        self.dequant = Q.DeQuantStub()
        self._init_weights()

    def _make_layer(self, block, block_setting, width_mult, dilation=1, norm_layer=nn.BatchNorm2d):
        layers = []
        for k, exp_size, c, se, nl, s in block_setting:
            out_channels = int(c * width_mult)
            stride = s if dilation == 1 else 1
            layers.append(block(self.in_channels, out_channels, int(exp_size * width_mult), k, stride, dilation, se, nl, norm_layer))
            self.in_channels = out_channels
        return nn.Sequential(*layers)

    def forward(self, x):
        x = self.quant(x)
        x = self.conv1(x)
        x = self.layer1(x)
        x = self.layer2(x)
        x = self.layer3(x)
        x = self.layer4(x)
        x = self.layer5(x)
        x = dropout(x, self.drop_rate, self.training)
        x = self.classifier(x)
        x = self.dequant(x)
        return x.view(x.size(0), x.size(1))

    def fuse_model(self):
        self.conv1.fuse_model()
        for stage in (self.layer1, self.layer2, self.layer3, self.layer4):
            for layer in stage:
                layer.fuse_model()
        self.layer5.fuse_model()
        if not self.dilated and self.mode == 'small':
            self.classifier[0].fuse_model()

    def _init_weights(self):
        """mobilenetv3.py:372-383 (the SE Linear layers are nn.Linear there: normal(0, 0.01))."""
        from .se import QATLinear
        for m in self.modules():
            if isinstance(m, (nn.Conv2d, QATConv1x1)):
                nn.init.kaiming_normal_(m.weight, mode='fan_out')
                if m.bias is not None:
                    nn.init.zeros_(m.bias)
            elif isinstance(m, nn.BatchNorm2d):
                nn.init.ones_(m.weight)
                nn.init.zeros_(m.bias)
            elif isinstance(m, QATLinear):
                nn.init.normal_(m.weight, 0, 0.01)


def get_mobilenet_v3(mode='small', width_mult=1.0, RE=False, pretrained=False, root='~/,torch/models', **kwargs):
    if pretrained:
        raise ValueError("Not support pretrained")
    return MobileNetV3(mode=mode, width_mult=width_mult, RE=RE, **kwargs)


def mobilenet_v3_large(**kwargs):
    return get_mobilenet_v3('large', 1.0, **kwargs)


def mobilenet_v3_small(**kwargs):
    return get_mobilenet_v3('small', 1.0, **kwargs)


def mobilenet_v3_ReLU_large(**kwargs):
    return get_mobilenet_v3('large', 1.0, RE=True, **kwargs)


def mobilenet_v3_ReLU_small(**kwargs):
    return get_mobilenet_v3('small', 1.0, RE=True, **kwargs)
