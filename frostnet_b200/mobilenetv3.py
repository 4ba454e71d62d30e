"""MobileNetV3 building blocks with the reference's names, module trees and state_dict keys
(Classification/models/imagenet/mobilenetv3.py:6-160): ``_ConvBNReLU`` (``cbr``), ``_ConvBN`` (``cb``),
``_Hswish``, ``_Hsigmoid``, ``_ConvBNHswish``, ``SEModule``, ``Identity`` and the inverted-residual ``Bottleneck``.

Float modules until ``fuse_model()`` + ``attach_fake_quant`` (this package's prepare_qat for arbitrary trees); afterwards every
member runs on the device through the per-module executor (block_engine.py): fused ConvBn(ReLU)2d on the tensor-core / depthwise
kernels, hard-swish as table passes, SE as in se.py, the residual through FloatFunctional.add - each boundary hands its
quantisation grid on with the tensor.  The plain ``nn.ReLU`` of the 'RE' blocks gets no observer from prepare_qat
(it is not in torch's propagation list): it is a ReLU on the incoming grid, and keeps that grid.

Not here: the MobileNetV3 network class itself (its dense 3x3 stem and biased 1x1 head convs have no stand-alone kernels yet).
"""
import torch
import torch.nn.functional as F
from torch import nn

from . import _lib as L
from . import qat as Q
from .frostnet import _ConvBlock
from .hswish import Hsigmoid as _Hsigmoid, Hswish as _Hswish
from .se import SEModule


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
