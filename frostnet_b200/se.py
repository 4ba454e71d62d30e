"""Quantization-aware squeeze-and-excite: the reference's ``SEModule``
(Classification/models/imagenet/mobilenetv3.py:85-102) with the same attribute names and
state_dict keys (``fc.0.weight``, ``fc.0.weight_fake_quant.*``, ``fc.0.activation_post_process.*``, ``fc.2.*``,
``fc.3.relu6.activation_post_process.*``, ``quant_mul.*``).

Float behaviour before ``fuse_model()`` + ``attach_fake_quant``; afterwards the block runs on the device through the
per-module calling convention of ``block_engine`` (the input carries its grid as ``_frost_qparams``):

    avg_pool          frost_pool_dropout_forward on the input's uint8 indices (no observer: prepare_qat adds none)
    fc.0 LinearReLU   weight fake-quant (symmetric int8, frost_fq_forward) -> frost_linear_forward -> frost_relu_forward
                      -> activation fake-quant (frost_fq_forward)                       (nniqat.LinearReLU)
    fc.2 Linear       the same without the ReLU                                          (nnqat.Linear)
    fc.3 Hsigmoid     frost_hsigmoid_forward (hswish.py)
    quant_mul.mul     frost_bcast_mul_forward -> frost_fq_forward with quant_mul's observer

Every step is its own autograd node over the matching C-ABI backward entry point; torch only adds the two gradients of x.
"""
import math

import torch
import torch.nn.functional as F
from torch import nn

from . import _lib as L
from . import qat as Q
from .hswish import Hsigmoid


def _scratch(mod, dev):
    sc = mod.__dict__.get("_frost_scratch")
    if sc is None or sc.device != dev:
        sc = torch.zeros(L.FQ_SCRATCH_FLOATS, dtype=torch.float32, device=dev)
        mod.__dict__["_frost_scratch"] = sc
    return sc


class QATLinear(nn.Module):
    """``nn.Linear(bias=False)`` that becomes nniqat.LinearReLU (``relu=True`` after SEModule.fuse_model) or nnqat.Linear once
    ``attach_fake_quant`` has given it ``weight_fake_quant`` and ``activation_post_process``."""

    def __init__(self, in_features, out_features, relu=False):
        super().__init__()
        self.in_features, self.out_features, self.relu = in_features, out_features, relu
        self.weight = nn.Parameter(torch.empty(out_features, in_features))
        nn.init.kaiming_uniform_(self.weight, a=math.sqrt(5))       # nn.Linear.reset_parameters

    def _prepared(self):
        return isinstance(getattr(self, "weight_fake_quant", None), Q.FrostFakeQuantize)

    def forward(self, x):
        if self._prepared():
            return _run_linear(self, x)
        y = F.linear(x, self.weight)
        return F.relu(y) if self.relu else y

    def extra_repr(self):
        return "in_features=%d, out_features=%d, bias=False, relu=%s" % (self.in_features, self.out_features, self.relu)


class SEModule(nn.Module):
    """mobilenetv3.py:85-102."""

    def __init__(self, in_channels, reduction=4):
        super().__init__()
        self.avg_pool = nn.AdaptiveAvgPool2d(1)
        self.fc = nn.Sequential(QATLinear(in_channels, in_channels // reduction), nn.ReLU(inplace=False),
                                QATLinear(in_channels // reduction, in_channels), Hsigmoid(True))
        self.quant_mul = Q.FloatFunctional()

    def _prepared(self):
        return isinstance(getattr(self.quant_mul, "activation_post_process", None), Q.FrostFakeQuantize)

    def fuse_model(self):
        """mobilenetv3.py:101-102: fuse_modules(self.fc, ['0', '1']) - Linear + ReLU become one module, fc.1 an Identity."""
        self.fc[0].relu = True
        self.fc[1] = nn.Identity()

    def forward(self, x):
        if self._prepared():
            return _run_se(self, x)
        n, c, _, _ = x.size()
        out = self.avg_pool(x).view(n, c)
        out = self.fc(out).view(n, c, 1, 1)
        return self.quant_mul.mul(x, out.expand_as(x))


# ---------------------------------------------------------------------- autograd nodes
def _fq(fq):
    return L.FQ(fq.activation_post_process.min_val.data_ptr(), fq.activation_post_process.max_val.data_ptr(), fq.scale.data_ptr(),
                fq.zero_point.data_ptr())


class _PoolFunction(torch.autograd.Function):
    """AdaptiveAvgPool2d(1) of an fp32 NCHW tensor that sits on the grid (scale, zp) -> [N, C]."""

    @staticmethod
    def forward(ctx, x, scale, zp, mod):
        from .block_engine import _to_nhwc_u8
        dev = x.device
        N, C, H, W = x.shape
        dummy = mod.__dict__.get("_frost_dummy")
        if dummy is None or dummy.device != dev:
            dummy = torch.zeros(2, dtype=torch.float32, device=dev)
            mod.__dict__["_frost_dummy"] = dummy
        pooled = torch.empty((N, C), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            q, _ = _to_nhwc_u8(x, (scale, zp), _scratch(mod, dev), dummy)
            L.call("frost_pool_dropout_forward", q.data_ptr(), scale.data_ptr(), zp.data_ptr(), N, H * W, C, None, 1.0,
                   pooled.data_ptr(), L.stream(dev))
        ctx.shape = (N, C, H, W)
        return pooled

    @staticmethod
    def backward(ctx, dp):
        N, C, H, W = ctx.shape
        dp = dp.contiguous().float()
        dy = torch.empty((N, H, W, C), dtype=torch.float32, device=dp.device)
        with torch.cuda.device(dp.device):
            L.call("frost_pool_dropout_backward", dp.data_ptr(), N, H * W, C, None, 1.0, dy.data_ptr(), L.stream(dp.device))
        return dy.permute(0, 3, 1, 2).contiguous(), None, None, None


def _pad4(n):
    return (n + 3) & ~3


class _LinearFunction(torch.autograd.Function):
    """relu?(F.linear(x, FQ_w(W)) + bias) on frost_linear_forward; K and cout are padded to multiples of 4 for its 4-wide loads.
    ``weight`` may be [cout, K] or a 1x1 conv's [cout, K, 1, 1]."""

    @staticmethod
    def forward(ctx, x, weight, bias, mod):
        dev = x.device
        wfq = mod.weight_fake_quant
        x = x.detach()
        if x.dtype != torch.float32 or not x.is_contiguous():
            x = x.float().contiguous()
        W = weight.detach().float().contiguous().view(weight.shape[0], -1)
        cout, K = W.shape
        N = x.shape[0]
        Kp, cp = _pad4(K), _pad4(cout)
        wy = torch.empty_like(W)
        wmask = torch.empty(W.shape, dtype=torch.uint8, device=dev)
        wq32 = torch.empty(W.shape, dtype=torch.int32, device=dev)
        with torch.cuda.device(dev):
            st = L.stream(dev)
            L.call("frost_fq_forward", W.data_ptr(), W.numel(), _fq(wfq), Q.W_QMIN, Q.W_QMAX, 1, 1 if wfq._observe else 0,
                   Q.AVERAGING_CONSTANT, wy.data_ptr(), wmask.data_ptr(), wq32.data_ptr(), _scratch(mod, dev).data_ptr(), st)
            if (Kp, cp) != (K, cout):
                # padding rows / columns hold the zero point: (wq - zp) * s == 0
                wq8 = wfq.zero_point.to(torch.int8).expand(cp, Kp).contiguous()
                wq8[:cout, :K] = wq32.to(torch.int8)
                xp = torch.zeros((N, Kp), dtype=torch.float32, device=dev)
                xp[:, :K] = x
            else:
                wq8, xp = wq32.to(torch.int8), x
            out = torch.empty((N, cp), dtype=torch.float32, device=dev)
            bp = None
            if bias is not None:
                bp = bias.detach().float().contiguous()
                if cp != cout:
                    bp = torch.cat([bp, torch.zeros(cp - cout, dtype=torch.float32, device=dev)])
            L.call("frost_linear_forward", xp.data_ptr(), wq8.data_ptr(), wfq.scale.data_ptr(), wfq.zero_point.data_ptr(), L.ptr(bp),
                   N, Kp, cp, out.data_ptr(), st)
            rmask = None
            if getattr(mod, "relu", False):
                rmask = torch.empty(out.shape, dtype=torch.uint8, device=dev)
                y = torch.empty_like(out)
                L.call("frost_relu_forward", out.data_ptr(), out.numel(), y.data_ptr(), rmask.data_ptr(), st)
                out = y
        ctx.saved = (xp, wq8, wmask, rmask, wfq, (N, K, cout, Kp, cp), tuple(weight.shape), bias is not None)
        return out[:, :cout].contiguous() if cp != cout else out

    @staticmethod
    def backward(ctx, dout):
        xp, wq8, wmask, rmask, wfq, (N, K, cout, Kp, cp), wshape, has_bias = ctx.saved
        dev = dout.device
        d = dout.contiguous().float()
        if cp != cout:
            dpad = torch.zeros((N, cp), dtype=torch.float32, device=dev)
            dpad[:, :cout] = d
            d = dpad
        with torch.cuda.device(dev):
            st = L.stream(dev)
            if rmask is not None:
                dr = torch.empty_like(d)
                L.call("frost_fq_backward", d.data_ptr(), rmask.data_ptr(), d.numel(), dr.data_ptr(), st)
                d = dr
            dx = torch.empty((N, Kp), dtype=torch.float32, device=dev)
            dwq = torch.empty((cp, Kp), dtype=torch.float32, device=dev)
            db = torch.empty(cp, dtype=torch.float32, device=dev) if has_bias else None
            L.call("frost_linear_backward", d.data_ptr(), xp.data_ptr(), wq8.data_ptr(), wfq.scale.data_ptr(),
                   wfq.zero_point.data_ptr(), N, Kp, cp, dx.data_ptr(), dwq.data_ptr(), L.ptr(db), st)
            dwq = dwq[:cout, :K].contiguous()
            dW = torch.empty_like(dwq)
            L.call("frost_fq_backward", dwq.data_ptr(), wmask.data_ptr(), dwq.numel(), dW.data_ptr(), st)
        return (dx[:, :K].contiguous() if Kp != K else dx), dW.view(wshape), (db[:cout].contiguous() if has_bias else None), None


class _BcastMulFunction(torch.autograd.Function):
    """torch.mul(x, gate.view(N, C, 1, 1).expand_as(x))"""

    @staticmethod
    def forward(ctx, x, gate):
        x = x.detach()
        if x.dtype != torch.float32 or not x.is_contiguous():
            x = x.float().contiguous()
        gate = gate.detach().float().contiguous()
        N, C, H, W = x.shape
        y = torch.empty_like(x)
        with torch.cuda.device(x.device):
            L.call("frost_bcast_mul_forward", x.data_ptr(), gate.data_ptr(), N * C, H * W, y.data_ptr(), L.stream(x.device))
        ctx.saved = (x, gate)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, gate = ctx.saved
        N, C, H, W = x.shape
        dy = dy.contiguous().float()
        dx = torch.empty_like(x)
        dg = torch.empty_like(gate)
        with torch.cuda.device(x.device):
            L.call("frost_bcast_mul_backward", dy.data_ptr(), x.data_ptr(), gate.data_ptr(), N * C, H * W, dx.data_ptr(), dg.data_ptr(),
                   L.stream(x.device))
        return dx, dg


def _run_linear(mod, x):
    from .block_engine import _FakeQuantFunction, attach_qparams
    if not x.is_cuda:
        raise RuntimeError("frostnet_b200: the QAT path runs on a CUDA device (B200) only; got a CPU tensor")
    if x.dim() != 2 or x.shape[1] != mod.in_features:
        raise RuntimeError("frostnet_b200: QATLinear expects [N, %d], got %s" % (mod.in_features, tuple(x.shape)))
    y = _LinearFunction.apply(x, mod.weight, None, mod)
    fq = mod.activation_post_process
    y = _FakeQuantFunction.apply(y, fq, _scratch(mod, x.device))
    return attach_qparams(y, fq.scale, fq.zero_point)


class QATConv1x1(nn.Module):
    """``nn.Conv2d(cin, cout, 1)`` with bias on pooled [N, cin, 1, 1] tensors (MobileNetV3's classifier convs,
    mobilenetv3.py:306-308, 320-323); nnqat.Conv2d once ``attach_fake_quant`` has given it its two fake-quants - the same
    ``weight`` / ``bias`` / ``weight_fake_quant.*`` / ``activation_post_process.*`` keys."""

    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        conv = nn.Conv2d(in_channels, out_channels, 1)
        self.weight, self.bias = conv.weight, conv.bias

    def _prepared(self):
        return isinstance(getattr(self, "weight_fake_quant", None), Q.FrostFakeQuantize)

    def forward(self, x):
        if not self._prepared():
            return F.conv2d(x, self.weight, self.bias)
        from .block_engine import _FakeQuantFunction, attach_qparams
        if not x.is_cuda:
            raise RuntimeError("frostnet_b200: the QAT path runs on a CUDA device (B200) only; got a CPU tensor")
        if x.dim() != 4 or x.shape[1] != self.in_channels or x.shape[2:] != (1, 1):
            raise RuntimeError("frostnet_b200: QATConv1x1 runs on pooled [N, %d, 1, 1] tensors, got %s" % (self.in_channels, tuple(x.shape)))
        y = _LinearFunction.apply(x.reshape(x.shape[0], -1), self.weight, self.bias, self)
        fq = self.activation_post_process
        y = _FakeQuantFunction.apply(y, fq, _scratch(self, x.device))
        return attach_qparams(y.view(y.shape[0], -1, 1, 1), fq.scale, fq.zero_point)

    def extra_repr(self):
        return "%d, %d, kernel_size=(1, 1), stride=(1, 1)" % (self.in_channels, self.out_channels)


class AvgPool(nn.AdaptiveAvgPool2d):
    """``nn.AdaptiveAvgPool2d(1)`` that, on a tensor carrying a quantisation grid, pools on the device (no observer: prepare_qat
    adds none to a pooling module, its output is off-grid fp32)."""

    def forward(self, x):
        qp = getattr(x, "_frost_qparams", None)
        if qp is None or not x.is_cuda:
            return super().forward(x)
        n, c = x.shape[:2]
        return _PoolFunction.apply(x, qp[0], qp[1], self).view(n, c, 1, 1)


def dropout(x, p, training):
    """``F.dropout(x, p, training)`` (mobilenetv3.py:353) on a tensor that carries a grid: the kept values are scaled by
    1 / (1 - p), i.e. the result sits on the grid (scale / (1 - p), zero point) and a dropped value is the index of zero."""
    qp = getattr(x, "_frost_qparams", None)
    if qp is None or not x.is_cuda:
        return F.dropout(x, p=p, training=training)
    if not training or p == 0.0:
        return x
    from .block_engine import attach_qparams
    keep = (torch.rand(x.shape, device=x.device) >= p).float().mul_(1.0 / (1.0 - p))
    y = _BcastMulFunction.apply(x.reshape(1, x.numel(), 1, 1), keep.view(1, -1)).view(x.shape)
    return attach_qparams(y, qp[0] / (1.0 - p), qp[1])


def _run_se(mod, x):
    from .block_engine import _FakeQuantFunction, attach_qparams, qparams_of
    scale, zp = qparams_of(x, "a prepared SEModule")
    if not x.is_cuda:
        raise RuntimeError("frostnet_b200: the QAT path runs on a CUDA device (B200) only; got a CPU tensor")
    if x.dim() != 4:
        raise RuntimeError("frostnet_b200: SEModule expects an NCHW tensor")
    if not isinstance(mod.fc[1], nn.Identity):
        raise RuntimeError("frostnet_b200: call fuse_model() before attaching the fake-quants (the reference fuses fc.0 + fc.1)")
    pooled = _PoolFunction.apply(x, scale, zp, mod)
    gate = mod.fc(pooled)                                   # fc.0 (+ReLU) -> fc.2 -> Hsigmoid, each with its fake-quant
    y = _BcastMulFunction.apply(x, gate)
    fq = mod.quant_mul.activation_post_process
    y = _FakeQuantFunction.apply(y, fq, _scratch(mod, x.device))
    return attach_qparams(y, fq.scale, fq.zero_point)


def run_functional_mul(ff, x, y):
    """FloatFunctional.mul(x, gate.expand_as(x)) (mobilenetv3.py:100) called on its own: y must be a per-(image, channel) gate
    broadcast over the plane."""
    from .block_engine import _FakeQuantFunction, attach_qparams
    if not x.is_cuda:
        raise RuntimeError("frostnet_b200: the QAT path runs on a CUDA device (B200) only; got a CPU tensor")
    if x.dim() != 4 or y.dim() != 4 or y.shape[:2] != x.shape[:2] or (y.shape[2:] != (1, 1) and y.stride()[2:] != (0, 0)):
        raise RuntimeError("frostnet_b200: stand-alone FloatFunctional.mul supports x[N,C,H,W] * gate[N,C,1,1].expand_as(x) only")
    out = _BcastMulFunction.apply(x, y[:, :, 0, 0])
    fq = ff.activation_post_process
    out = _FakeQuantFunction.apply(out, fq, _scratch(ff, x.device))
    return attach_qparams(out, fq.scale, fq.zero_point)
