"""Quantization-aware hard-swish: the reference's ``_Hswish`` (Classification/models/imagenet/mobilenetv3.py:43-56;
SURVEY.md 8f, row f4) with the same attribute names and state_dict keys
(``relu6.activation_post_process.*``, ``quant_mul1.*``, ``quant_mul2.*``, ``quant_add.*``).

Float behaviour before ``attach_fake_quant`` / prepare; afterwards ONE index pass + ONE table pass on the device
(csrc/hswish.cu) through the per-module calling convention of ``block_engine``: the input carries its grid as
``_frost_qparams``, the result carries ``(scale_mul / 6, zero_point_mul)``.
"""
import torch
from torch import nn

from . import _lib as L
from . import qat as Q


class Hswish(nn.Module):
    def __init__(self, inplace=True):
        super().__init__()
        self.relu6 = nn.ReLU6(inplace)
        self.quant_mul1 = Q.FloatFunctional()
        self.quant_mul2 = Q.FloatFunctional()
        self.quant_add = Q.FloatFunctional()

    def _prepared(self):
        return isinstance(getattr(self.relu6, "activation_post_process", None), Q.FrostFakeQuantize)

    def forward(self, x):
        if self._prepared():
            return _run(self, x)
        # through the functional modules, like the reference: the int8 export (export.py) swaps them for QFunctional
        out = self.quant_add.add_scalar(x, 3.0)
        out = self.relu6(out)
        out = self.quant_mul1.mul(x, out)
        return self.quant_mul2.mul_scalar(out, 1 / 6)


class Hsigmoid(nn.Module):
    """The reference's ``_Hsigmoid`` (mobilenetv3.py:59-69): add_scalar(3) -> ReLU6 (+ its fake-quant) -> mul_scalar(1/6)."""

    def __init__(self, inplace=True):
        super().__init__()
        self.relu6 = nn.ReLU6(inplace)
        self.quant_add = Q.FloatFunctional()
        self.quant_mul = Q.FloatFunctional()

    def _prepared(self):
        return isinstance(getattr(self.relu6, "activation_post_process", None), Q.FrostFakeQuantize)

    def forward(self, x):
        if self._prepared():
            return _run(self, x)
        return self.quant_mul.mul_scalar(self.relu6(self.quant_add.add_scalar(x, 3.0)), 1 / 6)


class _HswishFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, in_scale, in_zp, mod):
        x = x.detach()
        if x.dtype != torch.float32 or not x.is_contiguous():
            x = x.float().contiguous()
        dev = x.device
        sigmoid = isinstance(mod, Hsigmoid)
        fa = mod.relu6.activation_post_process
        fb = fa if sigmoid else mod.quant_mul1.activation_post_process
        n = x.numel()
        ws = torch.empty(L.load().frost_hswish_workspace_floats(), dtype=torch.float32, device=dev)
        q_in = torch.empty(x.shape, dtype=torch.uint8, device=dev)
        y = torch.empty_like(x)
        out_scale = mod.__dict__.get("_frost_out_scale")
        if out_scale is None or out_scale.device != dev:
            out_scale = torch.ones(1, dtype=torch.float32, device=dev)
            mod.__dict__["_frost_out_scale"] = out_scale
        fqa = L.FQ(fa.activation_post_process.min_val.data_ptr(), fa.activation_post_process.max_val.data_ptr(),
                   fa.scale.data_ptr(), fa.zero_point.data_ptr())
        with torch.cuda.device(dev):
            if sigmoid:
                L.call("frost_hsigmoid_forward", x.data_ptr(), n, in_scale.data_ptr(), in_zp.data_ptr(), fqa, 1 if fa._observe else 0,
                       Q.AVERAGING_CONSTANT, q_in.data_ptr(), y.data_ptr(), None, ws.data_ptr(), out_scale.data_ptr(), L.stream(dev))
            else:
                L.call("frost_hswish_forward", x.data_ptr(), n, in_scale.data_ptr(), in_zp.data_ptr(), fqa, 1 if fa._observe else 0,
                       L.FQ(fb.activation_post_process.min_val.data_ptr(), fb.activation_post_process.max_val.data_ptr(),
                            fb.scale.data_ptr(), fb.zero_point.data_ptr()), 1 if fb._observe else 0,
                       Q.AVERAGING_CONSTANT, q_in.data_ptr(), y.data_ptr(), None, ws.data_ptr(), out_scale.data_ptr(), L.stream(dev))
        ctx.q_in, ctx.ws = q_in, ws
        return y

    @staticmethod
    def backward(ctx, dy):
        dy = dy.contiguous().float()
        dx = torch.empty_like(dy)
        with torch.cuda.device(dy.device):
            L.call("frost_hswish_backward", dy.data_ptr(), ctx.q_in.data_ptr(), dy.numel(), ctx.ws.data_ptr(), dx.data_ptr(),
                   L.stream(dy.device))
        return dx, None, None, None


def _run(mod, x):
    from .block_engine import attach_qparams, qparams_of
    in_scale, in_zp = qparams_of(x, "a prepared Hswish")
    if not x.is_cuda:
        raise RuntimeError("frostnet_b200: the QAT path runs on a CUDA device (B200) only; got a CPU tensor")
    y = _HswishFunction.apply(x, in_scale, in_zp, mod)
    zp_mod = mod.relu6 if isinstance(mod, Hsigmoid) else mod.quant_mul1
    return attach_qparams(y, mod.__dict__["_frost_out_scale"], zp_mod.activation_post_process.zero_point)
