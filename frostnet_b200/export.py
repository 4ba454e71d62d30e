"""int8 inference export: the counterpart of ``torch.quantization.convert(model.eval(), inplace=True)`` in the
reference's evaluation script (Classification/evaluate.py:124-135).

A prepared (QAT) frostnet_b200 model keeps the reference's post-``prepare_qat`` state_dict, module for module.  The
export rebuilds, per module, the torch.ao QAT module the reference would hold at that place, loads our state into it and
lets torch's own ``from_float`` do the conversion (BN folding, weight observer pass + quantisation, output qparams from
the activation observer) - the arithmetic of the deployable model is therefore torch's, not a restatement.  The result
is a CPU module tree of ``torch.ao.nn.quantized`` modules driven by this package's float ``forward`` code; its
``state_dict()`` has the keys of the reference's converted model.  This is deployment plumbing: nothing here runs on the
training hot path.
"""
import copy

import torch
import torch.ao.nn.intrinsic.qat as nniqat
import torch.ao.nn.intrinsic.quantized as nniq
import torch.ao.nn.qat as nnqat
import torch.ao.nn.quantized as nnq
import torch.ao.quantization as taq
from torch import nn

from . import qat as Q


def _qconfig():
    return taq.get_default_qat_qconfig("qnnpack")          # Classification/evaluate.py:123


def _torch_fake_quant(fq, weight):
    """torch's FusedMovingAvgObsFakeQuantize in the state of our FrostFakeQuantize (same state_dict keys)."""
    qc = _qconfig()
    t = qc.weight() if weight else qc.activation()
    t.load_state_dict(fq.state_dict())
    return t


def _convert_conv_bn(m):
    """FrostConvBn2d -> nniq.ConvReLU2d / nnq.Conv2d through nniqat.ConvBn(ReLU)2d.from_float's target class."""
    cls = nniqat.ConvBnReLU2d if m.relu else nniqat.ConvBn2d
    t = cls(m.in_channels, m.out_channels, m.kernel_size, m.stride, m.padding, m.dilation, m.groups, bias=None,
            padding_mode="zeros", eps=m.bn.eps, momentum=m.bn.momentum, freeze_bn=False, qconfig=_qconfig())
    t.activation_post_process = _qconfig().activation()        # what prepare_qat attaches to every swapped module
    missing, unexpected = t.load_state_dict(m.state_dict(), strict=True)
    assert not missing and not unexpected
    t.eval()
    return (nniq.ConvReLU2d if m.relu else nnq.Conv2d).from_float(t)


def _convert_classifier(m):
    t = nnqat.Conv2d(m.in_channels, m.out_channels, 1, bias=m.bias is not None, qconfig=_qconfig())
    t.activation_post_process = _qconfig().activation()
    missing, unexpected = t.load_state_dict(m.state_dict(), strict=True)
    assert not missing and not unexpected
    t.eval()
    return nnq.Conv2d.from_float(t)


def _convert_functional(m):
    t = nnq.FloatFunctional()
    t.activation_post_process = _torch_fake_quant(m.activation_post_process, weight=False)
    return nnq.QFunctional.from_float(t)


def _convert_quant_stub(m):
    t = taq.QuantStub()
    t.activation_post_process = _torch_fake_quant(m.activation_post_process, weight=False)
    return nnq.Quantize.from_float(t)


def _convert_linear(m):
    """se.QATLinear -> nniq.LinearReLU / nnq.Linear through nniqat.LinearReLU / nnqat.Linear (what fuse + prepare_qat make of
    SEModule's nn.Linear layers, mobilenetv3.py:88-93, 101-102)."""
    cls = nniqat.LinearReLU if m.relu else nnqat.Linear
    t = cls(m.in_features, m.out_features, bias=False, qconfig=_qconfig())
    t.activation_post_process = _qconfig().activation()
    missing, unexpected = t.load_state_dict(m.state_dict(), strict=True)
    assert not missing and not unexpected
    t.eval()
    return (nniq.LinearReLU if m.relu else nnq.Linear).from_float(t)


def convert_int8(model):
    """Return the int8 inference model (CPU, eval) of a prepared frostnet_b200 network (FrostNet, MobileNetV3 or any tree of this
    package's modules).  `model` is left untouched."""
    if not Q.is_prepared(model):
        raise ValueError("convert_int8 expects a model after fuse_model() + prepare_qat()")
    # the copy is a plain module tree: QAT engines and device scratch hanging off the modules (``_frost_*`` entries of their
    # __dict__) are set aside while copying - an engine must not be deep-copied along with the module it belongs to
    aside = [(m, k, m.__dict__.pop(k)) for m in model.modules() for k in [k for k in m.__dict__ if k.startswith("_frost_")]]
    try:
        dst = copy.deepcopy(model)
    finally:
        for m, k, v in aside:
            m.__dict__[k] = v
    dst = dst.cpu().eval()

    from .se import QATConv1x1, QATLinear

    def swap(parent):
        for name, child in list(parent.named_children()):
            if isinstance(child, Q.FrostConvBn2d):
                setattr(parent, name, _convert_conv_bn(child))
            elif isinstance(child, (Q.FrostQATConv2d, QATConv1x1)):
                setattr(parent, name, _convert_classifier(child))
            elif isinstance(child, QATLinear):
                setattr(parent, name, _convert_linear(child))
            elif isinstance(child, nn.ReLU6):
                # nn.ReLU6 -> nnq.ReLU6 (stateless: the quantized op keeps its input's grid; the QAT observer is dropped)
                setattr(parent, name, nnq.ReLU6(child.inplace))
            elif isinstance(child, Q.FloatFunctional):
                setattr(parent, name, _convert_functional(child))
            elif isinstance(child, Q.QuantStub):
                setattr(parent, name, _convert_quant_stub(child))
            elif isinstance(child, Q.DeQuantStub):
                setattr(parent, name, nnq.DeQuantize())
            else:
                swap(child)
    swap(dst)
    return dst
