"""QAT module tree: the shapes torch's eager-mode QAT would produce for the reference, owned by us.

After ``model.fuse_model()`` + ``prepare_qat(model)`` the module tree has the same names, parameters
and buffers (hence the same ``state_dict`` keys, dtypes and shapes) as the reference after
``fuse_model(); model.qconfig = get_default_qat_qconfig('qnnpack'); prepare_qat(model)``
(Classification/train.py:166-173):

  <p>.conv.0           FrostConvBn2d      <-> nniqat.ConvBnReLU2d / ConvBn2d (conv_fused.py:613-710)
  <p>.conv.0.bn        nn.BatchNorm2d     (same module object as before fusion)
  *.weight_fake_quant / *.activation_post_process
                       FrostFakeQuantize  <-> FusedMovingAvgObsFakeQuantize (fake_quantize.py:340-438)
  classifier.2         FrostQATConv2d     <-> nnqat.Conv2d (torch/ao/nn/qat/modules/conv.py)
  skip_add / quant_cat FloatFunctional    <-> torch.ao.nn.quantized.FloatFunctional
  quant / dequant      QuantStub / DeQuantStub

Parameter OBJECTS are preserved across fuse+prepare (the optimizer built on the float model keeps
its state, S1 in SURVEY.md 8a).  These modules only hold state; the arithmetic of a prepared model
runs in frostnet_b200.engine.QATEngine.
"""
import torch
import torch.nn as nn

ACT_QMIN, ACT_QMAX = 0, 255
W_QMIN, W_QMAX = -128, 127
AVERAGING_CONSTANT = 0.01


class FrostObserverState(nn.Module):
    """Buffers of MovingAverageMinMaxObserver (torch/ao/quantization/observer.py:560-683)."""

    def __init__(self):
        super().__init__()
        self.register_buffer("eps", torch.tensor([torch.finfo(torch.float32).eps]))
        self.register_buffer("min_val", torch.tensor(float("inf")))
        self.register_buffer("max_val", torch.tensor(float("-inf")))
        self.averaging_constant = AVERAGING_CONSTANT


class FrostFakeQuantize(torch.ao.quantization.FakeQuantizeBase):
    """State of one FusedMovingAvgObsFakeQuantize (per-tensor).  K1 in SURVEY.md 8c.

    Derives from torch's FakeQuantizeBase so that ``model.apply(torch.ao.quantization.disable_observer)``
    (Classification/train.py:27-33 defines exactly that helper) reaches it."""

    flag_epoch = 0      # bumped whenever any instance toggles a flag; the engine re-encodes its tables

    def __init__(self, quant_min, quant_max, symmetric):
        super().__init__()
        self.quant_min, self.quant_max, self.is_symmetric_quant = quant_min, quant_max, symmetric
        self.activation_post_process = FrostObserverState()
        # FusedMovingAvgObsFakeQuantize keeps these two as int64 (state_dict compatibility)
        del self._buffers["fake_quant_enabled"], self._buffers["observer_enabled"]
        self.register_buffer("fake_quant_enabled", torch.tensor([1], dtype=torch.long))
        self.register_buffer("observer_enabled", torch.tensor([1], dtype=torch.long))
        self.register_buffer("scale", torch.tensor([1.0], dtype=torch.float))
        self.register_buffer("zero_point", torch.tensor([0], dtype=torch.int))
        # host mirrors of the two enable flags (the device copies are only state_dict payload)
        self._observe = True
        self._fake_quant = True

    @staticmethod
    def act():
        return FrostFakeQuantize(ACT_QMIN, ACT_QMAX, False)

    @staticmethod
    def weight():
        return FrostFakeQuantize(W_QMIN, W_QMAX, True)

    # torch.ao.quantization.{enable,disable}_{observer,fake_quant} protocol
    def enable_observer(self, enabled=True):
        self._observe = bool(enabled)
        self.observer_enabled[0] = 1 if enabled else 0
        FrostFakeQuantize.flag_epoch += 1

    def calculate_qparams(self, **kwargs):
        return self.scale, self.zero_point

    def disable_observer(self):
        self.enable_observer(False)

    def enable_fake_quant(self, enabled=True):
        if not enabled:
            raise RuntimeError("frostnet_b200: the QAT engine computes on quantize indices; "
                               "fake_quant cannot be disabled")
        self._fake_quant = True
        self.fake_quant_enabled[0] = 1

    def disable_fake_quant(self):
        self.enable_fake_quant(False)

    def _load_from_state_dict(self, state_dict, prefix, *args, **kwargs):
        super()._load_from_state_dict(state_dict, prefix, *args, **kwargs)
        k = prefix + "observer_enabled"
        if k in state_dict:
            self._observe = bool(int(state_dict[k].reshape(-1)[0]))
            FrostFakeQuantize.flag_epoch += 1

    def forward(self, x):
        raise RuntimeError("frostnet_b200: FrostFakeQuantize is state only; it runs inside the QAT engine")

    def extra_repr(self):
        return "quant_min=%d, quant_max=%d, symmetric=%s" % (self.quant_min, self.quant_max, self.is_symmetric_quant)


class FloatFunctional(nn.Module):
    """State holder for FloatFunctional.add / .cat (functional_modules.py:50-52,80-82)."""

    def __init__(self):
        super().__init__()
        self.activation_post_process = nn.Identity()

    def forward(self, x):
        raise RuntimeError("FloatFunctional is not intended to use the 'forward'. Please use the underlying operation")

    def add(self, x, y):
        if isinstance(self.activation_post_process, FrostFakeQuantize):
            # prepared and called on its own (the whole-network engine never comes here): per-module executor
            from .block_engine import run_functional_add
            return run_functional_add(self, x, y)
        return torch.add(x, y)

    def cat(self, xs, dim=0):
        if isinstance(self.activation_post_process, FrostFakeQuantize):
            from .block_engine import run_functional_cat
            return run_functional_cat(self, xs, dim)
        return torch.cat(xs, dim)

    def mul(self, x, y):
        if isinstance(self.activation_post_process, FrostFakeQuantize):
            from .se import run_functional_mul
            return run_functional_mul(self, x, y)
        return torch.mul(x, y)

    def add_scalar(self, x, y):
        """functional_modules.py:58-63: not observed (the quantized op needs no output qparams)"""
        return torch.add(x, y)

    def mul_scalar(self, x, y):
        """functional_modules.py:72-77: not observed"""
        return torch.mul(x, y)


class QuantStub(nn.Module):
    def forward(self, x):
        if isinstance(getattr(self, "activation_post_process", None), FrostFakeQuantize):
            # prepared and called on its own (the whole-network engine never comes here): observer + fake-quant, the result
            # carries its (scale, zero_point) for the frostnet_b200 module that consumes it
            from .block_engine import run_quant_stub
            return run_quant_stub(self, x)
        return x


class DeQuantStub(nn.Module):
    def forward(self, x):
        return x


class FrostConvBn2d(nn.Module):
    """Fused Conv+BN(+ReLU).  Float behaviour before prepare (== nni.ConvBn(ReLU)2d in train mode);
    after prepare it carries weight_fake_quant / activation_post_process like nniqat.ConvBn(ReLU)2d."""

    def __init__(self, conv, bn, relu):
        super().__init__()
        self.in_channels, self.out_channels = conv.in_channels, conv.out_channels
        self.kernel_size, self.stride, self.padding = conv.kernel_size, conv.stride, conv.padding
        self.dilation, self.groups = conv.dilation, conv.groups
        self.weight = conv.weight          # same Parameter object
        self.bias = None
        self.bn = bn                       # same module object
        self.relu = relu

    @property
    def is_depthwise(self):
        return self.groups == self.in_channels and self.groups > 1

    def forward(self, x):
        if isinstance(getattr(self, "weight_fake_quant", None), FrostFakeQuantize):
            # prepared and called directly (a fused conv sitting in somebody's nn.Sequential, e.g. torchvision-style
            # ConvBNReLU / InvertedResidual of Object_Detection/ssd_qmv2.py:40-110): per-module executor
            from .block_engine import run_block
            return run_block(self, x)
        y = nn.functional.conv2d(x, self.weight, None, self.stride, self.padding, self.dilation, self.groups)
        y = self.bn(y)
        return nn.functional.relu(y) if self.relu else y

    def extra_repr(self):
        return "%d, %d, kernel_size=%s, stride=%s, groups=%d, relu=%s" % (
            self.in_channels, self.out_channels, self.kernel_size, self.stride, self.groups, self.relu)


class FrostQATConv2d(nn.Module):
    """classifier.2 after prepare: nnqat.Conv2d(1280, nclass, 1) with bias."""

    def __init__(self, conv):
        super().__init__()
        self.in_channels, self.out_channels = conv.in_channels, conv.out_channels
        self.weight, self.bias = conv.weight, conv.bias
        self.weight_fake_quant = FrostFakeQuantize.weight()
        self.activation_post_process = FrostFakeQuantize.act()

    def forward(self, x):
        raise RuntimeError("frostnet_b200: a prepared FrostQATConv2d only runs inside the QAT engine")


def is_prepared(module):
    for m in module.modules():
        if isinstance(m, FrostFakeQuantize):
            return True
    return False


def prepare_qat(model, inplace=True):
    """Drop-in for ``torch.quantization.prepare_qat(model, inplace=True)`` with the qnnpack QAT
    qconfig (Classification/train.py:168-173).  ``model.qconfig`` is accepted and checked if set."""
    from . import engine as _engine
    from . import frostnet as _fn
    if not inplace:
        raise ValueError("frostnet_b200.prepare_qat only supports inplace=True (Parameter identity is the contract)")
    if not isinstance(model, _fn._FrostTrunk):
        raise TypeError("frostnet_b200.prepare_qat expects a frostnet_b200 FrostNet")
    if not getattr(model, "quantized", False):
        raise ValueError("prepare_qat needs a model built with quantized=True (frostnet_quant_* factories)")
    if is_prepared(model):
        return model
    model.fuse_model()      # no-op if the caller already fused (train.py:171)
    dev = next(model.parameters()).device
    for m in list(model.modules()):
        if isinstance(m, FrostConvBn2d):
            m.weight_fake_quant = FrostFakeQuantize.weight().to(dev)
            m.activation_post_process = FrostFakeQuantize.act().to(dev)
        elif isinstance(m, FloatFunctional):
            m.activation_post_process = FrostFakeQuantize.act().to(dev)
        elif isinstance(m, QuantStub):
            m.activation_post_process = FrostFakeQuantize.act().to(dev)
    cls = getattr(model, "classifier", None)
    if cls is not None and isinstance(cls[2], nn.Conv2d):
        cls[2] = FrostQATConv2d(cls[2]).to(dev)
    model.__dict__["_frost_engine"] = _engine.QATEngine(model)
    return model


def fuse_modules(container, names, inplace=True):
    """``torch.quantization.fuse_modules(container, ['0', '1'(, '2')], inplace=True)`` for Conv2d + BatchNorm2d (+ ReLU) children of
    any module (frostnet.py:27-28, ssd_qmv2.py:178-185, mobilenetv3.py:21-25): the first name becomes a FrostConvBn2d holding the
    SAME Parameter / BatchNorm objects (state_dict keys ``<first>.weight``, ``<first>.bn.*`` like nni.ConvBn(ReLU)2d after
    prepare_qat), the others nn.Identity."""
    if not inplace:
        raise ValueError("frostnet_b200.fuse_modules fuses in place")
    mods = [getattr(container, n) if not isinstance(container, nn.Sequential) else container[int(n)] for n in names]
    if isinstance(mods[0], FrostConvBn2d):
        return container
    if not (isinstance(mods[0], nn.Conv2d) and len(mods) in (2, 3) and isinstance(mods[1], nn.BatchNorm2d)
            and (len(mods) == 2 or isinstance(mods[2], nn.ReLU))):
        raise ValueError("frostnet_b200.fuse_modules: expected Conv2d, BatchNorm2d(, ReLU), got %s" % [type(m).__name__ for m in mods])
    if mods[0].bias is not None:
        raise ValueError("frostnet_b200.fuse_modules: the convolution must not have a bias")
    repl = [FrostConvBn2d(mods[0], mods[1], relu=len(mods) == 3)] + [nn.Identity() for _ in mods[1:]]
    for n, r in zip(names, repl):
        if isinstance(container, nn.Sequential):
            container[int(n)] = r
        else:
            setattr(container, n, r)
    return container


def attach_fake_quant(root):
    """prepare_qat for an arbitrary tree of frostnet_b200 modules (bottlenecks wired into another network, Hswish, a
    QuantStub in front): attaches the qnnpack-QAT fake-quants where torch.quantization.prepare_qat would - on every
    fused conv (weight + output), FloatFunctional, QuantStub, the ReLU6 inside Hswish / Hsigmoid and the Linear layers of SEModule.  Convs must be fused first
    (``fuse_model()``).  The modules then run through the per-module executor (block_engine.py)."""
    from .hswish import Hsigmoid, Hswish
    from .se import QATConv1x1, QATLinear
    dev = next((p.device for p in root.parameters()), torch.device("cpu"))
    for m in list(root.modules()):
        if isinstance(m, FrostConvBn2d) and not isinstance(getattr(m, "weight_fake_quant", None), FrostFakeQuantize):
            m.weight_fake_quant = FrostFakeQuantize.weight().to(dev)
            m.activation_post_process = FrostFakeQuantize.act().to(dev)
        elif isinstance(m, (FloatFunctional, QuantStub)) and not isinstance(getattr(m, "activation_post_process", None), FrostFakeQuantize):
            m.activation_post_process = FrostFakeQuantize.act().to(dev)
        elif isinstance(m, (Hswish, Hsigmoid)) and not m._prepared():
            m.relu6.activation_post_process = FrostFakeQuantize.act().to(dev)
        elif isinstance(m, (QATLinear, QATConv1x1)) and not m._prepared():
            m.weight_fake_quant = FrostFakeQuantize.weight().to(dev)
            m.activation_post_process = FrostFakeQuantize.act().to(dev)
    return root


def patch_torch_quantization():
    """Route torch.quantization.prepare_qat / fuse_modules to this package for frostnet_b200 models so
    that an unmodified caller script (Classification/train.py:166-173) works unchanged."""
    import torch.ao.quantization as taq
    from . import frostnet as _fn
    orig = taq.prepare_qat
    if getattr(orig, "_frost_patched", False):
        return

    def _prepare_qat(model, mapping=None, inplace=False):
        if isinstance(model, _fn._FrostTrunk):
            return prepare_qat(model, inplace=True)
        return orig(model, mapping, inplace)
    _prepare_qat._frost_patched = True
    taq.prepare_qat = _prepare_qat
    torch.quantization.prepare_qat = _prepare_qat
