"""FrostNet classifier - the reference's nn.Module surface (frostnet.py) on the B200-native engine.

Same constructors, attribute names, factory names and ``state_dict`` layout as
``/root/reference/frostnet.py`` (ConvBNReLU :14-28, ConvBN :46-60, _make_divisible :62-79,
CascadePreExBottleneck :81-145, FrostNet :150-351, factories :354-451), so a script written against
the reference only changes its import.  What differs is what runs underneath:

* float (FP warm-up) mode: stock torch ops - outside the QAT hot path (StatAssist phase 1);
* after ``fuse_model()`` + ``frostnet_b200.prepare_qat(model)``: ``forward`` is ONE autograd node
  executed by :class:`frostnet_b200.engine.QATEngine` - hand-written sm_100a kernels on uint8
  indices / int32 accumulators.  There is no CPU or eager fallback in that mode.
"""
import torch
import torch.nn as nn

from . import qat as _qat

__all__ = ["ConvBNReLU", "ConvBN", "CascadePreExBottleneck", "FrostNet", "_make_divisible", "SETTINGS"]


class _ConvBlock(nn.Module):
    """Shared body of ConvBNReLU / ConvBN: ``self.conv`` is the reference's nn.Sequential (``_seq_name`` lets the MobileNetV3
    wrappers, whose Sequential is called ``cbr`` / ``cb``, reuse it)."""
    _relu = False
    _seq_name = "conv"

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1):
        super().__init__()
        layers = [nn.Conv2d(in_channels, out_channels, kernel_size, stride, padding, dilation, groups, bias=False),
                  nn.BatchNorm2d(out_channels)]
        if self._relu:
            layers.append(nn.ReLU(False))
        setattr(self, self._seq_name, nn.Sequential(*layers))

    def _seq(self):
        return getattr(self, self._seq_name)

    def forward(self, x):
        seq = self._seq()
        if isinstance(getattr(seq[0], "weight_fake_quant", None), _qat.FrostFakeQuantize):
            # prepared and called on its own (the whole-network engine never comes here): per-module executor
            from .block_engine import run_block
            return run_block(self, x)
        return seq(x)

    def fuse_model(self):
        """Reference: torch.quantization.fuse_modules(self.conv, ['0','1'(,'2')], inplace=True)
        (frostnet.py:27-28, 59-60).  Produces the same module tree / state_dict keys
        (``conv.0.weight``, ``conv.0.bn.*``; Identity at 1(,2))."""
        seq = self._seq()
        if isinstance(seq[0], _qat.FrostConvBn2d):
            return
        fused = _qat.FrostConvBn2d(seq[0], seq[1], relu=self._relu)
        seq[0] = fused
        for i in range(1, len(seq)):
            seq[i] = nn.Identity()


class ConvBNReLU(_ConvBlock):
    _relu = True


class ConvBN(_ConvBlock):
    _relu = False


def _make_divisible(v, divisor=8, min_value=None):
    """frostnet.py:62-79."""
    if min_value is None:
        min_value = divisor
    new_v = max(min_value, int(v + divisor / 2) // divisor * divisor)
    if new_v < 0.9 * v:
        new_v += divisor
    return new_v


class CascadePreExBottleneck(nn.Module):
    """The Frost bottleneck (frostnet.py:81-145): optional squeeze 1x1 -> cat -> expand 1x1 ->
    depthwise kxk -> reduce 1x1 (-> skip add)."""

    def __init__(self, in_channels, out_channels, quantized=False, kernel_size=3, stride=1, dilation=1,
                 expand_ratio=6, reduce_factor=4, block_type='CAS'):
        super().__init__()
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.kernel_size = kernel_size
        self.stride = stride
        self.expand_ratio = expand_ratio
        self.quantized = quantized
        if in_channels // reduce_factor < 8:
            block_type = 'MB'
        self.block_type = block_type
        r_channels = _make_divisible(in_channels // reduce_factor)
        self.reduction = not (stride == 1 and in_channels == out_channels)
        if self.expand_ratio == 1:
            self.squeeze_conv = None
            self.conv1 = None
            n_channels = in_channels
        else:
            if block_type == 'CAS':
                self.squeeze_conv = ConvBNReLU(in_channels, r_channels, 1)
                n_channels = r_channels + in_channels
            else:
                n_channels = in_channels
            self.conv1 = ConvBNReLU(n_channels, n_channels * expand_ratio, 1)
        # NB: like the reference, `dilation` is accepted but conv2 is built with dilation 1 (:116-118)
        self.conv2 = ConvBNReLU(n_channels * expand_ratio, n_channels * expand_ratio, kernel_size, stride,
                                (kernel_size - 1) // 2, 1, groups=n_channels * expand_ratio)
        self.reduce_conv = ConvBN(n_channels * expand_ratio, out_channels, 1)
        if self.quantized:
            self.skip_add = _qat.FloatFunctional()
            self.quant_cat = _qat.FloatFunctional()

    def has_squeeze(self):
        return self.expand_ratio != 1 and self.block_type == 'CAS'

    def forward(self, x):
        if _qat.is_prepared(self):
            # called on its own (the whole-network engine never comes here): per-module executor, fp32 NCHW at the boundary
            from .block_engine import run_block
            return run_block(self, x)
        # float model and the int8 model made by frostnet_b200.convert_int8: cat / add go through the FloatFunctional /
        # QFunctional modules exactly where the reference's do (frostnet.py:124-145)
        if not self.expand_ratio == 1:
            if self.block_type == 'CAS':
                squeezed = self.squeeze_conv(x)
                out = self.quant_cat.cat([squeezed, x], 1) if self.quantized else torch.cat([squeezed, x], 1)
            else:
                out = x
            out = self.conv1(out)
        else:
            out = x
        out = self.conv2(out)
        out = self.reduce_conv(out)
        if not self.reduction:
            out = self.skip_add.add(x, out) if self.quantized else torch.add(x, out)
        return out


# kernel_size, c, e, r, s   (frostnet.py:157-269)
SETTINGS = {
    "large": [
        [[3, 16, 1, 1, 1], [3, 24, 6, 4, 2], [3, 24, 3, 4, 1]],
        [[5, 40, 6, 4, 2], [3, 40, 3, 4, 1]],
        [[5, 80, 6, 4, 2], [5, 80, 3, 4, 1], [5, 80, 3, 4, 1], [5, 96, 6, 4, 1],
         [5, 96, 3, 4, 1], [3, 96, 3, 4, 1], [3, 96, 3, 4, 1]],
        [[5, 192, 6, 2, 2], [5, 192, 6, 4, 1], [5, 192, 6, 4, 1], [5, 192, 3, 4, 1], [5, 192, 3, 4, 1]],
        [[5, 320, 6, 2, 1]],
    ],
    "base": [
        [[3, 16, 1, 1, 1], [5, 24, 6, 4, 2], [3, 24, 3, 4, 1]],
        [[5, 40, 3, 4, 2], [5, 40, 3, 4, 1]],
        [[5, 80, 3, 4, 2], [3, 80, 3, 4, 1], [5, 96, 3, 2, 1], [3, 96, 3, 4, 1],
         [5, 96, 3, 4, 1], [5, 96, 3, 4, 1]],
        [[5, 192, 6, 2, 2], [5, 192, 3, 2, 1], [5, 192, 3, 2, 1], [5, 192, 3, 2, 1]],
        [[5, 320, 6, 2, 1]],
    ],
    "small": [
        [[3, 16, 1, 1, 1], [5, 24, 3, 4, 2], [3, 24, 3, 4, 1]],
        [[5, 40, 3, 4, 2]],
        [[5, 80, 3, 4, 2], [5, 80, 3, 4, 1], [3, 80, 3, 4, 1], [5, 96, 3, 2, 1],
         [5, 96, 3, 4, 1], [5, 96, 3, 4, 1]],
        [[5, 192, 6, 4, 2], [5, 192, 6, 4, 1], [5, 192, 6, 4, 1]],
        [[5, 320, 6, 2, 1]],
    ],
}


class _FrostTrunk(nn.Module):
    """Stem + 5 stages shared by the classifier and the feature backbone."""

    def _build_trunk(self, mode, width_mult, bottleneck, quantized, dilated=False):
        if mode not in SETTINGS:
            raise ValueError('Unknown mode.')
        self.quantized = quantized
        self.in_channels = _make_divisible(int(32 * min(1.0, width_mult)))
        self.conv1 = ConvBNReLU(3, self.in_channels, 3, 2, 1)
        s = SETTINGS[mode]
        self.layer1 = self._make_layer(bottleneck, s[0], width_mult, 1)
        self.layer2 = self._make_layer(bottleneck, s[1], width_mult, 1)
        self.layer3 = self._make_layer(bottleneck, s[2], width_mult, 1)
        dilation = 2 if dilated else 1
        self.layer4 = self._make_layer(bottleneck, s[3], width_mult, dilation)
        self.layer5 = self._make_layer(bottleneck, s[4], width_mult, dilation)

    def _make_layer(self, block, block_setting, width_mult, dilation=1):
        layers = []
        for k, c, e, r, s in block_setting:
            out_channels = _make_divisible(int(c * width_mult))
            # reference quirk (frostnet.py:312-314): `stride` is computed and ignored; `s` is passed
            layers.append(block(self.in_channels, out_channels, quantized=self.quantized, kernel_size=k,
                                stride=s, dilation=dilation, expand_ratio=e, reduce_factor=r))
            self.in_channels = out_channels
        return nn.Sequential(*layers)

    def _init_weights(self):
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode='fan_out')
                if m.bias is not None:
                    nn.init.zeros_(m.bias)
            elif isinstance(m, nn.BatchNorm2d):
                nn.init.ones_(m.weight)
                nn.init.zeros_(m.bias)
            elif isinstance(m, nn.Linear):
                nn.init.normal_(m.weight, 0, 0.01)
                if m.bias is not None:
                    nn.init.zeros_(m.bias)

    def fuse_model(self):
        for m in self.modules():
            if type(m) in (ConvBNReLU, ConvBN):
                m.fuse_model()

    def stages(self):
        return [self.layer1, self.layer2, self.layer3, self.layer4, self.layer5]

    def _apply(self, fn, *args, **kwargs):
        out = super()._apply(fn, *args, **kwargs)
        eng = self.__dict__.get("_frost_engine")
        if eng is not None:
            eng.invalidate()       # tensors may have moved (.cuda()/.to()): rebuild pointer tables lazily
        return out


class FrostNet(_FrostTrunk):
    """frostnet.py:150-351."""

    def __init__(self, nclass=1000, mode='large', width_mult=1.0, quantized=False,
                 bottleneck=CascadePreExBottleneck, drop_rate=0.2, dilated=False, **kwargs):
        super().__init__()
        self._build_trunk(mode, width_mult, bottleneck, quantized, dilated)
        last_in_channels = self.in_channels
        self.last_layer = ConvBNReLU(last_in_channels, 1280, 1)
        self.classifier = nn.Sequential(nn.AdaptiveAvgPool2d(1), nn.Dropout(drop_rate), nn.Conv2d(1280, nclass, 1))
        self.mode = mode
        self.drop_rate = drop_rate
        self._init_weights()
        if self.quantized:
            self.quant = _qat.QuantStub()
            self.dequant = _qat.DeQuantStub()

    def forward(self, x):
        eng = self.__dict__.get("_frost_engine")
        if eng is not None:
            return eng.run(x)
        if self.quantized:
            x = self.quant(x)          # identity in the float model, nnq.Quantize after convert_int8 (frostnet.py:319-320)
        x = self.conv1(x)
        for st in self.stages():
            x = st(x)
        x = self.last_layer(x)
        x = self.classifier(x)
        if self.quantized:
            x = self.dequant(x)
        return x.view(x.size(0), x.size(1))


def _factory(mode, wm, quant):
    def f(**kwargs):
        return FrostNet(nclass=1000, mode=mode, width_mult=wm, quantized=quant,
                        bottleneck=CascadePreExBottleneck, **kwargs)
    return f


# the 30 factories of frostnet.py:354-451
for _q in (True, False):
    for _mode in ("large", "base", "small"):
        for _wm, _tag in ((1.25, "1_25"), (1.0, "1_0"), (0.75, "0_75"), (0.5, "0_5"), (0.35, "0_35")):
            _name = "frostnet_%s%s_%s" % ("quant_" if _q else "", _mode, _tag)
            globals()[_name] = _factory(_mode, _wm, _q)
            globals()[_name].__name__ = _name
            __all__.append(_name)
