"""FrostNet feature backbone - the reference's ``frostnet_features.py`` surface (an mmdet BACKBONE there).

``FrostNet(mode, width_mult, bottleneck, quantized, pretrained)`` builds the same stem + 5 stages as the
classifier (no QuantStub, no last_layer / classifier head) and ``forward`` returns ``[x1, x2, x3, x5]`` - the
stage outputs at strides 4 / 8 / 16 / 32 (frostnet_features.py:342-352).  After ``fuse_model()`` +
``frostnet_b200.prepare_qat`` the forward/backward run in the QAT engine: the stem convolves the raw fp32
image with the fake-quantised stem weights (there is no input fake-quant in this variant), everything after
it is the uint8-index path; the four maps come back dequantised, NCHW fp32, with gradients.
"""
import os
from collections import OrderedDict

import torch
import torch.nn as nn

from .frostnet import CascadePreExBottleneck, ConvBN, ConvBNReLU, _FrostTrunk

__all__ = ["FrostNet", "load_state_dict", "load_checkpoint"]


def load_state_dict(checkpoint_path, use_ema=False):
    """frostnet_features.py:10-30: strips a leading ``module.`` and prefers ``state_dict_ema``."""
    if checkpoint_path and os.path.isfile(checkpoint_path):
        checkpoint = torch.load(checkpoint_path, map_location='cpu')
        state_dict_key = 'state_dict'
        if isinstance(checkpoint, dict):
            if use_ema and 'state_dict_ema' in checkpoint:
                state_dict_key = 'state_dict_ema'
        if state_dict_key and state_dict_key in checkpoint:
            new_state_dict = OrderedDict()
            for k, v in checkpoint[state_dict_key].items():
                name = k[7:] if k.startswith('module') else k
                new_state_dict[name] = v
            state_dict = new_state_dict
        else:
            state_dict = checkpoint
        print("Loaded {} from checkpoint '{}'".format(state_dict_key, checkpoint_path))
        return state_dict
    print("No checkpoint found at '{}'".format(checkpoint_path))
    raise FileNotFoundError()


def load_checkpoint(model, checkpoint_path, use_ema=False, strict=True):
    model.load_state_dict(load_state_dict(checkpoint_path, use_ema), strict=strict)


class FrostNet(_FrostTrunk):
    """frostnet_features.py:170-359."""

    def __init__(self, mode='large', width_mult=1.0, bottleneck=CascadePreExBottleneck, quantized=False,
                 pretrained='', **kwargs):
        super().__init__()
        self._build_trunk(mode, width_mult, bottleneck, quantized, dilated=False)
        self.mode = mode

    def init_weights(self, pretrained):
        if pretrained != '':
            load_checkpoint(self, pretrained, use_ema=True, strict=False)
        else:
            print('No pretrained backbone provided')
            self._init_weights()

    def forward(self, x):
        eng = self.__dict__.get("_frost_engine")
        if eng is not None:
            return eng.run(x)
        x = self.conv1(x)
        x1 = self.layer1(x)
        x2 = self.layer2(x1)
        x3 = self.layer3(x2)
        x4 = self.layer4(x3)
        x5 = self.layer5(x4)
        return [x1, x2, x3, x5]

    def _freeze_stages(self):
        '''Freeze BatchNorm layers.'''
        print('Freeze BatchNorm layers.')
        for layer in self.modules():
            if isinstance(layer, nn.BatchNorm2d):
                layer.eval()
