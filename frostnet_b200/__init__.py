"""frostnet_b200 - B200-native (sm_100a) implementation of FrostNet's QAT forward/backward hot path
and the StatAssist/GradBoost optimizer step, behind the reference's nn.Module / Optimizer surface.

    from frostnet_b200 import frostnet_quant_large_1_0, prepare_qat, get_optimizer
"""
from .frostnet import *          # noqa: F401,F403  (FrostNet, blocks, 30 factories)
from .frostnet import FrostNet, CascadePreExBottleneck, ConvBNReLU, ConvBN
from .qat import (prepare_qat, patch_torch_quantization, attach_fake_quant, fuse_modules, FrostFakeQuantize, QuantStub,
                  DeQuantStub)
from .optimizer import QSGD, QRMSprop, QAdam, QAdamW, get_optimizer
from . import parallel
from .prefetch import DevicePrefetcher
from . import frostnet_features

__version__ = "0.1.0"
from .export import convert_int8
from .hswish import Hsigmoid, Hswish
from .se import SEModule, QATLinear
from . import mobilenetv3
from . import mobilenetv2
from .multibox import MultiBoxLoss, match_batch
