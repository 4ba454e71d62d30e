"""Golden vectors for the MobileNetV2 inverted-residual block of the SSDLite backbone: the REAL reference
classes (Object_Detection/ssd_qmv2.py:40-110: ConvBNReLU as an nn.Sequential, InvertedResidual with a bare Conv2d +
BatchNorm2d tail), fused by the reference's own rule (MobileNetV2.fuse_model, :178-185), prepared with the qnnpack QAT qconfig,
three training steps per configuration.  ssd_qmv2.py imports the detection data pipeline and torchvision names that no longer
exist; those imports (unrelated to the two classes) are stubbed.  Dilation 1 and 2 (the last two stages of the backbone).  Runs only in the build container; tests/golden/mbv2_block.pt is committed.

    python tests/golden/make_golden_mbv2_block.py
"""
import importlib.util
import os
import sys
import types
import warnings

import torch
import torch.nn as nn

warnings.filterwarnings("ignore")
HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/Object_Detection/ssd_qmv2.py"
#        inp, oup, stride, expand, H, dilation
CASES = [(32, 16, 1, 1, 12, 1),    # t = 1: no expand conv
         (16, 24, 2, 6, 12, 1),    # stride 2
         (24, 24, 1, 6, 8, 1),     # residual
         (24, 24, 1, 6, 10, 2),    # residual, dilated depthwise (the backbone's last two stages)
         (24, 32, 1, 6, 10, 2)]    # dilated, no residual


def load_reference():
    tv = types.ModuleType("torchvision.models.mobilenet")
    tv.InvertedResidual = tv.ConvBNReLU = tv.MobileNetV2 = object
    sys.modules["torchvision.models.mobilenet"] = tv
    for name in ("layers", "data"):
        m = types.ModuleType(name)
        m.__all__ = []
        sys.modules[name] = m
    sys.modules["data"].voc, sys.modules["data"].coco = {}, {}
    torch.quantization.fuse_modules = torch.ao.quantization.fuse_modules_qat     # the reference fuses in train mode (torch 1.6)
    spec = importlib.util.spec_from_file_location("ref_ssd", REF)
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    return ref


def fuse(ref, block):
    """MobileNetV2.fuse_model (ssd_qmv2.py:178-185) applied to one block"""
    for m in block.modules():
        if type(m) == ref.ConvBNReLU:
            torch.quantization.fuse_modules(m, ['0', '1', '2'], inplace=True)
        if type(m) == ref.InvertedResidual:
            for idx in range(len(m.conv)):
                if type(m.conv[idx]) == nn.Conv2d:
                    torch.quantization.fuse_modules(m.conv, [str(idx), str(idx + 1)], inplace=True)


def main():
    ref = load_reference()
    out = []
    for ci, (inp, oup, s, t, H, d) in enumerate(CASES):
        torch.manual_seed(1882 + ci)
        net = nn.Sequential(torch.ao.quantization.QuantStub(), ref.InvertedResidual(inp, oup, s, d, t))
        float_sd = {k: v.clone() for k, v in net.state_dict().items()}
        net.train()
        fuse(ref, net[1])
        net.qconfig = torch.ao.quantization.get_default_qat_qconfig("qnnpack")
        torch.ao.quantization.prepare_qat(net, inplace=True)
        sd0 = {k: v.clone() for k, v in net.state_dict().items()}
        g = torch.Generator().manual_seed(31 + ci)
        steps = []
        for i in range(3):
            net.zero_grad()
            x = (torch.randn(4, inp, H, H, generator=g) * (1.0 + 0.5 * i)).requires_grad_(True)
            y = net(x)
            dy = torch.randn(y.shape, generator=g)
            y.backward(dy)
            steps.append(dict(x=x.detach().clone(), dy=dy, y=y.detach().clone(), dx=x.grad.clone(),
                              grads={n: p.grad.clone() for n, p in net.named_parameters()},
                              state={k: v.clone() for k, v in net.state_dict().items()}))
        out.append(dict(case=(inp, oup, s, t, H, d), float_sd=float_sd, sd0=sd0, steps=steps))
        print("case", ci, (inp, oup, s, t, d), "keys", len(sd0), "y", tuple(steps[0]["y"].shape))
    torch.save(dict(cases=out, torch=torch.__version__), os.path.join(HERE, "mbv2_block.pt"))


if __name__ == "__main__":
    main()
