"""Golden vectors for the QAT squeeze-and-excite block: the REAL reference module
(Classification/models/imagenet/mobilenetv3.py:85-102 SEModule) behind a QuantStub, fused by its own fuse_model(),
prepared with the qnnpack QAT qconfig, three training steps + one with the observers off.  Runs only in the build container;
tests/golden/se.pt is committed.

    python tests/golden/make_golden_se.py
"""
import importlib.util
import os
import warnings

import torch

warnings.filterwarnings("ignore")
HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/Classification/models/imagenet/mobilenetv3.py"


def main():
    # the reference fuses in train mode (torch 1.6 picked the QAT fuser from module.training; torch >= 1.11 needs fuse_modules_qat)
    torch.quantization.fuse_modules = torch.ao.quantization.fuse_modules_qat
    spec = importlib.util.spec_from_file_location("ref_mbv3", REF)
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    torch.manual_seed(1882)
    C = 16
    net = torch.nn.Sequential(torch.ao.quantization.QuantStub(), ref.SEModule(C, reduction=4))
    with torch.no_grad():                       # default init gives gates that barely move; spread them
        net[1].fc[0].weight.mul_(3.0)
        net[1].fc[2].weight.mul_(6.0)
    float_sd = {k: v.clone() for k, v in net.state_dict().items()}
    net.train()
    net[1].fuse_model()
    net.qconfig = torch.ao.quantization.get_default_qat_qconfig("qnnpack")
    torch.ao.quantization.prepare_qat(net, inplace=True)
    sd0 = {k: v.clone() for k, v in net.state_dict().items()}
    g = torch.Generator().manual_seed(7)
    steps = []
    for i in range(4):
        if i == 3:
            net.apply(torch.ao.quantization.disable_observer)
        net.zero_grad()
        x = (torch.randn(3, C, 5, 6, generator=g) * (1.5 + i) + 0.3 * i).requires_grad_(True)
        dy = torch.randn(3, C, 5, 6, generator=g)
        y = net(x)
        y.backward(dy)
        steps.append(dict(x=x.detach().clone(), dy=dy, y=y.detach().clone(), dx=x.grad.clone(),
                          dw0=net[1].fc[0].weight.grad.clone(), dw2=net[1].fc[2].weight.grad.clone(),
                          state={k: v.clone() for k, v in net.state_dict().items()}, observers_off=(i == 3)))
    torch.save(dict(float_sd=float_sd, sd0=sd0, steps=steps, C=C, torch=torch.__version__), os.path.join(HERE, "se.pt"))
    print("se golden ok: keys", len(sd0))
    for k in sd0:
        print("  ", k, tuple(sd0[k].shape))
    print(net)


if __name__ == "__main__":
    main()
