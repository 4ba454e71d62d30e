"""Golden vectors for the QAT hard-swish: the REAL reference module
(Classification/models/imagenet/mobilenetv3.py:43-56 _Hswish) behind a QuantStub, prepared with the qnnpack QAT qconfig,
three training steps.  Runs only in the build container; tests/golden/hswish.pt is committed.

    python tests/golden/make_golden_hswish.py
"""
import importlib.util
import os
import warnings

import torch

warnings.filterwarnings("ignore")
HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/Classification/models/imagenet/mobilenetv3.py"


def main():
    spec = importlib.util.spec_from_file_location("ref_mbv3", REF)
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    torch.manual_seed(1882)
    net = torch.nn.Sequential(torch.ao.quantization.QuantStub(), ref._Hswish(True))
    net.train()
    net.qconfig = torch.ao.quantization.get_default_qat_qconfig("qnnpack")
    torch.ao.quantization.prepare_qat(net, inplace=True)
    sd0 = {k: v.clone() for k, v in net.state_dict().items()}
    g = torch.Generator().manual_seed(5)
    steps = []
    for i in range(3):
        x = (torch.randn(3, 8, 9, 7, generator=g) * (2.0 + i) + 0.5 * i).requires_grad_(True)
        dy = torch.randn(3, 8, 9, 7, generator=g)
        y = net(x)
        y.backward(dy)
        steps.append(dict(x=x.detach().clone(), dy=dy, y=y.detach().clone(), dx=x.grad.clone(),
                          state={k: v.clone() for k, v in net.state_dict().items()}))
    # observers off (late QAT): the tables still follow the frozen qparams
    net.apply(torch.ao.quantization.disable_observer)
    x = (torch.randn(3, 8, 9, 7, generator=g) * 5.0).requires_grad_(True)
    dy = torch.randn(3, 8, 9, 7, generator=g)
    y = net(x)
    y.backward(dy)
    steps.append(dict(x=x.detach().clone(), dy=dy, y=y.detach().clone(), dx=x.grad.clone(),
                      state={k: v.clone() for k, v in net.state_dict().items()}, observers_off=True))
    # the same for _Hsigmoid (mobilenetv3.py:59-69)
    torch.manual_seed(1883)
    net2 = torch.nn.Sequential(torch.ao.quantization.QuantStub(), ref._Hsigmoid(True))
    net2.train()
    net2.qconfig = torch.ao.quantization.get_default_qat_qconfig("qnnpack")
    torch.ao.quantization.prepare_qat(net2, inplace=True)
    sig_sd0 = {k: v.clone() for k, v in net2.state_dict().items()}
    sig_steps = []
    for i in range(3):
        if i == 2:
            net2.apply(torch.ao.quantization.disable_observer)
        x = (torch.randn(5, 12, generator=g) * (2.0 + i) - 0.5 * i).requires_grad_(True)       # SE feeds it [N, C]
        dy = torch.randn(5, 12, generator=g)
        y = net2(x)
        y.backward(dy)
        sig_steps.append(dict(x=x.detach().clone(), dy=dy, y=y.detach().clone(), dx=x.grad.clone(),
                              state={k: v.clone() for k, v in net2.state_dict().items()}, observers_off=(i == 2)))
    torch.save(dict(sd0=sd0, steps=steps, sig_sd0=sig_sd0, sig_steps=sig_steps, torch=torch.__version__), os.path.join(HERE, "hswish.pt"))
    print("hswish golden ok:", [tuple(s["y"].shape) for s in steps], "keys", len(sd0), "| hsigmoid:", len(sig_steps), "steps, keys", len(sig_sd0))


if __name__ == "__main__":
    main()
