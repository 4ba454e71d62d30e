"""Golden vectors for the MobileNetV3 inverted-residual block: the REAL reference Bottleneck
(Classification/models/imagenet/mobilenetv3.py:113-160) behind a QuantStub, fused by its own fuse_model(), prepared with the
qnnpack QAT qconfig, three training steps per configuration.  Runs only in the build container; tests/golden/mbv3_block.pt is
committed.

    python tests/golden/make_golden_mbv3_block.py
"""
import importlib.util
import os
import warnings

import torch

warnings.filterwarnings("ignore")
HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/Classification/models/imagenet/mobilenetv3.py"
#        in, out, exp, k, stride, se, nl, H
CASES = [(16, 16, 64, 3, 1, True, 'RE', 12),      # residual + SE + plain ReLU
         (16, 24, 48, 5, 2, False, 'HS', 12),     # stride 2, hard-swish, no SE
         (24, 24, 72, 5, 1, True, 'HS', 8)]       # residual + SE + hard-swish (exp/4 = 18: the padded Linear path)


def main():
    torch.quantization.fuse_modules = torch.ao.quantization.fuse_modules_qat
    spec = importlib.util.spec_from_file_location("ref_mbv3", REF)
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    out = []
    for ci, (cin, cout, exp, k, s, se, nl, H) in enumerate(CASES):
        torch.manual_seed(1882 + ci)
        net = torch.nn.Sequential(torch.ao.quantization.QuantStub(), ref.Bottleneck(cin, cout, exp, k, s, se=se, nl=nl))
        float_sd = {kk: v.clone() for kk, v in net.state_dict().items()}
        net.train()
        net[1].fuse_model()
        net.qconfig = torch.ao.quantization.get_default_qat_qconfig("qnnpack")
        torch.ao.quantization.prepare_qat(net, inplace=True)
        sd0 = {kk: v.clone() for kk, v in net.state_dict().items()}
        g = torch.Generator().manual_seed(11 + ci)
        steps = []
        for i in range(3):
            net.zero_grad()
            x = (torch.randn(4, cin, H, H, generator=g) * (1.0 + 0.5 * i)).requires_grad_(True)
            y = net(x)
            dy = torch.randn(y.shape, generator=g)
            y.backward(dy)
            steps.append(dict(x=x.detach().clone(), dy=dy, y=y.detach().clone(), dx=x.grad.clone(),
                              grads={n: p.grad.clone() for n, p in net.named_parameters()},
                              state={kk: v.clone() for kk, v in net.state_dict().items()}))
        out.append(dict(case=(cin, cout, exp, k, s, se, nl, H), float_sd=float_sd, sd0=sd0, steps=steps))
        print("case", ci, (cin, cout, exp, k, s, se, nl), "keys", len(sd0), "y", tuple(steps[0]["y"].shape))
    torch.save(dict(cases=out, torch=torch.__version__), os.path.join(HERE, "mbv3_block.pt"))


if __name__ == "__main__":
    main()
