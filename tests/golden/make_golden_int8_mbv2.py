"""Golden vectors for the int8 export of the SSDLite backbone block (SURVEY.md 8f, f1 + f2): the REAL reference InvertedResidual
(Object_Detection/ssd_qmv2.py:80-110; residual, dilation 2) behind a QuantStub, three QAT steps, then torch.quantization.convert
and the int8 output.  Runs only in the build container; tests/golden/int8_mbv2.pt is committed.

    python tests/golden/make_golden_int8_mbv2.py
"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(HERE, ".."))
import make_golden_mbv2_block as G  # noqa: E402
from util import qdigest_compact    # noqa: E402


def main():
    ref = G.load_reference()
    torch.backends.quantized.engine = "qnnpack"
    torch.manual_seed(7)
    case = (24, 24, 1, 6, 10, 2)
    inp, oup, s, t, H, d = case
    net = torch.nn.Sequential(torch.ao.quantization.QuantStub(), ref.InvertedResidual(inp, oup, s, d, t), torch.ao.quantization.DeQuantStub())
    net.train()
    G.fuse(ref, net[1])
    net.qconfig = torch.ao.quantization.get_default_qat_qconfig("qnnpack")
    torch.ao.quantization.prepare_qat(net, inplace=True)
    g = torch.Generator().manual_seed(5)
    for i in range(3):
        net.zero_grad()
        y = net(torch.randn(4, inp, H, H, generator=g))
        y.backward(torch.randn(y.shape, generator=g))
        with torch.no_grad():
            for p in net.parameters():
                p.add_(p.grad, alpha=-0.05)
    x = torch.randn(4, inp, H, H, generator=g)
    net.eval()
    with torch.no_grad():
        net(x)                                         # observers stay on in eval mode: this forward moves their state
        sd = {k: v.clone() for k, v in net.state_dict().items()}
        q = torch.ao.quantization.convert(net, inplace=False)
        out = q(x).clone()
    torch.save(dict(case=case, sd=sd, x=x, int8_out=out, converted=qdigest_compact(q.state_dict()), engine="qnnpack",
                    torch=torch.__version__), os.path.join(HERE, "int8_mbv2.pt"))
    print("int8 mbv2 golden ok:", len(sd), "state entries,", tuple(out.shape))


if __name__ == "__main__":
    main()
