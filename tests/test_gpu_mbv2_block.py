"""The MobileNetV2 inverted-residual block of the SSDLite backbone (frostnet_b200/mobilenetv2.py) on the
per-module executor against the reference's classes (Object_Detection/ssd_qmv2.py:40-110), golden vectors from
tests/golden/make_golden_mbv2_block.py: three QAT training steps per configuration (t = 1 / stride 2 / residual).  Every fused
conv here is a prepared FrostConvBn2d called directly by the nn.Sequential that holds it.
Asserted: outputs within one quantum on at most 0.5 % of the elements, observer / BatchNorm state 1e-4 (1e-5 absolute), gradients 2e-3
relative L2 (the measured values are printed)."""
import pytest
import torch

from util import load_golden

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _rel(a, b):
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


@pytest.mark.parametrize("ci", [0, 1, 2])
def test_inverted_residual_matches_reference_step_by_step(ci):
    import frostnet_b200 as F
    from frostnet_b200 import mobilenetv2 as M2
    c = load_golden("mbv2_block.pt")["cases"][ci]
    inp, oup, s, t, H = c["case"]
    net = torch.nn.Sequential(F.QuantStub(), M2.InvertedResidual(inp, oup, s, 1, t))
    M2.fuse_model(net)
    F.attach_fake_quant(net)
    net.load_state_dict(c["sd0"], strict=True)
    net.to(DEV).train()
    last = "1.skip_add" if net[1].use_res_connect else "1.conv.%d" % (len(net[1].conv) - 2)
    for i, st in enumerate(c["steps"]):
        net.zero_grad()
        x = st["x"].to(DEV).requires_grad_(True)
        y = net(x)
        assert hasattr(y, "_frost_qparams")
        quantum = float(st["state"][last + ".activation_post_process.scale"])
        diff = (y.detach().cpu() - st["y"]).abs()
        frac = float((diff > 0.5 * quantum).float().mean())
        y.backward(st["dy"].to(DEV))
        gerr = {n: _rel(p.grad.cpu(), st["grads"][n]) for n, p in net.named_parameters()}
        worst = max(gerr, key=gerr.get)
        dxerr = _rel(x.grad.cpu(), st["dx"])
        print("mbv2 case %d step %d: max |dy| %.2f quanta, %.3f %% of the elements off; dx %.2e; worst grad %s %.2e"
              % (ci, i, float(diff.max()) / quantum, 100 * frac, dxerr, worst, gerr[worst]))
        assert float(diff.max()) <= 1.01 * quantum and frac <= 0.005, (ci, i, float(diff.max()) / quantum, frac)
        assert dxerr < 2e-3 and gerr[worst] < 2e-3, (ci, i, dxerr, worst, gerr[worst])
        sd = net.state_dict()
        for kk, v in st["state"].items():
            a = sd[kk].cpu()
            if v.dtype.is_floating_point:
                fin = torch.isfinite(v)
                assert torch.equal(torch.isfinite(a), fin), (ci, i, kk)
                assert torch.allclose(a[fin], v[fin], rtol=1e-4, atol=1e-5), (ci, i, kk, float((a[fin] - v[fin]).abs().max()))
            else:
                assert int((a.long() - v.long()).abs().max()) <= (1 if kk.endswith("zero_point") else 0), (ci, i, kk, a, v)


def test_dilated_block_is_refused_in_qat():
    import frostnet_b200 as F
    from frostnet_b200 import mobilenetv2 as M2
    net = torch.nn.Sequential(F.QuantStub(), M2.InvertedResidual(16, 16, 1, 2, 6))
    M2.fuse_model(net)
    F.attach_fake_quant(net)
    net.to(DEV).train()
    with pytest.raises(RuntimeError, match="dilated"):
        net(torch.randn(2, 16, 8, 8, device=DEV))
