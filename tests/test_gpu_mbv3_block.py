"""The MobileNetV3 inverted-residual block (frostnet_b200/mobilenetv3.py) on the per-module executor against
the reference's Bottleneck (Classification/models/imagenet/mobilenetv3.py:113-160), golden vectors from
tests/golden/make_golden_mbv3_block.py: three QAT training steps per configuration (plain ReLU / hard-swish, with / without
squeeze-and-excite, stride 1 with the residual / stride 2).

Tolerances: every member is exact on equal inputs except the fp32 sums whose order differs from ATen's CPU kernels (BatchNorm
statistics, the SE pool and GEMMs); a value within 1e-7 relative of a rounding boundary flips by one quantum and the flip
travels on.  Measured on B200: all nine outputs BIT-IDENTICAL to the reference's, input gradients 2e-6 .. 7e-5, parameter
gradients 6e-6 .. 4.5e-4 relative L2.  Asserted: at most one quantum on at most 0.5 % of the elements; observer / BatchNorm
state 1e-4; gradients 2e-3."""
import pytest
import torch

from util import load_golden

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _rel(a, b):
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


@pytest.mark.parametrize("ci", [0, 1, 2])
def test_bottleneck_matches_reference_step_by_step(ci):
    import frostnet_b200 as F
    from frostnet_b200 import mobilenetv3 as M
    c = load_golden("mbv3_block.pt")["cases"][ci]
    cin, cout, exp, k, s, se, nl, H = c["case"]
    net = torch.nn.Sequential(F.QuantStub(), M.Bottleneck(cin, cout, exp, k, s, se=se, nl=nl))
    net[1].fuse_model()
    F.attach_fake_quant(net)
    net.load_state_dict(c["sd0"], strict=True)
    net.to(DEV).train()
    last = "1.skip_add" if net[1].use_res_connect else "1.conv.4.cb.0"
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
        print("case %d step %d: max |dy| %.2f quanta, %.3f %% of the elements off; dx %.2e; worst grad %s %.2e"
              % (ci, i, float(diff.max()) / quantum, 100 * frac, dxerr, worst, gerr[worst]))
        assert float(diff.max()) <= 1.01 * quantum and frac <= 0.005, (ci, i, float(diff.max()) / quantum, frac)
        assert dxerr < 2e-3 and gerr[worst] < 2e-3, (ci, i, dxerr, worst, gerr[worst])
        sd = net.state_dict()
        for kk, v in st["state"].items():
            a = sd[kk].cpu()
            if v.dtype.is_floating_point:
                fin = torch.isfinite(v)
                assert torch.equal(torch.isfinite(a), fin), (ci, i, kk)
                assert torch.allclose(a[fin], v[fin], rtol=1e-4, atol=1e-6), (ci, i, kk, float((a[fin] - v[fin]).abs().max()))
            else:
                assert int((a.long() - v.long()).abs().max()) <= (1 if kk.endswith("zero_point") else 0), (ci, i, kk, a, v)


def test_relu_keeps_the_grid_and_dilated_depthwise_is_refused():
    import frostnet_b200 as F
    from frostnet_b200 import mobilenetv3 as M
    stub = torch.nn.Sequential(F.QuantStub())
    F.attach_fake_quant(stub)
    stub.to(DEV)
    xq = stub(torch.randn(2, 8, 5, 5, device=DEV, requires_grad=True))
    r = M.ReLU(True)(xq)
    assert torch.equal(r.detach(), torch.relu(xq.detach())) and r._frost_qparams[0] is xq._frost_qparams[0]
    r.sum().backward()
    blk = M._ConvBN(8, 8, 3, 1, 2, 2, groups=8)
    blk.fuse_model()
    F.attach_fake_quant(blk)
    blk.to(DEV).train()
    with pytest.raises(RuntimeError, match="dilated"):
        blk(stub(torch.randn(2, 8, 5, 5, device=DEV)))
