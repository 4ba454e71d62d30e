"""QAT squeeze-and-excite (frostnet_b200/se.py + csrc/se.cu) against the reference's SEModule
(Classification/models/imagenet/mobilenetv3.py:85-102), golden vectors from tests/golden/make_golden_se.py.

Tolerances: the pooled mean and the two tiny GEMMs are fp32 sums in a different order than ATen's on the CPU (1e-7 relative
before each fake-quant); a value that lands within that distance of a rounding boundary may flip by one quantum.  The bar:
the broadcast product and every fake-quant are exact given the same gate, so y may differ from the reference by at most one
output quantum on at most 2 % of the elements; observer state 1e-5 relative; gradients 1e-4 relative L2."""
import pytest
import torch

from util import load_golden

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _net(C=16):
    import frostnet_b200 as F
    net = torch.nn.Sequential(F.QuantStub(), F.SEModule(C, reduction=4))
    net[1].fuse_model()
    F.attach_fake_quant(net)
    return net.to(DEV).train()


def _rel(a, b):
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def test_se_matches_reference_step_by_step():
    g = load_golden("se.pt")
    net = _net(g["C"])
    net.load_state_dict(g["sd0"])
    for i, s in enumerate(g["steps"]):
        if s.get("observers_off"):
            net.apply(torch.ao.quantization.disable_observer)
        net.zero_grad()
        x = s["x"].to(DEV).requires_grad_(True)
        y = net(x)
        sc, zp = y._frost_qparams
        quantum = float(s["state"]["1.quant_mul.activation_post_process.scale"])
        diff = (y.detach().cpu() - s["y"]).abs()
        assert float(diff.max()) <= 1.01 * quantum, (i, float(diff.max()), quantum)
        assert float((diff > 0.5 * quantum).float().mean()) <= 0.02, (i, float((diff > 0.5 * quantum).float().mean()))
        y.backward(s["dy"].to(DEV))
        assert _rel(x.grad.cpu(), s["dx"]) < 1e-4, (i, _rel(x.grad.cpu(), s["dx"]))
        assert _rel(net[1].fc[0].weight.grad.cpu(), s["dw0"]) < 1e-4, (i, "dw0", _rel(net[1].fc[0].weight.grad.cpu(), s["dw0"]))
        assert _rel(net[1].fc[2].weight.grad.cpu(), s["dw2"]) < 1e-4, (i, "dw2", _rel(net[1].fc[2].weight.grad.cpu(), s["dw2"]))
        sd = net.state_dict()
        for k, v in s["state"].items():
            a = sd[k].cpu()
            if v.dtype.is_floating_point:
                ok = torch.allclose(a, v, rtol=1e-5, atol=1e-7) or (not torch.isfinite(v).all() and torch.equal(a, v))
                assert ok, (i, k, a, v)
            else:
                assert int((a.long() - v.long()).abs().max()) <= (1 if k.endswith("zero_point") else 0), (i, k, a, v)
        # the result sits on the grid it advertises
        q = torch.round(y.detach() / sc) + zp
        assert float(q.min()) >= 0 and float(q.max()) <= 255
        assert float(((q - zp) * sc - y.detach()).abs().max()) <= 2e-7 * float(y.detach().abs().max()) + 1e-12


def test_se_replay_with_torch_ops_on_the_device():
    """MobileNetV3-sized SE (N=32, C=72 -> 18 -> 72 at 28x28: widths that are not multiples of 4 go through the padded path):
    replaying the block with plain torch ops from the module's updated state gives the same output up to gate rounding"""
    import frostnet_b200 as F
    torch.manual_seed(3)
    net = _net(72)
    x = torch.randn(32, 72, 28, 28, device=DEV) * 2
    y = net(x).detach()
    se = net[1]
    fqa = lambda t, f: (torch.clamp(torch.round(t * (1.0 / f.scale)) + f.zero_point, 0, 255) - f.zero_point) * f.scale
    fqw = lambda t, f: (torch.clamp(torch.round(t * (1.0 / f.scale)) + f.zero_point, -128, 127) - f.zero_point) * f.scale
    with torch.no_grad():
        xq = fqa(x, net[0].activation_post_process)
        p = xq.mean((2, 3))
        h = fqa(torch.relu(p @ fqw(se.fc[0].weight, se.fc[0].weight_fake_quant).t()), se.fc[0].activation_post_process)
        h = fqa(h @ fqw(se.fc[2].weight, se.fc[2].weight_fake_quant).t(), se.fc[2].activation_post_process)
        gate = fqa(torch.clamp(h + 3.0, 0, 6), se.fc[3].relu6.activation_post_process) * (1 / 6)
        ref = fqa(xq * gate[:, :, None, None], se.quant_mul.activation_post_process)
    quantum = float(se.quant_mul.activation_post_process.scale)
    diff = (y - ref).abs()
    assert float(diff.max()) <= 1.01 * quantum * 2 and float((diff > 0.5 * quantum).float().mean()) <= 0.01, \
        (float(diff.max()), float((diff > 0.5 * quantum).float().mean()))
    # gradients flow to the input and both weights
    xr = x.clone().requires_grad_(True)
    net(xr).sum().backward()
    assert xr.grad is not None and se.fc[0].weight.grad is not None and se.fc[2].weight.grad is not None
    assert torch.isfinite(xr.grad).all() and float(se.fc[2].weight.grad.abs().sum()) > 0


def test_bcast_mul_kernels_exact():
    """frost_bcast_mul_forward is torch.mul bit for bit; the backward's plane sums match torch to fp32 rounding"""
    from frostnet_b200 import _lib as L
    torch.manual_seed(1)
    x = torch.randn(7, 13, 9, 5, device=DEV)
    g = torch.rand(7, 13, device=DEV)
    dy = torch.randn_like(x)
    y, dx, dg = torch.empty_like(x), torch.empty_like(x), torch.empty_like(g)
    st = L.stream(x.device)
    L.call("frost_bcast_mul_forward", x.data_ptr(), g.data_ptr(), 7 * 13, 45, y.data_ptr(), st)
    L.call("frost_bcast_mul_backward", dy.data_ptr(), x.data_ptr(), g.data_ptr(), 7 * 13, 45, dx.data_ptr(), dg.data_ptr(), st)
    assert torch.equal(y, x * g[:, :, None, None])
    assert torch.equal(dx, dy * g[:, :, None, None])
    assert torch.allclose(dg, (dy * x).sum((2, 3)), rtol=1e-5, atol=1e-5)
    r, m = torch.empty_like(x), torch.empty(x.shape, dtype=torch.uint8, device=DEV)
    L.call("frost_relu_forward", x.data_ptr(), x.numel(), r.data_ptr(), m.data_ptr(), st)
    assert torch.equal(r, torch.relu(x)) and torch.equal(m.bool(), x > 0)


def test_functional_mul_standalone_and_errors():
    import frostnet_b200 as F
    ff = F.qat.FloatFunctional()
    a, b = torch.randn(2, 3, 4, 4), torch.rand(2, 3, 1, 1)
    assert torch.equal(ff.mul(a, b.expand_as(a)), a * b)           # float path
    net = torch.nn.Sequential(F.QuantStub())
    holder = torch.nn.Module()
    holder.ff = ff
    F.attach_fake_quant(net)
    F.attach_fake_quant(holder)
    net.to(DEV), holder.to(DEV)
    xq = net(torch.randn(2, 3, 4, 4, device=DEV))
    gate = torch.rand(2, 3, 1, 1, device=DEV)
    out = ff.mul(xq, gate.expand_as(xq))
    f = ff.activation_post_process
    ref = (torch.clamp(torch.round((xq * gate) * (1.0 / f.scale)) + f.zero_point, 0, 255) - f.zero_point) * f.scale
    assert torch.equal(out, ref)
    with pytest.raises(RuntimeError, match="supports"):
        ff.mul(xq, torch.rand(2, 3, 4, 4, device=DEV))
