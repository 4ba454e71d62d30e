"""QAT hard-swish (csrc/hswish.cu) against the reference's _Hswish module
(Classification/models/imagenet/mobilenetv3.py:43-56), golden vectors from tests/golden/make_golden_hswish.py."""
import pytest
import torch

from util import load_golden

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _net():
    import frostnet_b200 as F
    net = torch.nn.Sequential(F.QuantStub(), F.Hswish(True))
    F.attach_fake_quant(net)
    return net.to(DEV).train()


def test_state_dict_keys_match_reference():
    g = load_golden("hswish.pt")
    net = _net()
    assert sorted(net.state_dict().keys()) == sorted(g["sd0"].keys())
    missing, unexpected = net.load_state_dict(g["sd0"], strict=True)
    assert not missing and not unexpected


def test_hswish_matches_reference_step_by_step():
    """outputs BIT-EXACT (every stage is a function of the uint8 input index), both observers' state exact, gradients to fp32
    rounding; the last step runs with the observers switched off"""
    import frostnet_b200 as F
    g = load_golden("hswish.pt")
    net = _net()
    net.load_state_dict(g["sd0"])
    for i, s in enumerate(g["steps"]):
        if s.get("observers_off"):
            net.apply(torch.ao.quantization.disable_observer)
        x = s["x"].to(DEV).requires_grad_(True)
        y = net(x)
        assert hasattr(y, "_frost_qparams")
        assert torch.equal(y.detach().cpu(), s["y"]), (i, float((y.detach().cpu() - s["y"]).abs().max()))
        y.backward(s["dy"].to(DEV))
        dx = x.grad.cpu()
        assert float((dx - s["dx"]).abs().max()) <= 1e-6 * float(s["dx"].abs().max()), (i, float((dx - s["dx"]).abs().max()))
        sd = net.state_dict()
        for k, v in s["state"].items():
            assert torch.equal(sd[k].cpu(), v), (i, k, sd[k], v)
        # the result sits on the grid it advertises
        sc, zp = y._frost_qparams
        q = torch.round(y.detach() / sc) + zp
        assert float(((q - zp) * sc - y.detach()).abs().max()) <= 2e-7 * float(y.detach().abs().max()) + 1e-12
        assert float(q.min()) >= 0 and float(q.max()) <= 255


def test_hswish_at_a_headline_sized_tensor():
    """256 x 112 x 112 x 16 elements (51 M): table path == the same ops done one by one with torch on the device"""
    import frostnet_b200 as F
    net = _net()
    torch.manual_seed(0)
    x = torch.randn(256, 16, 112, 112, device=DEV) * 3
    y = net(x)
    # replay with torch ops from the module's (now updated) state
    stub_fq, fa, fb = net[0].activation_post_process, net[1].relu6.activation_post_process, net[1].quant_mul1.activation_post_process
    fq = lambda t, f: (torch.clamp(torch.round(t * (1.0 / f.scale)) + f.zero_point, 0, 255) - f.zero_point) * f.scale
    xq = fq(x, stub_fq)
    ra = fq(torch.clamp(xq + 3.0, 0, 6), fa)
    ref = fq(xq * ra, fb) * (1 / 6)
    assert torch.equal(y, ref), float((y - ref).abs().max())


def test_hswish_float_path_and_errors():
    import frostnet_b200 as F
    m = F.Hswish()
    x = torch.randn(4, 5)
    assert torch.allclose(m(x), x * torch.nn.functional.relu6(x + 3) / 6, atol=1e-6)
    net = _net()
    with pytest.raises(RuntimeError, match="without quantisation parameters"):
        net[1](torch.randn(2, 3, device=DEV))


def test_hsigmoid_matches_reference_step_by_step():
    """_Hsigmoid (mobilenetv3.py:59-69) on [N, C] tensors as the SE block feeds it; last step with the observers off"""
    import frostnet_b200 as F
    g = load_golden("hswish.pt")
    net = torch.nn.Sequential(F.QuantStub(), F.Hsigmoid(True))
    F.attach_fake_quant(net)
    net.to(DEV).train()
    assert sorted(net.state_dict().keys()) == sorted(g["sig_sd0"].keys())
    net.load_state_dict(g["sig_sd0"])
    for i, s in enumerate(g["sig_steps"]):
        if s.get("observers_off"):
            net.apply(torch.ao.quantization.disable_observer)
        x = s["x"].to(DEV).requires_grad_(True)
        y = net(x)
        assert torch.equal(y.detach().cpu(), s["y"]), (i, float((y.detach().cpu() - s["y"]).abs().max()))
        y.backward(s["dy"].to(DEV))
        assert float((x.grad.cpu() - s["dx"]).abs().max()) <= 1e-6 * float(s["dx"].abs().max()) + 1e-12, i
        sd = net.state_dict()
        for k, v in s["state"].items():
            assert torch.equal(sd[k].cpu(), v), (i, k)
