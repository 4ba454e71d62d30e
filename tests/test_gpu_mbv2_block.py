"""The MobileNetV2 inverted-residual block of the SSDLite backbone (frostnet_b200/mobilenetv2.py) on the
per-module executor against the reference's classes (Object_Detection/ssd_qmv2.py:40-110), golden vectors from
tests/golden/make_golden_mbv2_block.py: three QAT training steps per configuration (t = 1 / stride 2 / residual / dilation 2 with and without the residual).  Every fused
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


@pytest.mark.parametrize("ci", [0, 1, 2, 3, 4])
def test_inverted_residual_matches_reference_step_by_step(ci):
    import frostnet_b200 as F
    from frostnet_b200 import mobilenetv2 as M2
    c = load_golden("mbv2_block.pt")["cases"][ci]
    inp, oup, s, t, H, d = c["case"]
    net = torch.nn.Sequential(F.QuantStub(), M2.InvertedResidual(inp, oup, s, d, t))
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


@pytest.mark.parametrize("k,stride,dil,C,H,W,pad", [(3, 1, 2, 24, 10, 10, None), (5, 1, 2, 16, 12, 9, None), (3, 2, 3, 8, 11, 13, None),
                                                      (5, 2, 4, 40, 19, 19, None), (3, 1, 8, 12, 20, 20, None),
                                                      (3, 1, 1, 16, 10, 10, 0), (3, 2, 1, 24, 9, 11, 1), (5, 1, 1, 8, 9, 9, 0), (3, 1, 2, 8, 12, 12, 1)])
def test_dilated_depthwise_kernels_against_torch(k, stride, dil, C, H, W, pad):
    """frost_dw_conv_forward_dilated / _dgrad_dilated / _wgrad_dilated against F.conv2d and its autograd on integer-valued
    tensors: the accumulators and their per-channel statistics exactly, the gradients to fp32 rounding"""
    import ctypes
    import torch.nn.functional as Fn
    from frostnet_b200 import _lib as L
    torch.manual_seed(k * 100 + dil)
    N = 3
    pad = dil * (k - 1) // 2 if pad is None else pad            # None: the reference's 'same' padding; else explicit (SSD extras: 0)
    xq = torch.randint(0, 256, (N, H, W, C), dtype=torch.uint8, device=DEV)
    wq = torch.randint(-128, 128, (k * k, C), dtype=torch.int8, device=DEV)
    zp_a = torch.tensor([117], dtype=torch.int32, device=DEV)
    zp_w = torch.tensor([0], dtype=torch.int32, device=DEV)
    xf = (xq.double() - 117).permute(0, 3, 1, 2).contiguous().requires_grad_(True)
    wf = wq.double().t().reshape(C, 1, k, k).contiguous().requires_grad_(True)
    ref = Fn.conv2d(xf, wf, None, stride, pad, dil, groups=C)
    Ho, Wo = ref.shape[2:]
    acc = torch.empty((N, Ho, Wo, C), dtype=torch.int32, device=DEV)
    stats = torch.zeros(C * 8, dtype=torch.int32, device=DEV)       # FrostChanStats: 32 bytes per channel
    st = L.stream(acc.device)
    L.call("frost_stats_reset", stats.data_ptr(), C, st)
    L.call("frost_dw_conv_forward_dilated", xq.data_ptr(), C, zp_a.data_ptr(), wq.data_ptr(), zp_w.data_ptr(), N, H, W, C, k, stride, dil,
           pad, acc.data_ptr(), stats.data_ptr(), st)
    assert torch.equal(acc.permute(0, 3, 1, 2).double(), ref.detach())
    dz = torch.randn(N, Ho, Wo, C, device=DEV)
    ref.backward(dz.permute(0, 3, 1, 2).double())
    w_scale = torch.tensor([0.02], device=DEV)
    x_scale = torch.tensor([0.05], device=DEV)
    dx = torch.empty((N, H, W, C), device=DEV)
    L.call("frost_dw_dgrad_dilated", dz.data_ptr(), wq.data_ptr(), w_scale.data_ptr(), zp_w.data_ptr(), N, H, W, C, k, stride, dil,
           pad, dx.data_ptr(), 0, st)
    want_dx = (xf.grad * 0.02).permute(0, 2, 3, 1).float()
    assert _rel(dx, want_dx) < 1e-6, _rel(dx, want_dx)
    L.call("frost_dw_dgrad_dilated", dz.data_ptr(), wq.data_ptr(), w_scale.data_ptr(), zp_w.data_ptr(), N, H, W, C, k, stride, dil,
           pad, dx.data_ptr(), 1, st)
    assert _rel(dx, 2 * want_dx) < 1e-6
    dwq = torch.empty((k * k, C), device=DEV)
    L.call("frost_dw_wgrad_dilated", dz.data_ptr(), xq.data_ptr(), C, x_scale.data_ptr(), zp_a.data_ptr(), N, H, W, C, k, stride, dil,
           pad, dwq.data_ptr(), st)
    want_dw = (wf.grad.reshape(C, k * k).t() * 0.05).float()
    assert _rel(dwq, want_dw) < 1e-5, _rel(dwq, want_dw)


def test_mobilenetv2_backbone_trains_end_to_end():
    """The whole SSDLite backbone (stem, 17 inverted-residual blocks - the last 4 with dilated depthwise convs -, 1x1 to 1280) in
    QAT training mode on the device: output stride 16, every parameter receives a finite gradient, a few SGD steps on a fixed
    batch keep a regression loss finite and do not increase it"""
    import frostnet_b200 as F
    from frostnet_b200 import mobilenetv2 as M2
    torch.manual_seed(0)
    net = torch.nn.Sequential(F.QuantStub(), M2.MobileNetV2())
    net[1].fuse_model()
    F.attach_fake_quant(net)
    net.to(DEV).train()
    x = torch.randn(4, 3, 96, 96, device=DEV)
    target = torch.randn(4, 1280, 6, 6, device=DEV).abs() * 0.1
    opt = torch.optim.SGD(net.parameters(), lr=0.02, momentum=0.9)
    losses = []
    for i in range(6):
        opt.zero_grad()
        y = net(x)
        assert y.shape == (4, 1280, 6, 6) and hasattr(y, "_frost_qparams")
        loss = (y - target).square().mean()
        loss.backward()
        assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in net.parameters()), i
        opt.step()
        losses.append(float(loss.detach()))
    print("mbv2 backbone losses", losses)
    assert all(l == l and l < 10 for l in losses) and min(losses[1:]) < losses[0] * 1.02, losses


def test_valid_padded_depthwise_block_runs_on_the_executor():
    """the SSD extras' depthwise ConvBN (ssd_qmv2.py:188-203: 3x3, padding 0 at stride 1, padding 1 at stride 2) as a stand-alone
    fused conv: output sizes of a 'valid' convolution, gradients reach every parameter"""
    import frostnet_b200 as F
    stub = torch.nn.Sequential(F.QuantStub())
    F.attach_fake_quant(stub)
    stub.to(DEV)
    for stride, pad, want in ((1, 0, 8), (2, 1, 5)):
        blk = F.ConvBN(16, 16, 3, stride, pad, 1, groups=16)
        blk.fuse_model()
        F.attach_fake_quant(blk)
        blk.to(DEV).train()
        x = stub(torch.randn(2, 16, 10, 10, device=DEV))
        y = blk(x)
        ref = torch.nn.functional.conv2d(x.detach(), blk.conv[0].weight.detach(), None, stride, pad, 1, 16)
        assert y.shape == ref.shape == (2, 16, want, want)
        y.sum().backward()
        assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in blk.parameters())
