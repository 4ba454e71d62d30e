"""CPU side of the QAT squeeze-and-excite block (frostnet_b200/se.py): the module tree, its state_dict keys before and after
fuse_model() + attach_fake_quant against the reference's SEModule (tests/golden/se.pt), and the float path."""
import pytest
import torch

from util import load_golden


def test_state_dict_keys_match_reference_float_and_prepared():
    import frostnet_b200 as F
    g = load_golden("se.pt")
    net = torch.nn.Sequential(F.QuantStub(), F.SEModule(g["C"], reduction=4))
    assert sorted(net.state_dict().keys()) == sorted(g["float_sd"].keys())
    net.load_state_dict(g["float_sd"], strict=True)
    net[1].fuse_model()
    F.attach_fake_quant(net)
    assert sorted(net.state_dict().keys()) == sorted(g["sd0"].keys())
    missing, unexpected = net.load_state_dict(g["sd0"], strict=True)
    assert not missing and not unexpected
    for k, v in g["sd0"].items():
        assert net.state_dict()[k].shape == v.shape and net.state_dict()[k].dtype == v.dtype, k


def test_float_forward_is_the_reference_formula():
    import frostnet_b200 as F
    g = load_golden("se.pt")
    se = F.SEModule(g["C"], reduction=4)
    se.load_state_dict({k[2:]: v for k, v in g["float_sd"].items() if k.startswith("1.")})
    x = g["steps"][0]["x"]
    w0, w2 = se.fc[0].weight, se.fc[2].weight
    gate = torch.nn.functional.relu6(torch.relu(x.mean((2, 3)) @ w0.t()) @ w2.t() + 3.0) / 6.0
    assert torch.allclose(se(x), x * gate[:, :, None, None], atol=1e-6)
    se.fuse_model()                                  # fusing does not change the float function
    assert torch.allclose(se(x), x * gate[:, :, None, None], atol=1e-6)


def test_prepared_module_refuses_cpu_tensors():
    import frostnet_b200 as F
    from frostnet_b200.block_engine import attach_qparams
    se = F.SEModule(8)
    se.fuse_model()
    F.attach_fake_quant(se)
    x = attach_qparams(torch.randn(1, 8, 2, 2), torch.ones(1), torch.zeros(1, dtype=torch.int32))
    with pytest.raises(RuntimeError, match="CUDA device"):
        se(x)


def test_mobilenetv3_bottleneck_keys_match_reference_float_and_prepared():
    """frostnet_b200.mobilenetv3.Bottleneck against the reference's Bottleneck (tests/golden/mbv3_block.pt): same state_dict keys,
    shapes and dtypes before fuse_model() and after fuse_model() + attach_fake_quant; the float forward runs"""
    import frostnet_b200 as F
    from frostnet_b200 import mobilenetv3 as M
    g = load_golden("mbv3_block.pt")
    for c in g["cases"]:
        cin, cout, exp, k, s, se, nl, H = c["case"]
        net = torch.nn.Sequential(F.QuantStub(), M.Bottleneck(cin, cout, exp, k, s, se=se, nl=nl))
        assert sorted(net.state_dict().keys()) == sorted(c["float_sd"].keys()), c["case"]
        net.load_state_dict(c["float_sd"], strict=True)
        y = net(c["steps"][0]["x"])
        assert y.shape == c["steps"][0]["y"].shape
        net[1].fuse_model()
        F.attach_fake_quant(net)
        assert sorted(net.state_dict().keys()) == sorted(c["sd0"].keys()), c["case"]
        net.load_state_dict(c["sd0"], strict=True)
        sd = net.state_dict()
        for kk, v in c["sd0"].items():
            assert sd[kk].shape == v.shape and sd[kk].dtype == v.dtype, kk
        assert [n for n, _ in net.named_parameters()] == list(c["steps"][0]["grads"].keys())


def test_mobilenetv3_network_keys_match_reference():
    """frostnet_b200.mobilenetv3.MobileNetV3 against the reference network (tests/golden/mbv3_net.pt): the same state_dict keys
    in the same order before fuse_model(); the same keys, shapes, dtypes and parameter order after fuse + prepare; the factory
    names and the float forward work"""
    import frostnet_b200 as F
    from frostnet_b200 import mobilenetv3 as M
    g = load_golden("mbv3_net.pt")
    net = M.get_mobilenet_v3("small", 1.0, nclass=10)
    assert list(net.state_dict().keys()) == g["float_keys"]
    net.eval()
    assert net(torch.randn(2, 3, 64, 64)).shape == (2, 10)
    net.train()
    net.fuse_model()
    F.attach_fake_quant(net)
    sd = net.state_dict()
    assert sorted(sd.keys()) == sorted(g["sd0_keys"].keys())
    for k, (shape, dtype) in g["sd0_keys"].items():
        assert tuple(sd[k].shape) == shape and str(sd[k].dtype) == dtype, k
    assert [n for n, _ in net.named_parameters()] == g["param_names"]
    for f in (M.mobilenet_v3_large, M.mobilenet_v3_ReLU_small, M.mobilenet_v3_ReLU_large):
        assert sum(p.numel() for p in f(nclass=7).parameters()) > 0
    with pytest.raises(ValueError):
        M.MobileNetV3(mode="medium")


def test_mobilenetv2_block_keys_match_reference():
    """frostnet_b200.mobilenetv2.InvertedResidual (SSDLite backbone, f2) against the reference's block
    (tests/golden/mbv2_block.pt): state_dict keys, shapes, dtypes and parameter order before and after fuse + prepare"""
    import frostnet_b200 as F
    from frostnet_b200 import mobilenetv2 as M2
    g = load_golden("mbv2_block.pt")
    for c in g["cases"]:
        inp, oup, s, t, H, d = c["case"]
        net = torch.nn.Sequential(F.QuantStub(), M2.InvertedResidual(inp, oup, s, d, t))
        assert list(net.state_dict().keys()) == list(c["float_sd"].keys()), c["case"]
        net.load_state_dict(c["float_sd"], strict=True)
        assert net(c["steps"][0]["x"]).shape == c["steps"][0]["y"].shape
        M2.fuse_model(net)
        F.attach_fake_quant(net)
        assert sorted(net.state_dict().keys()) == sorted(c["sd0"].keys()), c["case"]
        net.load_state_dict(c["sd0"], strict=True)
        for kk, v in c["sd0"].items():
            assert net.state_dict()[kk].shape == v.shape and net.state_dict()[kk].dtype == v.dtype, kk
        assert [n for n, _ in net.named_parameters()] == list(c["steps"][0]["grads"].keys())
