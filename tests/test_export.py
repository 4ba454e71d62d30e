"""int8 inference export: frostnet_b200.convert_int8 against the reference's own
torch.quantization.convert (Classification/evaluate.py:131), golden vectors from tests/golden/make_golden_int8.py."""
import pytest
import torch

from util import load_golden


def _digest(sd):
    out = {}
    for k, v in sd.items():
        if isinstance(v, torch.Tensor) and v.is_quantized:
            out[k] = dict(int_repr=v.int_repr(), scale=float(v.q_scale()), zero_point=int(v.q_zero_point()))
        elif isinstance(v, torch.Tensor):
            out[k] = v
        elif isinstance(v, (int, float, type(None), torch.dtype)):
            out[k] = v
        else:
            out[k] = repr(v)
    return out


@pytest.fixture(scope="module")
def converted():
    import frostnet_b200 as F
    g = load_golden("int8_small035.pt")
    if g["engine"] not in torch.backends.quantized.supported_engines:
        pytest.skip("quantized engine %s not available in this torch build" % g["engine"])
    torch.backends.quantized.engine = g["engine"]
    model = F.FrostNet(nclass=g["nclass"], mode=g["mode"], width_mult=g["width_mult"], quantized=True, drop_rate=0.0)
    model.fuse_model()
    F.prepare_qat(model)
    missing, unexpected = model.load_state_dict(g["sd"], strict=True)
    assert not missing and not unexpected
    before = {k: v.clone() for k, v in model.state_dict().items()}
    q = F.convert_int8(model)
    after = model.state_dict()
    assert all(torch.equal(before[k], after[k]) for k in before), "convert_int8 must not touch the QAT model"
    assert model.__dict__.get("_frost_engine") is not None and q.__dict__.get("_frost_engine") is None
    return g, q


def test_converted_state_matches_reference_convert(converted):
    """every quantized weight (int8 indices, scale, zero point), bias and output qparam of the reference's converted
    model, key for key"""
    g, q = converted
    mine, ref = _digest(q.state_dict()), g["converted"]
    assert sorted(mine.keys()) == sorted(ref.keys())
    for k, r in ref.items():
        m = mine[k]
        if isinstance(r, dict):
            assert m["scale"] == r["scale"] and m["zero_point"] == r["zero_point"], k
            assert torch.equal(m["int_repr"], r["int_repr"]), k
        elif isinstance(r, torch.Tensor):
            assert torch.equal(m, r), k
        else:
            assert m == r, k


def test_int8_logits_match_reference_int8_model(converted):
    g, q = converted
    with torch.no_grad():
        logits = q(g["x"])
    assert logits.shape == g["int8_logits"].shape
    assert torch.equal(logits, g["int8_logits"]), float((logits - g["int8_logits"]).abs().max())


def test_convert_requires_prepared_model():
    import frostnet_b200 as F
    with pytest.raises(ValueError):
        F.convert_int8(F.FrostNet(nclass=8, mode="small", width_mult=0.35, quantized=True))


def test_float_forward_uses_functional_modules():
    """the eager (float) forward routes cat / add / quant / dequant through the same modules the int8 model swaps"""
    import frostnet_b200 as F
    torch.manual_seed(0)
    a = F.FrostNet(nclass=8, mode="small", width_mult=0.35, quantized=True, drop_rate=0.0).eval()
    b = F.FrostNet(nclass=8, mode="small", width_mult=0.35, quantized=False, drop_rate=0.0).eval()
    b.load_state_dict(a.state_dict())
    x = torch.randn(2, 3, 64, 64)
    with torch.no_grad():
        assert torch.equal(a(x), b(x))


def test_mobilenetv3_int8_export_matches_reference_convert():
    """convert_int8 on the prepared MobileNetV3-small against the reference's torch.quantization.convert of the real reference
    network (tests/golden/make_golden_int8_mbv3.py): every entry of the converted state_dict (int8 weight bytes by SHA-1,
    scales, zero points, biases) and the int8 logits"""
    import frostnet_b200 as F
    from frostnet_b200 import mobilenetv3 as M
    from util import fill_params_by_name, qdigest_compact
    g = load_golden("int8_mbv3.pt")
    if g["engine"] not in torch.backends.quantized.supported_engines:
        pytest.skip("quantized engine %s not available in this torch build" % g["engine"])
    torch.backends.quantized.engine = g["engine"]
    net = M.get_mobilenet_v3("small", 1.0, nclass=10)
    net.train()
    net.fuse_model()
    F.attach_fake_quant(net)
    fill_params_by_name(net)
    missing, unexpected = net.load_state_dict(g["state"], strict=False)
    assert not unexpected and all(k.endswith(".weight") or k.endswith(".bias") for k in missing)
    q = F.convert_int8(net)
    mine, ref = qdigest_compact(q.state_dict()), g["converted"]
    assert sorted(mine.keys()) == sorted(ref.keys())

    def same(a, b):
        if isinstance(b, torch.Tensor):
            return isinstance(a, torch.Tensor) and torch.equal(a, b)
        if isinstance(b, list):
            return isinstance(a, list) and len(a) == len(b) and all(same(x, y) for x, y in zip(a, b))
        return a == b
    bad = [k for k in ref if not same(mine[k], ref[k])]
    assert not bad, (len(bad), bad[:6])
    with torch.no_grad():
        logits = q(g["x"])
    assert torch.equal(logits, g["int8_logits"]), float((logits - g["int8_logits"]).abs().max())


def test_mobilenetv2_block_int8_export_matches_reference_convert():
    """convert_int8 on a prepared SSDLite backbone block (residual, dilated depthwise; fused convs sitting directly in an
    nn.Sequential) against the reference block's torch.quantization.convert: converted state and int8 output"""
    import frostnet_b200 as F
    from frostnet_b200 import mobilenetv2 as M2
    from util import qdigest_compact
    g = load_golden("int8_mbv2.pt")
    if g["engine"] not in torch.backends.quantized.supported_engines:
        pytest.skip("quantized engine %s not available in this torch build" % g["engine"])
    torch.backends.quantized.engine = g["engine"]
    inp, oup, s, t, H, d = g["case"]
    net = torch.nn.Sequential(F.QuantStub(), M2.InvertedResidual(inp, oup, s, d, t), F.DeQuantStub())
    M2.fuse_model(net)
    F.attach_fake_quant(net)
    missing, unexpected = net.load_state_dict(g["sd"], strict=True)
    assert not missing and not unexpected
    q = F.convert_int8(net)
    mine, ref = qdigest_compact(q.state_dict()), g["converted"]
    assert sorted(mine.keys()) == sorted(ref.keys())

    def same(a, b):
        if isinstance(b, torch.Tensor):
            return isinstance(a, torch.Tensor) and torch.equal(a, b)
        if isinstance(b, list):
            return isinstance(a, list) and len(a) == len(b) and all(same(x, y) for x, y in zip(a, b))
        return a == b
    bad = [k for k in ref if not same(mine[k], ref[k])]
    assert not bad, (len(bad), bad[:6])
    with torch.no_grad():
        out = q(g["x"])
    assert torch.equal(out, g["int8_out"]), float((out - g["int8_out"]).abs().max())
