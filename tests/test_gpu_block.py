"""Per-module executor (frostnet_b200/block_engine.py): a prepared CascadePreExBottleneck / QuantStub called on its own
must compute exactly what it computes inside the whole-network engine - same kernels, fp32 NCHW at the boundary."""
import copy

import pytest
import torch

from util import build_model_from_golden, load_golden

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _nchw(q_rows, N, H, W, C, scale, zp):
    """uint8 NHWC rows [M][C] -> dequantised fp32 NCHW (what a stand-alone producer would hand over)"""
    q = q_rows.view(N, H, W, C).permute(0, 3, 1, 2).float()
    return ((q - zp.float()) * scale).contiguous()


def _named_block(model, name):
    m = model
    for part in name.split("."):
        m = m[int(part)] if part.isdigit() else getattr(m, part)
    return m


# (block, the tap that holds its input, the module whose observer quantised that input)
CASES = [("layer1.0", "conv1.conv.0.out_q", "conv1.conv.0"),                 # e = 1: depthwise + reduce only
         ("layer1.1", "layer1.0.reduce_conv.conv.0.out_q", "layer1.0.reduce_conv.conv.0"),    # MB, stride 2
         ("layer3.1", None, None)]                                           # CAS with skip: input found below


@pytest.mark.parametrize("case", range(len(CASES)))
def test_standalone_block_matches_whole_network_engine(case):
    import frostnet_b200 as F
    g = load_golden("net_small035.pt")
    whole = build_model_from_golden(g, DEV)
    whole.train()
    alone = copy.deepcopy(whole)                      # same pre-step state: observers / BN statistics evolve identically
    eng = whole.__dict__["_frost_engine"]
    eng.record_taps = True
    x = g["xs"][0].to(DEV)
    logits = whole(x)
    loss = torch.nn.functional.cross_entropy(logits, g["ys"][0].to(DEV))
    loss.backward()
    taps, gtaps = eng.last_taps, eng.last_grad_taps
    name, in_tap, in_mod = CASES[case]
    blocks = [b for b in eng.blocks]
    bi = [b["name"] for b in blocks].index(name)
    b = blocks[bi]
    if in_tap is None:                                 # input = output of the previous block (its add or its reduce conv)
        prev = blocks[bi - 1]
        if prev["skip"]:
            in_tap, fqmod = prev["name"] + ".add_q", _named_block(whole, prev["name"]).skip_add.activation_post_process
        else:
            in_tap = prev["reduce"].name + ".out_q"
            fqmod = prev["reduce"].mod.activation_post_process
    else:
        fqmod = _named_block(whole, in_mod).activation_post_process
    blk_w, blk_a = _named_block(whole, name), _named_block(alone, name)
    q_in = taps[in_tap]
    # spatial size of the block input: from the saved quantised tensor of the block's first conv
    first = b["squeeze"] or b["conv1"] or b["conv2"]
    N = x.shape[0]
    C = blk_w.in_channels
    HW = q_in.shape[0] // N
    H = W = int(round(HW ** 0.5))
    assert q_in.shape == (N * H * W, C)
    xin = _nchw(q_in, N, H, W, C, fqmod.scale, fqmod.zero_point).requires_grad_(True)
    xin._frost_qparams = (fqmod.scale, fqmod.zero_point)
    blk_a.train()
    y = blk_a(xin)
    assert hasattr(y, "_frost_qparams")
    # ---- forward: the block output, index for index
    out_tap = name + ".add_q" if b["skip"] else b["reduce"].name + ".out_q"
    out_fq = blk_w.skip_add.activation_post_process if b["skip"] else b["reduce"].mod.activation_post_process
    Co = blk_w.out_channels
    Ho = y.shape[2]
    y_ref = _nchw(taps[out_tap], N, Ho, Ho, Co, out_fq.scale, out_fq.zero_point)
    assert y.shape == y_ref.shape
    assert torch.equal(y.detach(), y_ref), float((y.detach() - y_ref).abs().max())
    # every observer / BN buffer of the block moved exactly as inside the whole network
    sw, sa = blk_w.state_dict(), blk_a.state_dict()
    for k in sw:
        if not (k.endswith("weight") or k.endswith("bias")) or ".bn." in k:
            assert torch.equal(sw[k], sa[k]), k
    # ---- backward from the same upstream gradient
    dy = gtaps[name + ".out"].view(N, Ho, Ho, Co).permute(0, 3, 1, 2).contiguous()
    y.backward(dy)
    dx_ref = gtaps[name + ".in"].view(N, H, W, C).permute(0, 3, 1, 2)
    scale = float(dx_ref.abs().max())
    assert float((xin.grad - dx_ref).abs().max()) <= 1e-5 * scale, (float((xin.grad - dx_ref).abs().max()), scale)
    for (k, pw), (_, pa) in zip(blk_w.named_parameters(), blk_a.named_parameters()):
        assert pa.grad is not None, k
        tol = 1e-5 * float(pw.grad.abs().max()) + 1e-12
        assert float((pa.grad - pw.grad).abs().max()) <= tol, (k, float((pa.grad - pw.grad).abs().max()), tol)


def test_standalone_conv_block_matches_whole_network_engine():
    """last_layer (ConvBNReLU 1x1) and a depthwise ConvBNReLU called on their own"""
    g = load_golden("net_small035.pt")
    whole = build_model_from_golden(g, DEV)
    whole.train()
    alone = copy.deepcopy(whole)
    eng = whole.__dict__["_frost_engine"]
    eng.record_taps = True
    x = g["xs"][0].to(DEV)
    whole(x).sum().backward()
    taps = eng.last_taps
    N = x.shape[0]
    last_blk = eng.blocks[-1]
    for conv_name, in_tap, in_fq in (
            ("last_layer", (last_blk["name"] + ".add_q") if last_blk["skip"] else last_blk["reduce"].name + ".out_q",
             (_named_block(whole, last_blk["name"]).skip_add if last_blk["skip"] else last_blk["reduce"].mod).activation_post_process),
            ("layer1.0.conv2", "conv1.conv.0.out_q", whole.conv1.conv[0].activation_post_process)):
        cw, ca = _named_block(whole, conv_name), _named_block(alone, conv_name)
        q_in = taps[in_tap]
        C = cw.conv[0].in_channels
        H = int(round((q_in.shape[0] // N) ** 0.5))
        xin = _nchw(q_in, N, H, H, C, in_fq.scale, in_fq.zero_point)
        xin._frost_qparams = (in_fq.scale, in_fq.zero_point)
        ca.train()
        y = ca(xin)
        ofq = cw.conv[0].activation_post_process
        y_ref = _nchw(taps[conv_name + ".conv.0.out_q"], N, y.shape[2], y.shape[3], y.shape[1], ofq.scale, ofq.zero_point)
        assert torch.equal(y.detach(), y_ref), conv_name
        y.sum().backward()
        assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in ca.parameters())
    # the dense 3x3 stem runs on its own too (direct-convolution kernels; no input gradient, like the reference's stem)
    xs = torch.zeros(N, 3, 64, 64, device=DEV)
    xs._frost_qparams = (whole.quant.activation_post_process.scale, whole.quant.activation_post_process.zero_point)
    ys = alone.conv1(xs)
    assert ys.shape[:2] == (N, alone.conv1.conv[0].out_channels) and hasattr(ys, "_frost_qparams")
    ys.sum().backward()
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in alone.conv1.parameters())
    with pytest.raises(RuntimeError, match="no input gradient"):
        xg = torch.zeros(N, 3, 64, 64, device=DEV, requires_grad=True)
        xg._frost_qparams = xs._frost_qparams
        alone.conv1(xg)
    import frostnet_b200 as F
    big = F.ConvBN(16, 16, 3, 1, 1)                      # a dense 3x3 on 16 channels has no kernel
    big.fuse_model()
    F.attach_fake_quant(big)
    big.to(DEV).train()
    with pytest.raises(RuntimeError, match="stem-sized"):
        xb = torch.zeros(N, 16, 8, 8, device=DEV)
        xb._frost_qparams = xs._frost_qparams
        big(xb)


def test_quant_stub_then_blocks_chain_standalone():
    """QuantStub -> (foreign op) is rejected; QuantStub -> block -> block chains through `_frost_qparams`."""
    import frostnet_b200 as F
    torch.manual_seed(3)
    stub = F.qat.QuantStub()
    b1 = F.CascadePreExBottleneck(16, 24, quantized=True, kernel_size=3, stride=2, expand_ratio=6, reduce_factor=4)
    b2 = F.CascadePreExBottleneck(24, 24, quantized=True, kernel_size=5, stride=1, expand_ratio=3, reduce_factor=2)
    for m in (b1, b2):
        for c in m.modules():
            if isinstance(c, (F.ConvBNReLU, F.ConvBN)):
                c.fuse_model()
    net = torch.nn.Sequential(stub, b1, b2).to(DEV).train()
    for m in list(net.modules()):
        if isinstance(m, F.qat.FrostConvBn2d):
            m.weight_fake_quant = F.FrostFakeQuantize.weight().to(DEV)
            m.activation_post_process = F.FrostFakeQuantize.act().to(DEV)
        elif isinstance(m, (F.qat.FloatFunctional, F.qat.QuantStub)):
            m.activation_post_process = F.FrostFakeQuantize.act().to(DEV)
    x = torch.randn(2, 16, 32, 32, device=DEV)
    y = net(x)
    assert y.shape == (2, 24, 16, 16) and hasattr(y, "_frost_qparams")
    y.square().mean().backward()
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in net.parameters())
    # the output is on its grid: quantise-dequantise with its own qparams is the identity
    s, z = y._frost_qparams
    q = torch.clamp(torch.round(y.detach() / s) + z, 0, 255)
    assert float(((q - z) * s - y.detach()).abs().max()) <= 1e-6 * float(y.detach().abs().max())
    with pytest.raises(RuntimeError, match="without quantisation parameters"):
        b2(torch.relu(y.detach()))                      # a foreign op dropped the qparams


@pytest.mark.parametrize("op", ["add", "cat"])
def test_standalone_float_functional_matches_torch_ao(op):
    """FloatFunctional.add / .cat called on their own vs torch.ao's QAT FloatFunctional (the arithmetic the reference runs,
    functional_modules.py:50-52,80-82) on the CPU: values, observer state and input gradients over two steps."""
    import frostnet_b200 as F
    import torch.ao.quantization as taq
    qc = taq.get_default_qat_qconfig("qnnpack")
    # reference (CPU)
    r1, r2 = taq.QuantStub(), taq.QuantStub()
    r1.activation_post_process, r2.activation_post_process = qc.activation(), qc.activation()
    rff = torch.ao.nn.quantized.FloatFunctional()
    rff.activation_post_process = qc.activation()
    rq = lambda stub, t: stub.activation_post_process(stub(t))
    # ours (device)
    m = torch.nn.ModuleDict(dict(s1=F.QuantStub(), s2=F.QuantStub(), ff=F.qat.FloatFunctional()))
    F.attach_fake_quant(m)
    m.to(DEV).train()
    g = torch.Generator().manual_seed(11)
    for step in range(2):
        Cb = 24 if op == "add" else 40
        x = (torch.randn(3, 24, 10, 9, generator=g) * (1.5 + step)).requires_grad_(True)
        y = (torch.randn(3, Cb, 10, 9, generator=g) * 0.7 + 0.3).requires_grad_(True)
        ref = rff.add(rq(r1, x), rq(r2, y)) if op == "add" else rff.cat([rq(r1, x), rq(r2, y)], 1)
        w = torch.randn(ref.shape, generator=g)
        (ref * w).sum().backward()
        xd, yd = x.detach().to(DEV).requires_grad_(True), y.detach().to(DEV).requires_grad_(True)
        a, b = m["s1"](xd), m["s2"](yd)
        out = m["ff"].add(a, b) if op == "add" else m["ff"].cat([a, b], 1)
        assert hasattr(out, "_frost_qparams")
        assert torch.equal(out.detach().cpu(), ref.detach()), (step, float((out.detach().cpu() - ref.detach()).abs().max()))
        (out * w.to(DEV)).sum().backward()
        for ours, theirs in ((xd.grad, x.grad), (yd.grad, y.grad)):
            assert float((ours.cpu() - theirs).abs().max()) <= 1e-6 * float(theirs.abs().max()) + 1e-12
        for mine, theirs in ((m["s1"], r1), (m["s2"], r2), (m["ff"], rff)):
            fm, ft = mine.activation_post_process, theirs.activation_post_process
            assert float(fm.scale) == float(ft.scale) and int(fm.zero_point) == int(ft.zero_point)
            assert float(fm.activation_post_process.min_val) == float(ft.activation_post_process.min_val)
            assert float(fm.activation_post_process.max_val) == float(ft.activation_post_process.max_val)


def test_mobilenetv3_style_block_from_standalone_modules_trains():
    """An inverted-residual block in the reference's MobileNetV3 style (mobilenetv3.py:113-160: 1x1 ConvBN + h-swish, depthwise
    ConvBN + h-swish, 1x1 ConvBN, residual add) wired from frostnet_b200 modules and run through the per-module executor:
    every boundary hands its quantisation grid on, gradients reach every parameter, and a few SGD steps reduce the loss."""
    import frostnet_b200 as F

    class Block(torch.nn.Module):
        def __init__(self, c, e):
            super().__init__()
            self.quant = F.QuantStub()
            self.pw = F.ConvBN(c, e, 1)
            self.act1 = F.Hswish()
            self.dw = F.ConvBN(e, e, 3, 1, 1, 1, groups=e)
            self.act2 = F.Hswish()
            self.pwl = F.ConvBN(e, c, 1)
            self.skip_add = F.qat.FloatFunctional()

        def forward(self, x):
            x = self.quant(x)
            y = self.act1(self.pw(x))
            y = self.act2(self.dw(y))
            return self.skip_add.add(x, self.pwl(y))

    torch.manual_seed(5)
    blk = Block(16, 64)
    for m in blk.modules():
        if isinstance(m, F.ConvBN):
            m.fuse_model()
    F.attach_fake_quant(blk)
    blk.to(DEV).train()
    opt = torch.optim.SGD(blk.parameters(), lr=0.05)
    x = torch.randn(8, 16, 14, 14, device=DEV)
    target = torch.randn(8, 16, 14, 14, device=DEV) * 0.1
    losses = []
    for _ in range(6):
        opt.zero_grad()
        out = blk(x)
        assert hasattr(out, "_frost_qparams") and out.shape == x.shape
        loss = (out - target).square().mean()
        loss.backward()
        assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in blk.parameters())
        opt.step()
        losses.append(float(loss.detach()))
    assert losses[-1] < losses[0], losses
