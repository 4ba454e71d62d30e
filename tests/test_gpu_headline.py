"""Parity on the HEADLINE configuration (BASELINE.json configs[1]: FrostNet-Large-1.0, 224x224, bs=256):

* the whole network - `frostnet_quant_large_1_0` topology at 224x224 - against the CPU oracle, layer by layer from
  identical inputs (teacher forcing; same thresholds as tests/test_gpu_net.py);
* every kernel family at the bs=256 tensor sizes of the largest layers (3.2 M pixels x 96 channels), and at sizes whose
  byte offsets cross 2^31, against exact integer / float64 restatements (computed with plain torch ops on the device:
  the integer ranges make fp32 GEMM / conv exact, see the comments);
* engine contracts the reference's autograd gives for free: stale graphs, copies, odd class counts.
"""
import copy
import ctypes as C

import pytest
import torch
import torch.nn.functional as Fn

from util import rel_l2
from test_gpu_net import (GRAD_REL_L2, LAYER_MISMATCH_RATE, LOGIT_REL_L2, _check_state, _force_dict, _grad_rel_l2)

pytestmark = pytest.mark.gpu

DEV = "cuda:0"


def L():
    from frostnet_b200 import _lib
    return _lib


def stream():
    return torch.cuda.current_stream().cuda_stream


# ------------------------------------------------------------------------------------------ whole network
def test_large_1_0_at_224_teacher_forced_vs_oracle():
    _teacher_forced("large", 1.0, 224, 4, 1000)


# the other topologies / width multipliers of frostnet.py:354-451 (M3): every factory shares these code paths, the channel
# counts (and so the tile shapes, pitches, replication factors, chained-kernel eligibility) differ
@pytest.mark.parametrize("mode,wm,R,N,nclass", [("base", 1.0, 160, 2, 100), ("small", 0.75, 128, 3, 40), ("large", 1.25, 128, 2, 1000),
                                                 ("large", 0.5, 96, 2, 16), ("base", 0.35, 64, 5, 10)])
def test_other_factories_teacher_forced_vs_oracle(mode, wm, R, N, nclass):
    _teacher_forced(mode, wm, R, N, nclass)


def _teacher_forced(mode, wm, R, N, nclass):
    import frostnet_b200 as F
    from oracle import frost_oracle as O
    dev = torch.device(DEV)
    spec = O.net_spec(mode, wm, nclass)
    sd = O.fresh_state_dict(spec, seed=1882)
    gsd = torch.Generator().manual_seed(3)
    for k in sd:                                   # de-trivialise BN (fresh init is gamma=1, beta=0)
        if k.endswith("bn.weight"):
            sd[k] = 0.5 + torch.rand(sd[k].shape, generator=gsd)
        if k.endswith("bn.bias"):
            sd[k] = 0.2 * torch.randn(sd[k].shape, generator=gsd)
    onet = O.OracleNet(spec, sd)
    onet.record = True
    model = F.FrostNet(nclass=nclass, mode=mode, width_mult=wm, quantized=True, drop_rate=0.0)
    model.train()
    model.fuse_model()
    F.prepare_qat(model)
    model.load_state_dict(onet.state_dict(), strict=True)
    model.to(dev)
    g = torch.Generator().manual_seed(1882)
    x = torch.randn(N, 3, R, R, generator=g)
    y = torch.randint(0, nclass, (N,), generator=g)
    ologits = onet.forward(x, training=True, drop_rate=0.0)
    Fn.cross_entropy(ologits, y).backward()
    eng = model._frost_engine
    eng.force, eng.force_report = _force_dict(onet, dev), {}
    sd0 = {k: v.clone() for k, v in model.state_dict().items()}
    logits = model(x.to(dev))
    Fn.cross_entropy(logits, y.to(dev)).backward()
    rep = eng.force_report
    n_cat = sum(1 for b in eng.blocks if b["squeeze"] is not None)
    n_add = sum(1 for b in eng.blocks if b["skip"])
    if (mode, wm) == ("large", 1.0):
        assert (len(eng.layers), n_cat, n_add) == (70, 14, 11)
    assert len(rep) >= len(eng.layers) + n_cat + n_add                  # every conv, cat, add (+ QuantStub) was compared
    assert rep["quant"] == (0.0, 0)
    for ly in eng.layers:                                              # all 5.8 M weight quantize indices
        w_idx = onet.taps[ly.name + ".w_idx"].clamp(-128, 127).to(torch.int8)
        exp = (w_idx.reshape(-1) if ly.layout == 0 else
               w_idx.reshape(ly.cout, -1).t().contiguous().reshape(-1) if ly.layout == 1 else
               w_idx.permute(0, 2, 3, 1).contiguous().reshape(-1))
        assert torch.equal(ly.wq.cpu()[:exp.numel()], exp), ly.name
    worst = max(rep.items(), key=lambda kv: kv[1][0])
    assert all(mx <= 1 for _, mx in rep.values()), {k: v for k, v in rep.items() if v[1] > 1}
    for k, (rate, mx) in rep.items():
        ly = next((l for l in eng.layers if l.name == k), None)
        floor = 2.0 / ly.cout / 16 if ly is not None else 0.0
        assert rate <= max(LAYER_MISMATCH_RATE, floor), (k, rate)
    e_log = rel_l2(logits.detach().cpu(), ologits.detach())
    e_grad, wg = _grad_rel_l2(model, onet)
    print("%s-%g @%d N=%d teacher-forced: worst layer %s rate %.3g; logits rel-L2 %.3g; grad rel-L2 %.3g (worst %s)" % (
        mode, wm, R, N, worst[0], worst[1][0], e_log, e_grad, wg))
    assert e_log < LOGIT_REL_L2 and e_grad < GRAD_REL_L2
    _check_state(model, onet, [k for k in sd0 if "running_" in k or k.endswith("scale") or k.endswith("min_val")
                               or k.endswith("max_val") or k.endswith("zero_point")])


# ------------------------------------------------------------------------------------------ bs=256-scale operators
STAT_FIELDS = 4          # int64 words per FrostChanStats record: sum, sq_lo, sq_hi, (min | max << 32)


def _stats_buf(c):
    t = torch.zeros(c * 32, dtype=torch.uint8, device=DEV)
    L().call("frost_stats_reset", t.data_ptr(), c, stream())
    return t


def _check_stats_dev(buf, I):
    """I: [M, C] int32 on the device."""
    rec = buf.view(torch.int64).reshape(-1, STAT_FIELDS)
    I64 = I.long()
    assert torch.equal(rec[:, 0], I64.sum(0))
    sq = (I64 * I64).sum(0)                                  # < 2^63 for the sizes used here
    assert torch.equal(rec[:, 2] * (1 << 32) + rec[:, 1], sq)
    mn = (rec[:, 3] << 32) >> 32
    mx = rec[:, 3] >> 32
    assert torch.equal(mn, I64.min(0).values) and torch.equal(mx, I64.max(0).values)


@pytest.mark.parametrize("M", [3211264, 6000000])            # 256 x 112 x 112 (the bs=256 layer1.1 expand conv); > 2^31 bytes of int32
def test_pw_conv_forward_at_headline_size(M):
    K, cout = 16, 96
    g = torch.Generator(device=DEV).manual_seed(M % 1000)
    xq = torch.randint(0, 256, (M, K), generator=g, dtype=torch.uint8, device=DEV)
    wq = torch.randint(-128, 128, (cout, K), generator=g, dtype=torch.int8, device=DEV)
    zpa = 3
    # |sum| <= 16 * 255 * 128 < 2^24: the fp32 GEMM is exact whatever its summation order
    assert not torch.backends.cuda.matmul.allow_tf32
    I = ((xq.float() - zpa) @ wq.float().t()).int()
    za, zw = torch.tensor([zpa], dtype=torch.int32, device=DEV), torch.zeros(1, dtype=torch.int32, device=DEV)
    wsum = wq.long().sum(1).int()
    acc = torch.empty(M, cout, dtype=torch.int32, device=DEV)
    st = _stats_buf(cout)
    L().call("frost_pw_conv_forward", xq.data_ptr(), za.data_ptr(), wq.data_ptr(), zw.data_ptr(), wsum.data_ptr(), M, K, cout,
             acc.data_ptr(), st.data_ptr(), stream())
    assert torch.equal(acc, I)
    _check_stats_dev(st, I)


@pytest.mark.parametrize("M", [3211264, 6000000])
def test_bn_passes_at_headline_size(M):
    """bn_finalize + bnq_apply + bn_backward (reduce, apply fp32 and bf16 planes) on [M, 96] against float64 torch."""
    Cc, relu, s_a, s_w = 96, 1, 0.02, 0.003
    g = torch.Generator(device=DEV).manual_seed(5)
    I = (torch.randn(M, Cc, generator=g, device=DEV) * 3000 + torch.randn(Cc, generator=g, device=DEV) * 2000).round().int()
    gamma = 0.5 + torch.rand(Cc, generator=g, device=DEV)
    beta = 0.3 * torch.randn(Cc, generator=g, device=DEV)
    rm, rv = torch.zeros(Cc, device=DEV), torch.ones(Cc, device=DEV)
    sf = (gamma / torch.sqrt(rv + 1e-5)).contiguous()
    I64 = I.long()
    st = torch.zeros(Cc, STAT_FIELDS, dtype=torch.int64, device=DEV)
    sq = (I64 * I64).sum(0)
    st[:, 0], st[:, 1], st[:, 2] = I64.sum(0), sq & 0xffffffff, sq >> 32
    st[:, 3] = (I64.min(0).values & 0xffffffff) | (I64.max(0).values << 32)
    xs, ws = torch.tensor([s_a], device=DEV), torch.tensor([s_w], device=DEV)
    mn, mx = torch.tensor(float("inf"), device=DEV), torch.tensor(float("-inf"), device=DEV)
    scale, zp = torch.ones(1, device=DEV), torch.zeros(1, dtype=torch.int32, device=DEV)
    nbt = torch.zeros((), dtype=torch.int64, device=DEV)
    A, B, meanI, kfac = (torch.zeros(Cc, device=DEV) for _ in range(4))
    mm = torch.zeros(2, device=DEV)
    a = L().BnFinalizeArgs()
    a.stats, a.C, a.count = st.data_ptr(), Cc, M
    a.x_scale, a.w_scale, a.sf = xs.data_ptr(), ws.data_ptr(), sf.data_ptr()
    a.gamma, a.beta, a.running_mean, a.running_var = gamma.data_ptr(), beta.data_ptr(), rm.data_ptr(), rv.data_ptr()
    a.num_batches_tracked = nbt.data_ptr()
    a.momentum, a.eps, a.training, a.relu, a.observe, a.averaging_const = 0.1, 1e-5, 1, relu, 1, 0.01
    a.afq = L().FQ(mn.data_ptr(), mx.data_ptr(), scale.data_ptr(), zp.data_ptr())
    a.A, a.B, a.mean_I, a.kfac, a.cur_minmax = A.data_ptr(), B.data_ptr(), meanI.data_ptr(), kfac.data_ptr(), mm.data_ptr()
    L().call("frost_bn_finalize", C.byref(a), stream())
    q = torch.empty(M, Cc, dtype=torch.uint8, device=DEV)
    L().call("frost_bnq_apply", I.data_ptr(), 0, M, Cc, A.data_ptr(), B.data_ptr(), relu, scale.data_ptr(), zp.data_ptr(),
             q.data_ptr(), Cc, stream())
    # float64 restatement of conv_fused.py:156-167 on conv = s_a*s_w*I
    sasw = float(xs.double() * ws.double())
    u = I.double() * sasw / sf.double()
    mean, var = u.mean(0), u.var(0, unbiased=False)
    v = (u - mean) / torch.sqrt(var + 1e-5) * gamma.double() + beta.double()
    r = torch.relu(v)
    assert abs(float(mx) - float(r.max())) <= 2e-6 * float(r.max()) and float(mn) == 0.0
    s_o, zp_o = float(scale), int(zp)
    idx = torch.round(r.float() * (1.0 / torch.tensor(s_o))) + zp_o
    d = (q.float() - idx.clamp(0, 255)).abs()
    assert float(d.max()) <= 1 and float((d > 0).float().mean()) < 2e-3
    del d, u
    # ---- backward
    dy = torch.randn(M, Cc, generator=g, device=DEV)
    vv = v.detach().requires_grad_(True)
    # dv as the kernels define it (mask from the device's own qparams), then BN backward in float64 by the textbook formula
    maskf = ((idx >= 0) & (idx <= 255) & (v > 0)).double()
    dv = dy.double() * maskf
    xhat = (I.double() * sasw / sf.double() - mean) / torch.sqrt(var + 1e-5)
    S1, T = dv.sum(0), (dv * xhat).sum(0)
    du = gamma.double() / torch.sqrt(var + 1e-5) * (dv - S1 / M - xhat * T / M)
    dz_ref = du / sf.double()
    del xhat, maskf, vv, v, r, idx
    dz = torch.empty(M, Cc, device=DEV)
    sums = torch.zeros(2 * Cc, dtype=torch.float64, device=DEV)
    coef = torch.zeros(3 * Cc, device=DEV)
    dgb, dbeta, dsf = (torch.zeros(Cc, device=DEV) for _ in range(3))
    b = L().BnBackwardArgs()
    b.dy, b.acc, b.M, b.C, b.relu = dy.data_ptr(), I.data_ptr(), M, Cc, relu
    b.A, b.B, b.mean_I, b.kfac = A.data_ptr(), B.data_ptr(), meanI.data_ptr(), kfac.data_ptr()
    b.gamma, b.sf, b.x_scale, b.w_scale = gamma.data_ptr(), sf.data_ptr(), xs.data_ptr(), ws.data_ptr()
    b.out_scale, b.out_zp, b.eps = scale.data_ptr(), zp.data_ptr(), 1e-5
    b.sums, b.coef, b.dz = sums.data_ptr(), coef.data_ptr(), dz.data_ptr()
    b.dgamma_bn, b.dbeta, b.dsf_bn = dgb.data_ptr(), dbeta.data_ptr(), dsf.data_ptr()
    L().call("frost_bn_backward", C.byref(b), stream())
    ref_scale = float(dz_ref.abs().max())
    # the few elements whose mask differs between the fp32 device formula and float64 (v within an ulp of 0 or of a
    # clamp boundary) are excluded by a robust statistic: 99.99 % of the elements agree to 2e-4 of the scale
    err = (dz.double() - dz_ref).abs()
    assert float((err > 2e-4 * ref_scale).float().mean()) < 1e-4, float(err.max())
    torch.testing.assert_close(dbeta.double(), S1, rtol=1e-4, atol=1e-4 * float(S1.abs().max()))
    torch.testing.assert_close(dgb.double(), T, rtol=1e-3, atol=2e-4 * float(T.abs().max()))
    p_hi = torch.empty(M, Cc, dtype=torch.bfloat16, device=DEV)
    p_lo = torch.empty(M, Cc, dtype=torch.bfloat16, device=DEV)
    b.dz, b.dz_lo, b.dz_format = p_hi.data_ptr(), p_lo.data_ptr(), 1
    L().call("frost_bn_backward", C.byref(b), stream())
    assert torch.equal(p_hi, dz.to(torch.bfloat16))
    assert float(((p_hi.float() + p_lo.float()) - dz).abs().max()) <= 2.0 ** -16 * float(dz.abs().max())


@pytest.mark.parametrize("N", [256, 448])                   # 448 x 112 x 112 x 96 x 4 B > 2^31
def test_depthwise_family_at_headline_size(N):
    """dw forward (exact), dgrad and wgrad (fp32) on N x 112 x 112 x 96, k=3 s=1 - the shape of the largest depthwise
    tensor class at bs=256 - against torch's native (non-cuDNN) depthwise convolution."""
    H = W = 112
    Cc, k, s = 96, 3, 1
    g = torch.Generator(device=DEV).manual_seed(N)
    xq = torch.randint(0, 256, (N, H, W, Cc), generator=g, dtype=torch.uint8, device=DEV)
    w = torch.randint(-128, 128, (Cc, 1, k, k), generator=g, dtype=torch.int8, device=DEV)
    zpa = 7
    wd = w.reshape(Cc, k * k).t().contiguous()
    za, zw = torch.tensor([zpa], dtype=torch.int32, device=DEV), torch.zeros(1, dtype=torch.int32, device=DEV)
    acc = torch.empty(N, H, W, Cc, dtype=torch.int32, device=DEV)
    st = _stats_buf(Cc)
    L().call("frost_dw_conv_forward", xq.data_ptr(), Cc, za.data_ptr(), wd.data_ptr(), zw.data_ptr(), N, H, W, Cc, k, s,
             acc.data_ptr(), st.data_ptr(), stream())
    with torch.backends.cudnn.flags(enabled=False):
        # |sum| <= 9 * 255 * 128 < 2^24: exact in fp32
        xf = (xq.float() - zpa).permute(0, 3, 1, 2)
        ref = Fn.conv2d(xf, w.float(), None, s, 1, 1, Cc).permute(0, 2, 3, 1)
        assert torch.equal(acc, ref.int())
        _check_stats_dev(st, acc.reshape(-1, Cc))
        del ref, acc
        # backward: dz fp32
        dz = torch.randn(N, H, W, Cc, generator=g, device=DEV)
        sw, sa = torch.tensor([0.004], device=DEV), torch.tensor([0.03], device=DEV)
        dx = torch.empty(N, H, W, Cc, device=DEV)
        L().call("frost_dw_dgrad", dz.data_ptr(), wd.data_ptr(), sw.data_ptr(), zw.data_ptr(), N, H, W, Cc, k, s, dx.data_ptr(), 0,
                 stream())
        dz_nchw = dz.permute(0, 3, 1, 2)
        dx_ref = Fn.conv_transpose2d(dz_nchw, w.float() * sw, None, s, 1, 0, Cc).permute(0, 2, 3, 1)
        assert float((dx - dx_ref).abs().max()) <= 1e-5 * float(dx_ref.abs().max())
        del dx, dx_ref
        dwq = torch.empty(k * k, Cc, device=DEV)
        L().call("frost_dw_wgrad", dz.data_ptr(), xq.data_ptr(), Cc, sa.data_ptr(), za.data_ptr(), N, H, W, Cc, k, s, dwq.data_ptr(),
                 stream())
        # float64 reference of the weight gradient, image chunks at a time
        ref = torch.zeros(Cc, 1, k, k, dtype=torch.float64, device=DEV)
        for n0 in range(0, N, 32):
            xs_ = xf[n0:n0 + 32].double().requires_grad_(False)
            wv = torch.zeros(Cc, 1, k, k, dtype=torch.float64, device=DEV, requires_grad=True)
            out = Fn.conv2d(xs_, wv, None, s, 1, 1, Cc)
            out.backward(dz_nchw[n0:n0 + 32].double())
            ref += wv.grad
        ref = (ref * float(sa)).reshape(Cc, k * k).t()
        assert float((dwq.double() - ref).abs().max()) <= 2e-4 * float(ref.abs().max())


@pytest.mark.parametrize("M", [3211264])
def test_pw_backward_tc_at_headline_size(M):
    """Tensor-core dgrad / wgrad of the bs=256 layer1.1 expand conv (dz planes of 2 x 616 MB)."""
    K, cout = 16, 96
    g = torch.Generator(device=DEV).manual_seed(9)
    dz = torch.randn(M, cout, generator=g, device=DEV)
    hi = dz.to(torch.bfloat16)
    lo = (dz - hi.float()).to(torch.bfloat16)
    xq = torch.randint(0, 256, (M, K), generator=g, dtype=torch.uint8, device=DEV)
    wq = torch.randint(-128, 128, (cout, K), generator=g, dtype=torch.int8, device=DEV)
    wt = wq.float().t().contiguous().to(torch.bfloat16)
    sw, sa = torch.tensor([0.004], device=DEV), torch.tensor([0.03], device=DEV)
    za = torch.tensor([5], dtype=torch.int32, device=DEV)
    dx = torch.empty(M, K, device=DEV)
    L().call("frost_pw_dgrad_tc", hi.data_ptr(), lo.data_ptr(), wt.data_ptr(), sw.data_ptr(), M, K, cout, dx.data_ptr(), 0, stream())
    dx_ref = (dz.double() @ wq.double()) * float(sw)
    assert float((dx.double() - dx_ref).abs().max()) <= 1e-4 * float(dx_ref.abs().max())
    dwq = torch.empty(cout, K, device=DEV)
    L().call("frost_pw_wgrad_tc", hi.data_ptr(), lo.data_ptr(), xq.data_ptr(), K, sa.data_ptr(), za.data_ptr(), M, K, cout,
             dwq.data_ptr(), stream())
    ref = (dz.double().t() @ (xq.double() - 5.0)) * float(sa)
    assert float((dwq.double() - ref).abs().max()) <= 2e-4 * float(ref.abs().max())


# ------------------------------------------------------------------------------------------ engine contracts
def _small_model(nclass=16, dev=DEV):
    import frostnet_b200 as F
    torch.manual_seed(0)
    m = F.FrostNet(nclass=nclass, mode="small", width_mult=0.35, quantized=True, drop_rate=0.0)
    m.train()
    m.fuse_model()
    F.prepare_qat(m)
    return m.to(dev)


def test_backward_through_a_stale_graph_raises():
    m = _small_model()
    x = torch.randn(2, 3, 64, 64, device=DEV)
    out1 = m(x)
    with torch.no_grad():
        m(x)                                      # e.g. a validation / EMA forward between forward and backward
    with pytest.raises(RuntimeError, match="no longer the latest"):
        out1.sum().backward()
    out2 = m(x)
    out2.sum().backward()                         # the latest graph is fine
    assert all(p.grad is not None for p in m.parameters())


def test_deepcopy_gets_its_own_engine_and_leaves_the_original_alone():
    m = _small_model()
    x = torch.randn(2, 3, 64, 64, device=DEV)
    m(x).sum().backward()
    ema = copy.deepcopy(m)                        # timm ModelEma / mmdet EMA hooks do exactly this
    assert ema._frost_engine is not m._frost_engine and ema._frost_engine.model is ema
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    ema.eval()
    with torch.no_grad():
        a = ema(x)
    for k, v in m.state_dict().items():           # the copy's forward did not write into the original's buffers
        assert torch.equal(v, sd[k]), k
    m.eval()
    with torch.no_grad():
        b = m(x)
    assert torch.equal(a, b)                      # same weights, same state -> same logits
    # moving a parameter without Module._apply: the engine notices and rebuilds its pointer tables
    m.train()
    p = m.conv1.conv[0].weight
    p.data = p.data.clone()
    m(x).sum().backward()
    assert torch.isfinite(p.grad).all()


def test_class_count_not_a_multiple_of_4_trains():
    m = _small_model(nclass=10)                   # CIFAR-10 head (reference: Classification/train.py num_classes)
    x = torch.randn(4, 3, 64, 64, device=DEV)
    y = torch.randint(0, 10, (4,), device=DEV)
    logits = m(x)
    assert logits.shape == (4, 10)
    Fn.cross_entropy(logits, y).backward()
    cls = m.classifier[2]
    assert cls.weight.grad.shape == cls.weight.shape and torch.isfinite(cls.weight.grad).all()
    assert float(cls.weight.grad.abs().sum()) > 0 and float(cls.bias.grad.abs().sum()) > 0
    # against the 12-class run restricted to... no: against float64 autograd of the head alone
    eng = m._frost_engine
    eng.record_taps = True
    m.zero_grad()
    logits = m(x)
    pooled, pre = eng.last_taps["pooled"], eng.last_taps["classifier.2.pre"]
    eng.record_taps = False
    dlog = torch.randn_like(logits)
    logits.backward(dlog)
    wq = (eng.cls.wq[:10 * 1280].reshape(10, 1280).double() - float(cls.weight_fake_quant.zero_point)) * float(cls.weight_fake_quant.scale)
    torch.testing.assert_close(pre.double(), pooled.double() @ wq.t() + cls.bias.double(), rtol=1e-4, atol=1e-5)
    s, zp = float(cls.activation_post_process.scale), int(cls.activation_post_process.zero_point)
    idx = torch.round(pre * (1.0 / torch.tensor(s, device=DEV))) + zp
    dpre = dlog * ((idx >= 0) & (idx <= 255))
    torch.testing.assert_close(cls.bias.grad.double(), dpre.double().sum(0), rtol=1e-4, atol=1e-6)


def test_requires_grad_input_is_rejected():
    m = _small_model()
    x = torch.randn(2, 3, 64, 64, device=DEV, requires_grad=True)
    with pytest.raises(RuntimeError, match="requires grad"):
        m(x)
