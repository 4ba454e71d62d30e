"""Per-entry-point parity of libfrost_b200.so (called through the C ABI via ctypes) against CPU
restatements: the oracle's fake-quant for everything that quantises (bit-exact), exact integer
arithmetic for the convolutions (bit-exact), float64 torch autograd for the backward kernels."""
import ctypes as C

import numpy as np
import pytest
import torch
import torch.nn.functional as Fn

pytestmark = pytest.mark.gpu

DEV = "cuda:0"


def L():
    from frostnet_b200 import _lib
    return _lib


def O():
    from oracle import frost_oracle
    return frost_oracle


class DevFQ:
    """Device buffers of one fake-quant module + the matching oracle FQ."""

    def __init__(self, qmin, qmax, sym):
        self.min_val = torch.tensor(float("inf"), device=DEV)
        self.max_val = torch.tensor(float("-inf"), device=DEV)
        self.scale = torch.ones(1, device=DEV)
        self.zp = torch.zeros(1, dtype=torch.int32, device=DEV)
        self.oracle = O().FQ(qmin, qmax, sym)
        self.qmin, self.qmax, self.sym = qmin, qmax, sym

    def c(self):
        return L().FQ(self.min_val.data_ptr(), self.max_val.data_ptr(), self.scale.data_ptr(), self.zp.data_ptr())

    def set(self, scale, zp):
        self.scale.fill_(scale)
        self.zp.fill_(zp)
        self.oracle.scale, self.oracle.zero_point = np.float32(scale), int(zp)

    def assert_state_equal(self):
        o = self.oracle
        assert float(self.min_val) == float(o.min_val) and float(self.max_val) == float(o.max_val)
        assert float(self.scale) == float(o.scale), (float(self.scale), float(o.scale))
        assert int(self.zp) == int(o.zero_point)


def scratch():
    return torch.zeros(L().FQ_SCRATCH_FLOATS, device=DEV)


def stream():
    return torch.cuda.current_stream().cuda_stream


# ------------------------------------------------------------------------------------------ FQ
@pytest.mark.parametrize("sym", [False, True])
def test_fq_forward_bit_exact_vs_oracle(sym):
    rng = np.random.RandomState(1)
    qmin, qmax = (-128, 127) if sym else (0, 255)
    for case in range(24):
        n = [1, 3, 4, 5, 1000, 4099, 300001][case % 7]
        x = torch.from_numpy(rng.randn(n).astype(np.float32)) * float(10 ** rng.uniform(-5, 3))
        if case % 4 == 1:
            x = x.abs()
        if case % 4 == 2:
            x = -x.abs()
        fq = DevFQ(qmin, qmax, sym)
        sc = scratch()
        for it in range(3):
            xi = (x * (1 + 0.7 * it)).contiguous()
            xd = xi.to(DEV)
            y = torch.empty_like(xd)
            mask = torch.empty(n, dtype=torch.uint8, device=DEV)
            q = torch.empty(n, dtype=torch.int32, device=DEV)
            L().call("frost_fq_forward", xd.data_ptr(), n, fq.c(), qmin, qmax, int(sym), 1, 0.01, y.data_ptr(),
                     mask.data_ptr(), q.data_ptr(), sc.data_ptr(), stream())
            xo = xi.clone().requires_grad_(True)
            yo = fq.oracle(xo)
            yo.sum().backward()
            fq.assert_state_equal()
            assert torch.equal(y.cpu(), yo.detach())
            assert torch.equal(mask.cpu().float(), xo.grad)
            assert torch.equal(q.cpu().float(), fq.oracle.last_idx.clamp(qmin, qmax))
            dy = torch.randn(n, device=DEV)
            dx = torch.empty_like(dy)
            L().call("frost_fq_backward", dy.data_ptr(), mask.data_ptr(), n, dx.data_ptr(), stream())
            assert torch.equal(dx, dy * mask)


def test_fq_ties_and_mask_known_answers():
    fq = DevFQ(-128, 127, True)
    fq.set(1.0, 0)
    x = torch.tensor([0.5, 1.5, 2.5, -0.5, -1.5, 127.4999, 127.5, -128.5, -128.51], device=DEV)
    y, mask = torch.empty_like(x), torch.empty(9, dtype=torch.uint8, device=DEV)
    L().call("frost_fq_forward", x.data_ptr(), 9, fq.c(), -128, 127, 1, 0, 0.01, y.data_ptr(), mask.data_ptr(), None,
             scratch().data_ptr(), stream())
    assert y.tolist() == [0.0, 2.0, 2.0, -0.0, -2.0, 127.0, 127.0, -128.0, -128.0]
    assert mask.tolist() == [1, 1, 1, 1, 1, 1, 0, 1, 0]


@pytest.mark.parametrize("H,W", [(17, 13), (16, 12), (224, 224)])      # odd plane: per-pixel kernel; HW % 4 == 0: rgb4 kernel
def test_input_quant_matches_oracle(H, W):
    torch.manual_seed(0)
    N, Cc = 3, 3
    fq = DevFQ(0, 255, False)
    sc = scratch()
    for it in range(2):
        x = torch.randn(N, Cc, H, W) * (1 + it)
        xd = x.to(DEV)
        q = torch.empty(N, H, W, Cc, dtype=torch.uint8, device=DEV)
        mm = torch.empty(2, device=DEV)
        L().call("frost_input_quant", xd.data_ptr(), N, Cc, H, W, fq.c(), 1, 0.01, q.data_ptr(), mm.data_ptr(),
                 sc.data_ptr(), stream())
        yo = fq.oracle(x)
        fq.assert_state_equal()
        exp = fq.oracle.last_idx.clamp(0, 255).to(torch.uint8).permute(0, 2, 3, 1)
        assert torch.equal(q.cpu(), exp)
        assert float(mm[0]) == float(yo.min()) and float(mm[1]) == float(yo.max())


# ------------------------------------------------------------------------------------------ weights
def _weight_desc(w, gamma, var, eps, layout, fq, bufs):
    d = L().WeightDesc()
    d.weight = w.data_ptr()
    d.bn_weight = gamma.data_ptr() if gamma is not None else None
    d.bn_var = var.data_ptr() if var is not None else None
    d.bn_eps = eps
    d.cout, d.cin_g, d.kh, d.kw = w.shape
    d.layout, d.observe, d.averaging_const = layout, 1, 0.01
    d.wfq = fq.c()
    for k in ("wq", "wt_bf16", "wmask", "sf", "rstd_run", "wsum", "dwq", "dgamma_bn", "dsf_bn", "dweight", "dgamma"):
        setattr(d, k, bufs[k].data_ptr() if bufs.get(k) is not None else None)
    return d


def _alloc_wbufs(w, with_bn=True):
    n, co = w.numel(), w.shape[0]
    b = dict(wq=torch.zeros(n, dtype=torch.int8, device=DEV),
             wt_bf16=torch.zeros(n, dtype=torch.bfloat16, device=DEV) if (w.shape[2] == 1 and w.shape[1] > 1) else None, wmask=torch.zeros(n, dtype=torch.uint8, device=DEV),
             sf=torch.zeros(co, device=DEV), rstd_run=torch.zeros(co, device=DEV),
             wsum=torch.zeros(co, dtype=torch.int32, device=DEV), dwq=torch.zeros(n, device=DEV),
             dweight=torch.zeros(n, device=DEV))
    if with_bn:
        b.update(dgamma_bn=torch.zeros(co, device=DEV), dsf_bn=torch.zeros(co, device=DEV), dgamma=torch.zeros(co, device=DEV))
    return b


def _to_layout(t, layout):
    co = t.shape[0]
    if layout == 0:
        return t.reshape(-1)
    if layout == 1:
        return t.reshape(co, -1).t().contiguous().reshape(-1)
    return t.permute(0, 2, 3, 1).contiguous().reshape(-1)


@pytest.mark.parametrize("shape,layout", [((24, 16, 1, 1), 0), ((40, 1, 5, 5), 1), ((16, 3, 3, 3), 2), ((1000, 64, 1, 1), 0)])
def test_weight_prep_and_backward_vs_oracle(shape, layout):
    torch.manual_seed(3)
    descs, items = [], []
    for variant in range(3):
        w = torch.randn(shape) * 0.2
        if variant == 1:
            w = w.abs() + 0.01                    # one-signed: symmetric degrades to affine (K5), zp = -128
        with_bn = variant != 2
        gamma = (0.5 + torch.rand(shape[0])) if with_bn else None
        var = (0.5 + torch.rand(shape[0])) if with_bn else None
        fq = DevFQ(-128, 127, True)
        wd = w.to(DEV)
        gd, vd = (gamma.to(DEV), var.to(DEV)) if with_bn else (None, None)
        bufs = _alloc_wbufs(wd, with_bn)
        bufs["dwq"].normal_()
        if with_bn:
            bufs["dgamma_bn"].normal_()
            bufs["dsf_bn"].normal_()
        descs.append(_weight_desc(wd, gd, vd, 1e-5, layout, fq, bufs))
        items.append((w, gamma, var, fq, bufs, wd, gd, vd))
    arr = (L().WeightDesc * len(descs))(*descs)
    tab = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).to(DEV)
    n_el = int(np.prod(shape))
    chunks, n_chunks = L().chunk_table([n_el] * len(descs), L().WEIGHT_CHUNK, DEV)
    bchunks, n_bchunks = L().chunk_table([shape[0]] * len(descs), L().WEIGHT_BWD_CHANNELS, DEV)
    wscratch = torch.tensor([float("inf"), float("-inf")] * len(descs), device=DEV)
    for it in range(2):
        L().call("frost_weight_prep_multi", tab.data_ptr(), len(descs), chunks.data_ptr(), n_chunks, wscratch.data_ptr(), stream())
        L().call("frost_weight_backward_multi", tab.data_ptr(), len(descs), bchunks.data_ptr(), n_bchunks, stream())
        for (w, gamma, var, fq, bufs, *_rest) in items:
            wl = w.clone().requires_grad_(True)
            if gamma is not None:
                gl = gamma.clone().requires_grad_(True)
                sf = gl / torch.sqrt(var + 1e-5)
                ws = wl * sf.reshape(-1, 1, 1, 1)
            else:
                gl, sf, ws = None, torch.ones(shape[0]), wl
            wqf = fq.oracle(ws)
            fq.assert_state_equal()
            idx = fq.oracle.last_idx
            assert torch.equal(bufs["wq"].cpu(), _to_layout(idx.clamp(-128, 127).to(torch.int8), layout))
            if bufs["wt_bf16"] is not None:
                exp_t = (bufs["wq"].cpu().reshape(shape[0], shape[1]).float() - float(int(fq.zp))).t()
                assert torch.equal(bufs["wt_bf16"].cpu().float().reshape(shape[1], shape[0]), exp_t)
            # ATen's vectorised CPU sqrt/div may differ from IEEE scalar code in the last ulp of a vector tail
            torch.testing.assert_close(bufs["sf"].cpu(), sf.detach(), rtol=2.5e-7, atol=0)
            assert torch.equal(bufs["wsum"].cpu(), idx.clamp(-128, 127).reshape(shape[0], -1).sum(1).int())
            # backward: dL/dWq given (in kernel layout)
            dwq_pt = torch.empty(shape)
            perm = _to_layout(torch.arange(w.numel()).reshape(shape), layout)
            dwq_pt.reshape(-1)[perm] = bufs["dwq"].cpu()
            wqf.backward(dwq_pt)
            assert torch.equal(bufs["wmask"].cpu().reshape(shape).float(), ((idx >= -128) & (idx <= 127)).float())
            torch.testing.assert_close(bufs["dweight"].cpu().reshape(shape), wl.grad, rtol=1e-6, atol=1e-8)
            if gamma is not None:
                exp = gl.grad + bufs["dgamma_bn"].cpu() + bufs["dsf_bn"].cpu() / torch.sqrt(var + 1e-5)
                torch.testing.assert_close(bufs["dgamma"].cpu(), exp, rtol=2e-5, atol=1e-5)


def _pitched(t2d, ld):
    """[M, C] uint8 -> a [M, ld] buffer (ld >= C) whose padding bytes are garbage; returns the buffer."""
    M, Cc = t2d.shape
    buf = torch.randint(0, 256, (M, ld), dtype=torch.uint8, device=t2d.device)
    buf[:, :Cc] = t2d
    return buf


def _r16(v):
    return (v + 15) // 16 * 16


# ------------------------------------------------------------------------------------------ convs
STAT_DT = np.dtype([("sum", "<i8"), ("lo", "<u8"), ("hi", "<u8"), ("mn", "<i4"), ("mx", "<i4")])


def _stats_buf(c):
    t = torch.zeros(c * 32, dtype=torch.uint8, device=DEV)
    L().call("frost_stats_reset", t.data_ptr(), c, stream())
    return t


def _check_stats(buf, I):
    """I: [M, C] int64 expected accumulators."""
    raw = np.frombuffer(buf.cpu().numpy().tobytes(), dtype=STAT_DT)
    I = I.numpy().astype(object)
    for c in range(I.shape[1]):
        col = I[:, c]
        assert int(raw["sum"][c]) == int(col.sum()), c
        assert int(raw["hi"][c]) * (1 << 32) + int(raw["lo"][c]) == int((col * col).sum()), c
        assert int(raw["mn"][c]) == int(col.min()) and int(raw["mx"][c]) == int(col.max()), c


@pytest.mark.parametrize("entry", ["frost_pw_conv_forward", "frost_pw_conv_forward_simt"])
@pytest.mark.parametrize("M,K,cout,zpw", [(1, 16, 16, 0), (130, 24, 24, 0), (257, 104, 312, 0), (1000, 1728, 320, 0),
                                          (300, 56, 40, -128), (64, 320, 1280, 0), (513, 96, 16, 127), (40000, 16, 96, 0),
                                          (12544, 288, 1728, 0), (5000, 168, 40, 127), (777, 1440, 192, -128)])
def test_pw_conv_forward_exact(M, K, cout, zpw, entry):
    g = torch.Generator().manual_seed(M + K)
    xq = torch.randint(0, 256, (M, K), generator=g, dtype=torch.int64)
    wq = torch.randint(-128, 128, (cout, K), generator=g, dtype=torch.int64)
    zpa = int(torch.randint(0, 256, (1,), generator=g))
    I = (xq - zpa) @ (wq - zpw).t()
    xd, wd = xq.to(torch.uint8).to(DEV), wq.to(torch.int8).to(DEV)
    za, zw = torch.tensor([zpa], dtype=torch.int32, device=DEV), torch.tensor([zpw], dtype=torch.int32, device=DEV)
    wsum = wq.sum(1).int().to(DEV)
    acc = torch.empty(M, cout, dtype=torch.int32, device=DEV)
    st = _stats_buf(cout)
    L().call(entry, xd.data_ptr(), za.data_ptr(), wd.data_ptr(), zw.data_ptr(), wsum.data_ptr(), M, K,
             cout, acc.data_ptr(), st.data_ptr(), stream())
    torch.cuda.synchronize()
    assert torch.equal(acc.cpu().long(), I)
    _check_stats(st, I)


def _conv_int_ref(xq, zpa, w, zpw, stride, pad, groups):
    """xq [N,H,W,C] int64, w [cout,cin_g,k,k] int64 -> I [N,Ho,Wo,cout] (float64 conv is exact here)."""
    x = (xq - zpa).permute(0, 3, 1, 2).double()
    y = Fn.conv2d(x, (w - zpw).double(), None, stride, pad, 1, groups)
    return y.permute(0, 2, 3, 1).round().long()


@pytest.mark.parametrize("N,H,W,Cc,k,s,zpw", [(2, 9, 11, 32, 3, 1, 0), (1, 14, 14, 96, 5, 2, 0), (3, 7, 7, 1728, 5, 1, 0),
                                              (2, 16, 15, 72, 3, 2, -128), (1, 5, 6, 1440, 5, 1, 0), (2, 3, 3, 16, 5, 1, 0),
                                              (6, 56, 56, 144, 5, 2, 0), (40, 28, 30, 32, 3, 1, 0), (5, 33, 57, 96, 3, 2, 127)])
@pytest.mark.parametrize("padded", [False, True])
def test_dw_conv_forward_exact(N, H, W, Cc, k, s, zpw, padded):
    g = torch.Generator().manual_seed(H * W + Cc)
    xq = torch.randint(0, 256, (N, H, W, Cc), generator=g, dtype=torch.int64)
    w = torch.randint(-128, 128, (Cc, 1, k, k), generator=g, dtype=torch.int64)
    zpa = int(torch.randint(0, 256, (1,), generator=g))
    I = _conv_int_ref(xq, zpa, w, zpw, s, (k - 1) // 2, Cc)
    xd = xq.to(torch.uint8).to(DEV)
    ldx = _r16(Cc) + 16 if padded else Cc                 # bytes between consecutive pixels
    if padded:
        xd = _pitched(xd.reshape(-1, Cc), ldx)
    wd = w.reshape(Cc, k * k).t().contiguous().to(torch.int8).to(DEV)
    za, zw = torch.tensor([zpa], dtype=torch.int32, device=DEV), torch.tensor([zpw], dtype=torch.int32, device=DEV)
    acc = torch.empty(I.shape, dtype=torch.int32, device=DEV)
    st = _stats_buf(Cc)
    L().call("frost_dw_conv_forward", xd.data_ptr(), ldx, za.data_ptr(), wd.data_ptr(), zw.data_ptr(), N, H, W, Cc, k, s,
             acc.data_ptr(), st.data_ptr(), stream())
    assert torch.equal(acc.cpu().long(), I)
    _check_stats(st, I.reshape(-1, Cc))


@pytest.mark.parametrize("N,H,W,cout", [(2, 224, 224, 32), (3, 33, 31, 16), (1, 8, 8, 24)])
def test_stem_conv_forward_exact(N, H, W, cout):
    g = torch.Generator().manual_seed(H + cout)
    xq = torch.randint(0, 256, (N, H, W, 3), generator=g, dtype=torch.int64)
    w = torch.randint(-128, 128, (cout, 3, 3, 3), generator=g, dtype=torch.int64)
    zpa = int(torch.randint(0, 256, (1,), generator=g))
    I = _conv_int_ref(xq, zpa, w, 0, 2, 1, 1)
    xd = xq.to(torch.uint8).to(DEV)
    wd = w.permute(0, 2, 3, 1).contiguous().to(torch.int8).to(DEV)
    za, zw = torch.tensor([zpa], dtype=torch.int32, device=DEV), torch.zeros(1, dtype=torch.int32, device=DEV)
    acc = torch.empty(I.shape, dtype=torch.int32, device=DEV)
    st = _stats_buf(cout)
    L().call("frost_stem_conv_forward", xd.data_ptr(), za.data_ptr(), wd.data_ptr(), zw.data_ptr(), N, H, W, 3, cout, 3,
             2, 1, acc.data_ptr(), st.data_ptr(), stream())
    assert torch.equal(acc.cpu().long(), I)
    _check_stats(st, I.reshape(-1, cout))


# ------------------------------------------------------------------------------------------ BN + FQ
def _bn_setup(M, Cc, seed):
    g = torch.Generator().manual_seed(seed)
    I = (torch.randn(M, Cc, generator=g) * 3000 + torch.randn(Cc, generator=g) * 2000).round().long()
    gamma = 0.5 + torch.rand(Cc, generator=g)
    gamma[0] = -0.7                                   # a negative gamma exercises the decreasing branch
    beta = 0.3 * torch.randn(Cc, generator=g)
    rm, rv = 0.1 * torch.randn(Cc, generator=g), 0.5 + torch.rand(Cc, generator=g)
    sf = gamma / torch.sqrt(rv + 1e-5)
    return I, gamma, beta, rm, rv, sf, 0.02, 0.003


def _bn_reference(I, gamma, beta, rm, rv, sf, s_a, s_w, relu, training, fq):
    """float64 restatement of conv_fused.py:156-167 + hook, on conv = s_a*s_w*I."""
    conv = (I.double() * (float(np.float32(s_a)) * float(np.float32(s_w)))).requires_grad_(True)
    g64 = gamma.double().requires_grad_(True)
    b64 = beta.double().requires_grad_(True)
    u = conv / sf.double()
    rm64, rv64 = rm.double().clone(), rv.double().clone()
    v = Fn.batch_norm(u, rm64, rv64, g64, b64, training, 0.1, 1e-5)
    r = torch.relu(v) if relu else v
    y = fq(r.float())
    return conv, g64, b64, v, r, y, rm64, rv64


@pytest.mark.parametrize("M,Cc,relu,training", [(64, 16, True, True), (1000, 24, False, True), (513, 1728, True, True),
                                                (200, 40, True, False), (20000, 360, True, True), (70000, 72, True, True)])
def test_bn_finalize_apply_and_backward(M, Cc, relu, training):
    I, gamma, beta, rm, rv, sf, s_a, s_w = _bn_setup(M, Cc, M + Cc)
    Id = I.int().to(DEV)
    rec = np.zeros(Cc, dtype=STAT_DT)
    In = I.numpy().astype(object)
    for c in range(Cc):
        col = In[:, c]
        sq = int((col * col).sum())
        rec[c] = (int(col.sum()), sq & 0xffffffff, sq >> 32, int(col.min()), int(col.max()))
    st = torch.frombuffer(bytearray(rec.tobytes()), dtype=torch.uint8).to(DEV)
    fq = DevFQ(0, 255, False)
    t = lambda x, dt=torch.float32: x.to(dt).to(DEV).contiguous()
    xs, ws = t(torch.tensor([s_a])), t(torch.tensor([s_w]))
    gd, bd, rmd, rvd, sfd = t(gamma), t(beta), t(rm), t(rv), t(sf)
    nbt = torch.zeros((), dtype=torch.int64, device=DEV)
    A, B, meanI, kfac = (torch.zeros(Cc, device=DEV) for _ in range(4))
    mm = torch.zeros(2, device=DEV)
    a = L().BnFinalizeArgs()
    a.stats, a.C, a.count = st.data_ptr(), Cc, M
    a.x_scale, a.w_scale, a.sf = xs.data_ptr(), ws.data_ptr(), sfd.data_ptr()
    a.gamma, a.beta, a.running_mean, a.running_var = gd.data_ptr(), bd.data_ptr(), rmd.data_ptr(), rvd.data_ptr()
    a.num_batches_tracked = nbt.data_ptr()
    a.momentum, a.eps, a.training, a.relu, a.observe, a.averaging_const = 0.1, 1e-5, int(training), int(relu), 1, 0.01
    a.afq = fq.c()
    a.A, a.B, a.mean_I, a.kfac, a.cur_minmax = A.data_ptr(), B.data_ptr(), meanI.data_ptr(), kfac.data_ptr(), mm.data_ptr()
    L().call("frost_bn_finalize", C.byref(a), stream())
    ldq = _r16(Cc) + (16 if M % 2 else 0)                 # row-padded output (dense when C % 16 == 0 and M is even)
    qbuf = torch.full((M, ldq), 0xAB, dtype=torch.uint8, device=DEV)
    L().call("frost_bnq_apply", Id.data_ptr(), 0, M, Cc, A.data_ptr(), B.data_ptr(), int(relu), fq.scale.data_ptr(),
             fq.zp.data_ptr(), qbuf.data_ptr(), ldq, stream())
    q = qbuf[:, :Cc].contiguous()
    assert bool((qbuf[:, Cc:] == 0xAB).all())             # the row padding is not touched
    conv, g64, b64, v, r, y, rm64, rv64 = _bn_reference(I, gamma, beta, rm, rv, sf, s_a, s_w, relu, training, fq.oracle)
    # observer state: min/max of the pre-quant tensor agree to fp32 rounding; qparams follow
    assert abs(float(fq.min_val) - float(fq.oracle.min_val)) <= 2e-6 * max(1.0, abs(float(fq.oracle.min_val)))
    assert abs(float(fq.max_val) - float(fq.oracle.max_val)) <= 2e-6 * max(1.0, abs(float(fq.oracle.max_val)))
    assert abs(float(fq.scale) - float(fq.oracle.scale)) <= 2e-6 * float(fq.oracle.scale)
    assert abs(int(fq.zp) - int(fq.oracle.zero_point)) <= 1
    exp_q = fq.oracle.last_idx.clamp(0, 255)
    d = (q.cpu().float() - exp_q).abs()
    assert float(d.max()) <= 1 and float((d > 0).float().mean()) < 2e-3
    # the dequantised min/max that downstream cat observers consume
    s_o, zp_o = float(fq.scale), int(fq.zp)
    deq = (q.cpu().float() - zp_o) * s_o
    assert abs(float(mm[0]) - float(deq.min())) < 1e-6 and abs(float(mm[1]) - float(deq.max())) < 1e-6
    if training:
        torch.testing.assert_close(rmd.cpu().double(), rm64, rtol=1e-5, atol=1e-6)
        torch.testing.assert_close(rvd.cpu().double(), rv64, rtol=1e-5, atol=1e-6)
        assert int(nbt) == 1
    else:
        assert torch.equal(rmd.cpu(), rm) and torch.equal(rvd.cpu(), rv) and int(nbt) == 0
        return
    # ---- backward: STE mask taken at the device's own qparams so the comparison is about the BN algebra
    gen = torch.Generator().manual_seed(7)
    dy = torch.randn(M, Cc, generator=gen)
    idx = torch.round(r.detach().float() * float(np.float32(1.0) / np.float32(s_o))) + zp_o
    maskf = ((idx >= 0) & (idx <= 255)).double()
    (r * (dy.double() * maskf)).sum().backward()
    dyd = t(dy)
    dz = torch.empty(M, Cc, device=DEV)
    sums = torch.zeros(2 * Cc, dtype=torch.float64, device=DEV)
    coef = torch.zeros(3 * Cc, device=DEV)
    dgb, dbeta, dsf = (torch.zeros(Cc, device=DEV) for _ in range(3))
    b = L().BnBackwardArgs()
    b.dy, b.acc, b.M, b.C, b.relu = dyd.data_ptr(), Id.data_ptr(), M, Cc, int(relu)
    b.A, b.B, b.mean_I, b.kfac = A.data_ptr(), B.data_ptr(), meanI.data_ptr(), kfac.data_ptr()
    b.gamma, b.sf, b.x_scale, b.w_scale = gd.data_ptr(), sfd.data_ptr(), xs.data_ptr(), ws.data_ptr()
    b.out_scale, b.out_zp, b.eps = fq.scale.data_ptr(), fq.zp.data_ptr(), 1e-5
    b.sums, b.coef, b.dz = sums.data_ptr(), coef.data_ptr(), dz.data_ptr()
    b.dgamma_bn, b.dbeta, b.dsf_bn = dgb.data_ptr(), dbeta.data_ptr(), dsf.data_ptr()
    L().call("frost_bn_backward", C.byref(b), stream())
    scale = float(conv.grad.abs().max())
    err = float((dz.cpu().double() - conv.grad).abs().max())
    assert err <= 2e-4 * scale, (err, scale)
    # bf16 hi/lo plane format (operand format of the tensor-core dgrad / wgrad)
    p_hi = torch.empty(M, Cc, dtype=torch.bfloat16, device=DEV)
    p_lo = torch.empty(M, Cc, dtype=torch.bfloat16, device=DEV)
    b.dz, b.dz_lo, b.dz_format = p_hi.data_ptr(), p_lo.data_ptr(), 1
    L().call("frost_bn_backward", C.byref(b), stream())
    rec = p_hi.float() + p_lo.float()
    assert float((rec - dz).abs().max()) <= 2.0 ** -16 * float(dz.abs().max())
    assert torch.equal(p_hi, dz.to(torch.bfloat16))
    torch.testing.assert_close(dbeta.cpu().double(), b64.grad, rtol=1e-4, atol=1e-4 * float(b64.grad.abs().max()))
    torch.testing.assert_close(dgb.cpu().double(), g64.grad, rtol=1e-3, atol=2e-4 * float(g64.grad.abs().max()))


# ------------------------------------------------------------------------------------------ cat / add
def _rand_qt(M, Cc, seed, ld=None):
    g = torch.Generator().manual_seed(seed)
    q = torch.randint(0, 256, (M, Cc), generator=g, dtype=torch.int64)
    scale = float(np.float32(0.01 + 0.05 * float(torch.rand(1, generator=g))))
    zp = int(torch.randint(0, 200, (1,), generator=g))
    val = ((q - zp).float() * scale)
    qd = q.to(torch.uint8).to(DEV)
    d = dict(q=qd if ld is None else _pitched(qd, ld), ld=Cc if ld is None else ld, scale=torch.tensor([scale], device=DEV),
             zp=torch.tensor([zp], dtype=torch.int32, device=DEV), mm=torch.tensor([float(val.min()), float(val.max())], device=DEV))
    return d, val


def _qt(d, Cc):
    return L().QTensor(d["q"].data_ptr(), d["scale"].data_ptr(), d["zp"].data_ptr(), d["mm"].data_ptr(), Cc, d["ld"])


@pytest.mark.parametrize("padded", [False, True])
def test_cat_forward_backward_bit_exact(padded):
    M, C1, C2 = 333, 16, 40
    fq = DevFQ(0, 255, False)
    for it in range(2):
        a, va = _rand_qt(M, C1, 10 + it, 32 if padded else None)
        b, vb = _rand_qt(M, C2, 20 + it, 48 if padded else None)
        ld_out = 64 if padded else C1 + C2
        obuf = torch.full((M, ld_out), 0xAB, dtype=torch.uint8, device=DEV)
        mm = torch.empty(2, device=DEV)
        L().call("frost_cat_forward", _qt(a, C1), _qt(b, C2), M, fq.c(), 1, 0.01, obuf.data_ptr(), ld_out, mm.data_ptr(), stream())
        out = obuf[:, :C1 + C2]
        assert bool((obuf[:, C1 + C2:] == 0xAB).all())
        x = torch.cat([va, vb], 1).requires_grad_(True)
        y = fq.oracle(x)
        fq.assert_state_equal()
        assert torch.equal(out.cpu().float(), fq.oracle.last_idx.clamp(0, 255))
        assert float(mm[0]) == float(y.min()) and float(mm[1]) == float(y.max())
        dcat = torch.randn(M, C1 + C2)
        y.backward(dcat)
        da = torch.empty(M, C1, device=DEV)
        db = torch.ones(M, C2, device=DEV)
        L().call("frost_cat_backward", dcat.to(DEV).data_ptr(), _qt(a, C1), _qt(b, C2), M, fq.scale.data_ptr(),
                 fq.zp.data_ptr(), da.data_ptr(), db.data_ptr(), 1, stream())
        assert torch.equal(da.cpu(), x.grad[:, :C1])
        assert torch.equal(db.cpu(), x.grad[:, C1:] + 1.0)


@pytest.mark.parametrize("padded", [False, True])
def test_add_forward_backward_bit_exact(padded):
    M, Cc = 257, 24
    fq = DevFQ(0, 255, False)
    sc = scratch()
    for it in range(2):
        a, va = _rand_qt(M, Cc, 30 + it, 32 if padded else None)
        b, vb = _rand_qt(M, Cc, 40 + it, 48 if padded else None)
        ld_out = 32 if padded else Cc
        obuf = torch.full((M, ld_out), 0xAB, dtype=torch.uint8, device=DEV)
        mm = torch.empty(2, device=DEV)
        L().call("frost_add_forward", _qt(a, Cc), _qt(b, Cc), M * Cc, fq.c(), 1, 0.01, obuf.data_ptr(), ld_out, mm.data_ptr(),
                 sc.data_ptr(), stream())
        out = obuf[:, :Cc]
        assert bool((obuf[:, Cc:] == 0xAB).all())
        x = (va + vb).requires_grad_(True)
        y = fq.oracle(x)
        fq.assert_state_equal()
        assert torch.equal(out.cpu().float(), fq.oracle.last_idx.clamp(0, 255))
        assert float(mm[0]) == float(y.min()) and float(mm[1]) == float(y.max())
        dout = torch.randn(M, Cc)
        y.backward(dout)
        dsum = torch.empty(M, Cc, device=DEV)
        da = torch.empty(M, Cc, device=DEV)
        L().call("frost_add_backward", dout.to(DEV).data_ptr(), _qt(a, Cc), _qt(b, Cc), M * Cc, fq.scale.data_ptr(),
                 fq.zp.data_ptr(), dsum.data_ptr(), da.data_ptr(), 0, stream())
        assert torch.equal(dsum.cpu(), x.grad) and torch.equal(da.cpu(), x.grad)


def _split_bf16(x):
    hi = x.to(torch.bfloat16)
    lo = (x - hi.float()).to(torch.bfloat16)
    return hi.contiguous(), lo.contiguous()


# ------------------------------------------------------------------------------------------ dgrad / wgrad
@pytest.mark.parametrize("M,K,cout", [(70, 16, 24), (1000, 104, 312), (333, 1728, 320), (64, 320, 1280), (20000, 56, 168),
                                      (5000, 288, 1728)])
def test_pw_dgrad_wgrad(M, K, cout):
    g = torch.Generator().manual_seed(M)
    dz = torch.randn(M, cout, generator=g)
    wq = torch.randint(-128, 128, (cout, K), generator=g)
    xq = torch.randint(0, 256, (M, K), generator=g)
    s_w, zp_w, s_a, zp_a = 0.004, 0, 0.03, 77
    wf = (wq - zp_w).double() * float(np.float32(s_w))
    xf = (xq - zp_a).double() * float(np.float32(s_a))
    dx_ref = dz.double() @ wf
    dw_ref = dz.double().t() @ xf
    sw_t, sa_t = torch.tensor([s_w], device=DEV), torch.tensor([s_a], device=DEV)
    zw_t, za_t = torch.tensor([zp_w], dtype=torch.int32, device=DEV), torch.tensor([zp_a], dtype=torch.int32, device=DEV)
    dzd, wd, xd = dz.to(DEV), wq.to(torch.int8).to(DEV), xq.to(torch.uint8).to(DEV)
    dz_hi, dz_lo = _split_bf16(dzd)
    dx = torch.ones(M, K, device=DEV)
    L().call("frost_pw_dgrad", dzd.data_ptr(), wd.data_ptr(), sw_t.data_ptr(), zw_t.data_ptr(), M, K, cout,
             dx.data_ptr(), 1, stream())
    torch.testing.assert_close(dx.cpu().double(), dx_ref + 1.0, rtol=1e-4, atol=1e-4 * float(dx_ref.abs().max()))
    # tensor-core version (bf16 hi/lo split of dz, exact integer weights) from the transposed weights
    wtd = (wq - zp_w).t().contiguous().to(torch.bfloat16).to(DEV)
    for accumulate in (0, 1):
        dx2 = torch.ones(M, K, device=DEV)
        L().call("frost_pw_dgrad_tc", dz_hi.data_ptr(), dz_lo.data_ptr(), wtd.data_ptr(), sw_t.data_ptr(), M, K, cout,
                 dx2.data_ptr(), accumulate, stream())
        torch.cuda.synchronize()
        ref2 = dx_ref + (1.0 if accumulate else 0.0)
        err = float((dx2.cpu().double() - ref2).abs().max()) / float(dx_ref.abs().max())
        assert err < 5e-5, (accumulate, err)
    # integer-valued dz: products and sums are exact in the tensor-core path
    dzi = torch.randint(-64, 65, (M, cout), generator=g).float()
    dx3 = torch.empty(M, K, device=DEV)
    one = torch.ones(1, device=DEV)
    i_hi, i_lo = _split_bf16(dzi.to(DEV))
    L().call("frost_pw_dgrad_tc", i_hi.data_ptr(), i_lo.data_ptr(), wtd.data_ptr(), one.data_ptr(), M, K, cout,
             dx3.data_ptr(), 0, stream())
    torch.cuda.synchronize()
    assert torch.equal(dx3.cpu().double(), dzi.double() @ (wq - zp_w).double())
    dwq = torch.empty(cout, K, device=DEV)
    L().call("frost_pw_wgrad", dzd.data_ptr(), xd.data_ptr(), sa_t.data_ptr(), za_t.data_ptr(), M, K, cout,
             dwq.data_ptr(), stream())
    torch.testing.assert_close(dwq.cpu().double(), dw_ref, rtol=1e-4, atol=1e-4 * float(dw_ref.abs().max()))
    dwq2 = torch.empty(cout, K, device=DEV)
    L().call("frost_pw_wgrad_tc", dz_hi.data_ptr(), dz_lo.data_ptr(), xd.data_ptr(), K, sa_t.data_ptr(), za_t.data_ptr(), M, K,
             cout, dwq2.data_ptr(), stream())
    torch.cuda.synchronize()
    ldx = _r16(K) + 16                                   # the same through row-padded x
    xp = _pitched(xd, ldx)
    dwq3 = torch.empty(cout, K, device=DEV)
    L().call("frost_pw_wgrad_tc", dz_hi.data_ptr(), dz_lo.data_ptr(), xp.data_ptr(), ldx, sa_t.data_ptr(), za_t.data_ptr(), M, K,
             cout, dwq3.data_ptr(), stream())
    assert float((dwq3 - dwq2).abs().max()) <= 1e-4 * float(dwq2.abs().max())
    torch.cuda.synchronize()
    errw = float((dwq2.cpu().double() - dw_ref).abs().max()) / float(dw_ref.abs().max())
    assert errw < 5e-5, errw
    if M <= 1000:      # integer-valued dz: exact
        dzs = torch.randint(-8, 9, (M, cout), generator=g).float()
        one_w = torch.ones(1, device=DEV)
        s_hi, s_lo = _split_bf16(dzs.to(DEV))
        L().call("frost_pw_wgrad_tc", s_hi.data_ptr(), s_lo.data_ptr(), xd.data_ptr(), K, one_w.data_ptr(), za_t.data_ptr(), M, K,
                 cout, dwq2.data_ptr(), stream())
        torch.cuda.synchronize()
        assert torch.equal(dwq2.cpu().double(), dzs.double().t() @ (xq - zp_a).double())


@pytest.mark.parametrize("N,H,W,Cc,k,s", [(2, 9, 11, 32, 3, 1), (1, 14, 14, 96, 5, 2), (2, 7, 7, 1728, 5, 1), (2, 16, 15, 72, 3, 2),
                                          (6, 56, 56, 144, 5, 2), (40, 28, 30, 32, 3, 1), (5, 33, 57, 96, 3, 1)])
def test_dw_dgrad_wgrad(N, H, W, Cc, k, s):
    g = torch.Generator().manual_seed(H + Cc)
    pad = (k - 1) // 2
    xq = torch.randint(0, 256, (N, H, W, Cc), generator=g)
    wq = torch.randint(-128, 128, (Cc, 1, k, k), generator=g)
    s_w, zp_w, s_a, zp_a = 0.004, 0, 0.03, 77
    xf = ((xq - zp_a).double() * float(np.float32(s_a))).permute(0, 3, 1, 2).requires_grad_(True)
    wf = ((wq - zp_w).double() * float(np.float32(s_w))).requires_grad_(True)
    y = Fn.conv2d(xf, wf, None, s, pad, 1, Cc)
    dz = torch.randn(y.shape, generator=g, dtype=torch.float64)
    y.backward(dz)
    dz_nhwc = dz.permute(0, 2, 3, 1).contiguous().float().to(DEV)
    sw_t, sa_t = torch.tensor([s_w], device=DEV), torch.tensor([s_a], device=DEV)
    zw_t, za_t = torch.tensor([zp_w], dtype=torch.int32, device=DEV), torch.tensor([zp_a], dtype=torch.int32, device=DEV)
    wd = wq.reshape(Cc, k * k).t().contiguous().to(torch.int8).to(DEV)
    xd = xq.to(torch.uint8).to(DEV)
    dx = torch.zeros(N, H, W, Cc, device=DEV)
    L().call("frost_dw_dgrad", dz_nhwc.data_ptr(), wd.data_ptr(), sw_t.data_ptr(), zw_t.data_ptr(), N, H, W, Cc, k,
             s, dx.data_ptr(), 0, stream())
    ref = xf.grad.permute(0, 2, 3, 1)
    torch.testing.assert_close(dx.cpu().double(), ref, rtol=1e-4, atol=1e-5 * float(ref.abs().max()))
    dwq = torch.empty(k * k, Cc, device=DEV)
    L().call("frost_dw_wgrad", dz_nhwc.data_ptr(), xd.data_ptr(), Cc, sa_t.data_ptr(), za_t.data_ptr(), N, H, W, Cc, k,
             s, dwq.data_ptr(), stream())
    refw = wf.grad.reshape(Cc, k * k).t()
    torch.testing.assert_close(dwq.cpu().double(), refw, rtol=1e-4, atol=1e-4 * float(refw.abs().max()))
    ldx = _r16(Cc) + 16                                   # the same through row-padded x
    xp = _pitched(xd.reshape(-1, Cc), ldx)
    dwq_p = torch.empty(k * k, Cc, device=DEV)
    L().call("frost_dw_wgrad", dz_nhwc.data_ptr(), xp.data_ptr(), ldx, sa_t.data_ptr(), za_t.data_ptr(), N, H, W, Cc, k,
             s, dwq_p.data_ptr(), stream())
    torch.testing.assert_close(dwq_p.cpu().double(), refw, rtol=1e-4, atol=1e-4 * float(refw.abs().max()))


@pytest.mark.parametrize("N,H,W,cout", [(2, 33, 31, 16), (5, 224, 224, 32), (1, 8, 8, 24)])
def test_stem_wgrad(N, H, W, cout):
    g = torch.Generator().manual_seed(5)
    xq = torch.randint(0, 256, (N, H, W, 3), generator=g)
    s_a, zp_a = 0.02, 120
    xf = ((xq - zp_a).double() * float(np.float32(s_a))).permute(0, 3, 1, 2)
    wf = torch.zeros(cout, 3, 3, 3, dtype=torch.float64, requires_grad=True)
    y = Fn.conv2d(xf, wf, None, 2, 1)
    dz = torch.randn(y.shape, generator=g, dtype=torch.float64)
    y.backward(dz)
    dzd = dz.permute(0, 2, 3, 1).contiguous().float().to(DEV)
    dwq = torch.empty(cout, 3, 3, 3, device=DEV)
    xd = xq.to(torch.uint8).to(DEV)
    sa_t = torch.tensor([s_a], device=DEV)
    za_t = torch.tensor([zp_a], dtype=torch.int32, device=DEV)
    L().call("frost_stem_wgrad", dzd.data_ptr(), xd.data_ptr(), sa_t.data_ptr(), za_t.data_ptr(), N, H, W, 3, cout, 3, 2, 1,
             dwq.data_ptr(), stream())
    ref = wf.grad.permute(0, 2, 3, 1)
    torch.testing.assert_close(dwq.cpu().double(), ref, rtol=1e-4, atol=1e-4 * float(ref.abs().max()))


# ------------------------------------------------------------------------------------------ head
def test_pool_dropout_and_linear():
    N, HW, Cc, cout = 5, 49, 1280, 1000
    g = torch.Generator().manual_seed(9)
    a, va = _rand_qt(N * HW, Cc, 50)
    keep = (torch.rand(N, Cc, generator=g) > 0.2).float()
    keepd = keep.to(DEV)
    pooled = torch.empty(N, Cc, device=DEV)
    L().call("frost_pool_dropout_forward", a["q"].data_ptr(), a["scale"].data_ptr(), a["zp"].data_ptr(), N, HW, Cc,
             keepd.data_ptr(), 1.25, pooled.data_ptr(), stream())
    ref = va.reshape(N, HW, Cc).double().mean(1) * keep.double() * 1.25
    torch.testing.assert_close(pooled.cpu().double(), ref, rtol=2e-6, atol=1e-7)
    dpooled = torch.randn(N, Cc, generator=g)
    dpd = dpooled.to(DEV)
    dy = torch.empty(N, HW, Cc, device=DEV)
    L().call("frost_pool_dropout_backward", dpd.data_ptr(), N, HW, Cc, keepd.data_ptr(), 1.25, dy.data_ptr(), stream())
    refd = (dpooled * keep * 1.25 / HW).reshape(N, 1, Cc).expand(N, HW, Cc)
    torch.testing.assert_close(dy.cpu(), refd, rtol=1e-6, atol=1e-9)
    # classifier GEMMs
    wq = torch.randint(-128, 128, (cout, Cc), generator=g)
    bias = torch.randn(cout, generator=g)
    s_w = 0.002
    x = pooled.cpu()
    wf = wq.double() * float(np.float32(s_w))
    out = torch.empty(N, cout, device=DEV)
    t32 = torch.tensor([s_w], device=DEV)
    zi = torch.zeros(1, dtype=torch.int32, device=DEV)
    wd = wq.to(torch.int8).to(DEV)
    biasd = bias.to(DEV)
    L().call("frost_linear_forward", pooled.data_ptr(), wd.data_ptr(), t32.data_ptr(), zi.data_ptr(), biasd.data_ptr(),
             N, Cc, cout, out.data_ptr(), stream())
    refo = x.double() @ wf.t() + bias.double()
    torch.testing.assert_close(out.cpu().double(), refo, rtol=1e-5, atol=1e-5 * float(refo.abs().max()))
    dout = torch.randn(N, cout, generator=g)
    doutd = dout.to(DEV)
    dx, dwq, db = torch.empty(N, Cc, device=DEV), torch.empty(cout, Cc, device=DEV), torch.empty(cout, device=DEV)
    L().call("frost_linear_backward", doutd.data_ptr(), pooled.data_ptr(), wd.data_ptr(), t32.data_ptr(), zi.data_ptr(),
             N, Cc, cout, dx.data_ptr(), dwq.data_ptr(), db.data_ptr(), stream())
    torch.testing.assert_close(dx.cpu().double(), dout.double() @ wf, rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(dwq.cpu().double(), dout.double().t() @ x.double(), rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(db.cpu().double(), dout.double().sum(0), rtol=1e-5, atol=1e-5)


def test_bad_arguments_are_reported_not_crashed():
    with pytest.raises(RuntimeError, match="multiple of 8"):
        L().call("frost_pw_conv_forward", 8, 8, 8, 8, 8, 4, 12, 16, 8, 8, None)
    with pytest.raises(RuntimeError):
        L().call("frost_dw_conv_forward", 8, 16, 8, 8, 8, 1, 4, 4, 16, 7, 1, 8, 8, None)


# ------------------------------------------------------------------------------------------ launch-shape knobs
@pytest.mark.parametrize("knob,values", [(0, [1, 16]), (4, [1, 8]), (11, [2, 3])])
def test_results_do_not_depend_on_launch_knobs(knob, values):
    """FROST_TUNE_* only reshape grids: integer outputs and statistics must be identical for every value."""
    lib = L().load()
    outs = []
    try:
        for v in values:
            assert lib.frost_set_tunable(knob, v) == 0 and lib.frost_get_tunable(knob) == v
            if knob in (0, 11):
                test_dw_conv_forward_exact(3, 7, 7, 1728, 5, 1, 0, True)
                test_dw_conv_forward_exact(2, 9, 11, 32, 3, 1, 0, False)
                test_dw_conv_forward_exact(6, 56, 56, 144, 5, 2, 0, True)
                test_dw_conv_forward_exact(2, 112, 112, 32, 3, 1, 0, False)
            else:
                test_stem_conv_forward_exact(2, 224, 224, 32)
            outs.append(v)
    finally:
        lib.frost_set_tunable(knob, 0)
    assert lib.frost_set_tunable(99, 1) != 0 and b"unknown knob" in lib.frost_last_error()


def test_device_prefetcher_yields_every_batch_in_order():
    import frostnet_b200 as F
    batches = [(torch.full((4, 3, 8, 8), float(i)).pin_memory(), torch.full((4,), i, dtype=torch.int64).pin_memory())
               for i in range(5)]
    seen = []
    for x, y in F.DevicePrefetcher(batches, DEV):
        assert x.is_cuda and y.is_cuda
        seen.append((float(x.mean()), int(y[0])))
        x.mul_(2.0)       # the consumer may scribble on the buffer; the next copy waits for it
    assert seen == [(float(i), i) for i in range(5)]
    assert list(F.DevicePrefetcher([], DEV)) == []
