"""The fused 1x1 ConvBn(ReLU)2d kernels (csrc/pw_fused.cu: tcgen05 GEMM + BatchNorm + ReLU + fake-quant in one launch,
accumulators recomputed instead of stored) against the first-generation chain
    frost_pw_conv_forward -> frost_bn_finalize -> frost_bnq_apply      /  frost_bn_backward_reduce -> _apply
which tests/test_gpu_ops.py pins against exact integer arithmetic, the oracle's fake-quant and float64 autograd.
Forward: every output and every side effect BIT-EXACT (uint8 indices, BN affine, running statistics, observer state,
qparams).  Backward: the per-channel sums to fp64 rounding (different summation order), the bf16 planes accordingly.
Operands are staged by TMA: row pitches are multiples of 16 bytes (the engine pads; padding bytes are garbage here)."""
import ctypes as C

import pytest
import torch

pytestmark = pytest.mark.gpu

DEV = "cuda:0"


def L():
    from frostnet_b200 import _lib
    return _lib


def stream():
    return torch.cuda.current_stream().cuda_stream


def _r16(v):
    return (v + 15) // 16 * 16


class Layer:
    """Device state of one ConvBn(ReLU)2d + its activation fake-quant, cloneable (legacy vs fused runs)."""

    def __init__(self, K, cout, zpw, relu, seed):
        g = torch.Generator().manual_seed(seed)
        self.K, self.cout, self.zpw, self.relu = K, cout, zpw, relu
        wq = torch.randint(-128, 128, (cout, K), generator=g, dtype=torch.int64)
        self.wq = wq.to(torch.int8).to(DEV)
        self.ldw = _r16(K)
        flip = {0: 0, -128: 0x80, 127: 0x7f}[zpw]
        mma = torch.zeros(cout, self.ldw, dtype=torch.uint8)
        mma[:, :K] = (wq & 0xff).to(torch.uint8) ^ flip
        self.w_mma = mma.to(DEV)
        self.wsum = wq.sum(1).int().to(DEV)
        self.w_zp = torch.tensor([zpw], dtype=torch.int32, device=DEV)
        self.gamma = (0.5 + torch.rand(cout, generator=g)).to(DEV)
        self.gamma[0] = -0.7
        self.beta = (0.3 * torch.randn(cout, generator=g)).to(DEV)
        self.rm = (0.1 * torch.randn(cout, generator=g)).to(DEV)
        self.rv = (0.5 + torch.rand(cout, generator=g)).to(DEV)
        self.sf = (self.gamma / torch.sqrt(self.rv + 1e-5)).contiguous()
        self.nbt = torch.zeros((), dtype=torch.int64, device=DEV)
        self.x_scale = torch.tensor([0.02], device=DEV)
        self.w_scale = torch.tensor([0.003], device=DEV)
        self.mn = torch.tensor(float("inf"), device=DEV)
        self.mx = torch.tensor(float("-inf"), device=DEV)
        self.scale = torch.ones(1, device=DEV)
        self.zp = torch.zeros(1, dtype=torch.int32, device=DEV)
        self.A, self.B, self.mean_I, self.kfac = (torch.zeros(cout, device=DEV) for _ in range(4))
        self.mm = torch.zeros(2, device=DEV)
        self.stats = torch.zeros(cout * 32, dtype=torch.uint8, device=DEV)
        self.sums = torch.zeros(2 * cout, dtype=torch.float64, device=DEV)
        self.coef = torch.zeros(3 * cout, device=DEV)
        self.dgb, self.dbeta, self.dsf = (torch.zeros(cout, device=DEV) for _ in range(3))

    STATE = ("rm", "rv", "nbt", "mn", "mx", "scale", "zp", "A", "B", "mean_I", "kfac", "mm")

    def clone_state_from(self, o):
        for k in self.STATE:
            getattr(self, k).copy_(getattr(o, k))

    def fin_args(self, M, training, observe):
        a = L().BnFinalizeArgs()
        L().call("frost_stats_reset", self.stats.data_ptr(), self.cout, stream())
        a.stats, a.C, a.count = self.stats.data_ptr(), self.cout, M
        a.x_scale, a.w_scale, a.sf = self.x_scale.data_ptr(), self.w_scale.data_ptr(), self.sf.data_ptr()
        a.gamma, a.beta, a.running_mean, a.running_var = self.gamma.data_ptr(), self.beta.data_ptr(), self.rm.data_ptr(), self.rv.data_ptr()
        a.num_batches_tracked = self.nbt.data_ptr()
        a.momentum, a.eps, a.training, a.relu, a.observe, a.averaging_const = 0.1, 1e-5, int(training), int(self.relu), int(observe), 0.01
        a.afq = L().FQ(self.mn.data_ptr(), self.mx.data_ptr(), self.scale.data_ptr(), self.zp.data_ptr())
        a.A, a.B, a.mean_I, a.kfac, a.cur_minmax = (self.A.data_ptr(), self.B.data_ptr(), self.mean_I.data_ptr(), self.kfac.data_ptr(),
                                                    self.mm.data_ptr())
        return a

    def bwd_args(self, M, dy, acc, dz_hi, dz_lo, frozen):
        b = L().BnBackwardArgs()
        b.dy, b.acc, b.M, b.C, b.relu = dy.data_ptr(), (acc.data_ptr() if acc is not None else None), M, self.cout, int(self.relu)
        b.A, b.B, b.mean_I, b.kfac = self.A.data_ptr(), self.B.data_ptr(), self.mean_I.data_ptr(), self.kfac.data_ptr()
        b.gamma, b.sf, b.x_scale, b.w_scale = self.gamma.data_ptr(), self.sf.data_ptr(), self.x_scale.data_ptr(), self.w_scale.data_ptr()
        b.out_scale, b.out_zp, b.eps = self.scale.data_ptr(), self.zp.data_ptr(), 1e-5
        b.sums, b.coef, b.dz, b.dz_lo, b.dz_format = self.sums.data_ptr(), self.coef.data_ptr(), dz_hi.data_ptr(), dz_lo.data_ptr(), 1
        b.dgamma_bn, b.dbeta, b.dsf_bn = self.dgb.data_ptr(), self.dbeta.data_ptr(), self.dsf.data_ptr()
        b.frozen = int(frozen)
        return b

    def operands(self, x, M, ldx, x_zp):
        op = L().PwOperands()
        op.x, op.M, op.K, op.ldx = x.data_ptr(), M, self.K, ldx
        op.w_mma, op.ldw, op.cout = self.w_mma.data_ptr(), self.ldw, self.cout
        op.x_zp, op.w_zp, op.wsum = x_zp.data_ptr(), self.w_zp.data_ptr(), self.wsum.data_ptr()
        return op


def _run_case(M, K, cout, zpw, relu, training, observe, extra_pad=0, steps=2, check_bwd=True):
    if True:
        ref, fus = Layer(K, cout, zpw, relu, M + K), Layer(K, cout, zpw, relu, M + K)
        ldx = _r16(K) + extra_pad
        ldq = _r16(cout) + extra_pad
        gd = torch.Generator(device=DEV).manual_seed(M + cout)
        x_zp = torch.tensor([7], dtype=torch.int32, device=DEV)
        for step in range(steps):
            xp = torch.randint(0, 256, (M, ldx), generator=gd, dtype=torch.uint8, device=DEV)   # pad bytes are garbage on purpose
            xd = xp[:, :K].contiguous()
            # ---- first-generation chain
            a = ref.fin_args(M, training, observe)
            acc = torch.empty(M, cout, dtype=torch.int32, device=DEV)
            L().call("frost_pw_conv_forward", xd.data_ptr(), x_zp.data_ptr(), ref.wq.data_ptr(), ref.w_zp.data_ptr(), ref.wsum.data_ptr(),
                     M, K, cout, acc.data_ptr(), ref.stats.data_ptr(), stream())
            L().call("frost_bn_finalize", C.byref(a), stream())
            q_ref = torch.empty(M, cout, dtype=torch.uint8, device=DEV)
            L().call("frost_bnq_apply", acc.data_ptr(), 0, M, cout, ref.A.data_ptr(), ref.B.data_ptr(), int(relu), ref.scale.data_ptr(),
                     ref.zp.data_ptr(), q_ref.data_ptr(), cout, stream())
            # ---- fused
            f = L().PwFusedFwdArgs()
            f.op, f.bn = fus.operands(xp, M, ldx, x_zp), fus.fin_args(M, training, observe)
            bar = torch.zeros(1, dtype=torch.int32, device=DEV)
            q = torch.full((M, ldq), 0xAB, dtype=torch.uint8, device=DEV)
            f.grid_barrier, f.q, f.ldq = bar.data_ptr(), q.data_ptr(), ldq
            L().call("frost_pw_fused_forward", C.byref(f), stream())
            torch.cuda.synchronize()
            assert torch.equal(q[:, :cout], q_ref), "indices differ: %d of %d" % (int((q[:, :cout] != q_ref).sum()), q_ref.numel())
            # (the row padding [cout, ldq) is don't-care: the TMA store moves whole 16-byte units)
            if training or observe:
                # FrostChanStats: sum, sq_lo, sq_hi, (min | max << 32); the lo/hi split of the sum of squares depends on
                # how the partial sums were grouped - compare the 96-bit value, not its representation
                fs, rs = fus.stats.view(torch.int64).reshape(-1, 4), ref.stats.view(torch.int64).reshape(-1, 4)
                assert torch.equal(fs[:, 0], rs[:, 0]) and torch.equal(fs[:, 3], rs[:, 3]), "channel sum / min / max differ"
                tot = lambda t: [int(h) * (1 << 32) + int(l) for l, h in zip(t[:, 1].tolist(), t[:, 2].tolist())]
                assert tot(fs) == tot(rs), "channel sum of squares differs"
            for k in Layer.STATE:
                if k == "mm" and not (training or observe):
                    continue
                assert torch.equal(getattr(fus, k), getattr(ref, k)), k
            if not check_bwd:
                continue
            # ---- backward
            dy = torch.randn(M, cout, generator=gd, device=DEV)
            hi_r, lo_r = (torch.empty(M, cout, dtype=torch.bfloat16, device=DEV) for _ in range(2))
            br = ref.bwd_args(M, dy, acc, hi_r, lo_r, not training)
            L().call("frost_bn_backward_reduce", C.byref(br), stream())
            L().call("frost_bn_backward_apply", C.byref(br), stream())
            hi_f, lo_f = (torch.empty(M, cout, dtype=torch.bfloat16, device=DEV) for _ in range(2))
            fb = L().PwFusedBwdArgs()
            fb.op, fb.bn = fus.operands(xp, M, ldx, x_zp), fus.bwd_args(M, dy, None, hi_f, lo_f, not training)
            L().call("frost_pw_fused_bwd_reduce", C.byref(fb), stream())
            L().call("frost_pw_fused_bwd_apply", C.byref(fb), stream())
            torch.cuda.synchronize()
            scale = float(ref.sums.abs().max())
            assert float((fus.sums - ref.sums).abs().max()) <= 1e-5 * max(scale, 1e-30), "S1/S2 differ"
            dz_r = hi_r.float() + lo_r.float()
            dz_f = hi_f.float() + lo_f.float()
            # dz = P*dv + R*I + Q in the fused kernel vs c1*(dv - a0 - a1*(I - mean)) in the chain: same value, different
            # rounding of terms of size |c1*dy|
            c1max = float((ref.A.abs() / (ref.x_scale * ref.w_scale)).max())
            tol = 1e-5 * float(dz_r.abs().max()) + 2e-6 * c1max * float(dy.abs().max()) + 1e-30
            assert float((dz_f - dz_r).abs().max()) <= tol, float((dz_f - dz_r).abs().max())
            for k in ("dgb", "dbeta", "dsf"):
                r, v = getattr(ref, k), getattr(fus, k)
                assert float((v - r).abs().max()) <= 1e-5 * float(r.abs().max()) + 1e-30, k


SHAPES = [(1, 16, 16, 0), (130, 24, 24, 0), (257, 104, 312, 0), (1000, 1728, 320, 0), (300, 56, 40, -128), (64, 320, 1280, 0),
          (513, 96, 16, 127), (40000, 16, 96, 0), (12544, 288, 1728, 0), (5000, 168, 40, 127), (777, 1440, 192, -128),
          (50176, 56, 336, 0), (200704, 24, 144, 0)]


@pytest.mark.parametrize("extra_pad", [0, 32])
@pytest.mark.parametrize("M,K,cout,zpw", SHAPES)
def test_fused_matches_first_generation_chain(M, K, cout, zpw, extra_pad):
    _run_case(M, K, cout, zpw, relu=(cout % 3 != 0), training=True, observe=True, extra_pad=extra_pad)


def test_fused_rejects_rows_tma_cannot_address():
    ly = Layer(24, 24, 0, True, 1)
    x = torch.zeros(64, 24, dtype=torch.uint8, device=DEV)
    zp = torch.zeros(1, dtype=torch.int32, device=DEV)
    f = L().PwFusedFwdArgs()
    f.op, f.bn = ly.operands(x, 64, 24, zp), ly.fin_args(64, True, True)          # dense rows of 24 bytes
    bar = torch.zeros(1, dtype=torch.int32, device=DEV)
    q = torch.zeros(64, 32, dtype=torch.uint8, device=DEV)
    f.grid_barrier, f.q, f.ldq = bar.data_ptr(), q.data_ptr(), 32
    with pytest.raises(RuntimeError, match="multiple of 16"):
        L().call("frost_pw_fused_forward", C.byref(f), stream())


@pytest.mark.parametrize("training,observe", [(False, True), (False, False), (True, False)])
def test_fused_frozen_bn_and_observer_off(training, observe):
    """eval-mode BatchNorm (running statistics; frozen in backward) and / or observer off; (False, False) is the
    single-pass mode: no statistics phase, no grid barrier."""
    _run_case(3000, 104, 312, 0, relu=True, training=training, observe=observe)
    _run_case(3000, 40, 16, 0, relu=False, training=training, observe=observe, extra_pad=16)


def test_fused_at_headline_size():
    """256 x 112 x 112 pixels, 16 -> 96: the largest 1x1 layer of FrostNet-L at bs=256."""
    _run_case(3211264, 16, 96, 0, relu=True, training=True, observe=True, steps=1)


def test_engine_fused_and_first_generation_paths_agree():
    """The whole network through the engine with fused_pw on / off: identical indices everywhere, gradients to fp32 noise."""
    import frostnet_b200 as F
    torch.manual_seed(0)
    m = F.FrostNet(nclass=16, mode="small", width_mult=0.35, quantized=True, drop_rate=0.0)
    m.train()
    m.fuse_model()
    F.prepare_qat(m)
    m.to(DEV)
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    x = torch.randn(8, 3, 96, 96, device=DEV)
    y = torch.randint(0, 16, (8,), device=DEV)
    outs = []
    for fused in (True, False):
        m.load_state_dict(sd)
        m.zero_grad()
        eng = m._frost_engine
        eng.fused_pw = fused
        eng.record_taps = True
        logits = m(x)
        torch.nn.functional.cross_entropy(logits, y).backward()
        taps = {k: v.clone() for k, v in eng.last_taps.items() if k.endswith("_q")}
        eng.record_taps = False
        outs.append((logits.detach().clone(), taps, [p.grad.clone() for p in m.parameters()],
                     {k: v.clone() for k, v in m.state_dict().items()}))
    (la, ta, ga, sa), (lb, tb, gb, sb) = outs
    for k in ta:
        assert torch.equal(ta[k], tb[k]), k
    assert torch.equal(la, lb)
    for k in sa:
        assert torch.equal(sa[k], sb[k]), k
    num = sum(float((a - b).double().pow(2).sum()) for a, b in zip(ga, gb))
    den = sum(float(b.double().pow(2).sum()) for b in gb)
    assert (num / den) ** 0.5 < 1e-4          # S1 / S2 are summed in a different order (fp32 partials of 32 vs 64 terms)


# ------------------------------------------------------------------------------------------ chained backward
CHAIN_SHAPES = [(3000, 16, 96, 0), (5000, 24, 72, 0), (4100, 24, 144, 0), (2000, 56, 168, -128), (1500, 56, 336, 0),
                (130, 64, 128, 127), (200704, 56, 336, 0), (802816, 24, 144, 0)]


def _chain_case(M, K, cout, zpw, accumulate, frozen=False):
    assert L().load().frost_pw_chain_supported(K, cout) == 1
    ly = Layer(K, cout, zpw, True, M + K)
    ldx = _r16(K)
    gd = torch.Generator(device=DEV).manual_seed(M + cout)
    x_zp = torch.tensor([7], dtype=torch.int32, device=DEV)
    x = torch.randint(0, 256, (M, ldx), generator=gd, dtype=torch.uint8, device=DEV)
    f = L().PwFusedFwdArgs()
    f.op, f.bn = ly.operands(x, M, ldx, x_zp), ly.fin_args(M, not frozen, True)
    bar = torch.zeros(1, dtype=torch.int32, device=DEV)
    q = torch.empty(M, _r16(cout), dtype=torch.uint8, device=DEV)
    f.grid_barrier, f.q, f.ldq = bar.data_ptr(), q.data_ptr(), _r16(cout)
    L().call("frost_pw_fused_forward", C.byref(f), stream())
    dy = torch.randn(M, cout, generator=gd, device=DEV)
    hi, lo = (torch.empty(M, cout, dtype=torch.bfloat16, device=DEV) for _ in range(2))
    fb = L().PwFusedBwdArgs()
    fb.op, fb.bn = ly.operands(x, M, ldx, x_zp), ly.bwd_args(M, dy, None, hi, lo, frozen)
    L().call("frost_pw_fused_bwd_reduce", C.byref(fb), stream())
    # ---- unchained: apply -> planes -> tensor-core dgrad / wgrad
    L().call("frost_pw_fused_bwd_apply", C.byref(fb), stream())
    wt = (ly.wq.float() - zpw).t().contiguous().to(torch.bfloat16)
    base = torch.randn(M, K, generator=gd, device=DEV) if accumulate else None
    dx_a = base.clone() if accumulate else torch.empty(M, K, device=DEV)
    L().call("frost_pw_dgrad_tc", hi.data_ptr(), lo.data_ptr(), wt.data_ptr(), ly.w_scale.data_ptr(), M, K, cout, dx_a.data_ptr(),
             int(accumulate), stream())
    dwq_a = torch.empty(cout, K, device=DEV)
    L().call("frost_pw_wgrad_tc", hi.data_ptr(), lo.data_ptr(), x.data_ptr(), ldx, ly.x_scale.data_ptr(), x_zp.data_ptr(), M, K, cout,
             dwq_a.data_ptr(), stream())
    grads_a = [t.clone() for t in (ly.dgb, ly.dbeta, ly.dsf)]
    for t in (ly.dgb, ly.dbeta, ly.dsf):
        t.zero_()
    # ---- chained
    ch = L().PwChainArgs()
    ch.op, ch.bn = ly.operands(x, M, ldx, x_zp), ly.bwd_args(M, dy, None, hi, lo, frozen)
    dx_b = base.clone() if accumulate else torch.full((M, K), float("nan"), device=DEV)
    dwq_b = torch.full((cout, K), float("nan"), device=DEV)
    ch.wt_bf16, ch.dx, ch.accumulate, ch.dwq = wt.data_ptr(), dx_b.data_ptr(), int(accumulate), dwq_b.data_ptr()
    L().call("frost_pw_chain_backward", C.byref(ch), stream())
    torch.cuda.synchronize()
    sx = float(dx_a.abs().max())
    assert float((dx_b - dx_a).abs().max()) <= 2e-4 * sx, ("dx", float((dx_b - dx_a).abs().max()), sx)
    sw = float(dwq_a.abs().max())
    assert float((dwq_b - dwq_a).abs().max()) <= 2e-4 * sw, ("dwq", float((dwq_b - dwq_a).abs().max()), sw)
    for a, b in zip(grads_a, (ly.dgb, ly.dbeta, ly.dsf)):
        assert torch.equal(a, b)
    # and against float64 on the planes' own values: dz = hi + lo
    dz = hi.double() + lo.double()
    ref_dx = dz @ (ly.wq.double() - zpw) * float(ly.w_scale)
    if accumulate:
        ref_dx += base.double()
    assert float((dx_b.double() - ref_dx).abs().max()) <= 1e-4 * float(ref_dx.abs().max())
    ref_dw = dz.t() @ (x[:, :K].double() - 7.0) * float(ly.x_scale)
    # fp32 accumulation in TMEM over up to 3.2 M pixels: the tensor core adds each k-step into the accumulator with
    # truncation, a bias of ~1e-7 per step that both the chained and the unchained kernel carry (they agree to 2e-4 above);
    # against float64 the budget is the north-star 1e-3 of the largest entry (measured 6e-4 at M = 3.2 M, 2e-5 at M = 2e5)
    assert float((dwq_b.double() - ref_dw).abs().max()) <= 1e-3 * float(ref_dw.abs().max())


@pytest.mark.parametrize("M,K,cout,zpw", CHAIN_SHAPES)
def test_chained_backward_matches_unchained(M, K, cout, zpw):
    _chain_case(M, K, cout, zpw, accumulate=(M % 2 == 0))


def test_chained_backward_frozen_bn():
    _chain_case(3000, 24, 144, 0, accumulate=False, frozen=True)


def test_chained_backward_at_headline_size():
    _chain_case(3211264, 16, 96, 0, accumulate=True)


def test_chain_supported_rule():
    f = L().load().frost_pw_chain_supported
    assert [f(16, 96), f(24, 72), f(24, 144), f(56, 168), f(56, 336)] == [1, 1, 1, 1, 1]     # FrostNet-L's expand convs up to 28x28
    assert [f(104, 312), f(96, 24), f(32, 16), f(16, 40), f(56, 512)] == [0, 0, 0, 0, 0]


# ------------------------------------------------------------------------------------------ QuantStub -> im2col (the stem as a GEMM)
@pytest.mark.parametrize("N,H,W", [(2, 224, 224), (3, 33, 31), (1, 8, 9)])
def test_input_quant_im2col_matches_quantstub_plus_unfold(N, H, W):
    from frostnet_b200.engine import QATEngine

    class Stem:
        kh = kw = 3
        stride, pad, ldw = 2, 1, 32
    g = torch.Generator().manual_seed(H)
    x = torch.randn(N, 3, H, W, generator=g).to(DEV)
    mk = lambda: (torch.tensor(float("inf"), device=DEV), torch.tensor(float("-inf"), device=DEV), torch.ones(1, device=DEV),
                  torch.zeros(1, dtype=torch.int32, device=DEV))
    sc = torch.zeros(L().FQ_SCRATCH_FLOATS, device=DEV)
    a, b = mk(), mk()
    xq = torch.empty(N, H, W, 3, dtype=torch.uint8, device=DEV)
    mm_a, mm_b = torch.empty(2, device=DEV), torch.empty(2, device=DEV)
    L().call("frost_input_quant", x.data_ptr(), N, 3, H, W, L().FQ(*[t.data_ptr() for t in a]), 1, 0.01, xq.data_ptr(), mm_a.data_ptr(),
             sc.data_ptr(), stream())
    Ho, Wo = (H + 2 - 3) // 2 + 1, (W + 2 - 3) // 2 + 1
    cols = torch.full((N * Ho * Wo, 32), 0xAB, dtype=torch.uint8, device=DEV)
    L().call("frost_input_quant_im2col", x.data_ptr(), N, 3, H, W, 3, 2, 1, L().FQ(*[t.data_ptr() for t in b]), 1, 0.01, cols.data_ptr(), 32,
             mm_b.data_ptr(), sc.data_ptr(), stream())
    for u, v in zip(a, b):
        assert torch.equal(u, v)                                   # same observer state, same qparams
    assert torch.equal(mm_a, mm_b)
    ref = QATEngine._im2col_from_indices(xq, Stem, int(a[3]))
    assert torch.equal(cols[:, :27], ref[:, :27])
