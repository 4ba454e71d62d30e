"""Per-module executor: a prepared ``CascadePreExBottleneck``, ``ConvBNReLU`` / ``ConvBN`` (1x1 or depthwise) or
``QuantStub`` called on its own, outside
``FrostNet.forward`` - what a caller that wires Frost bottlenecks into another network (SSDLite / ESPNetV2 style,
SURVEY.md 8f) does.

The whole-network engine keeps activations as uint8 NHWC indices between blocks.  A stand-alone module has to honour
PyTorch's calling convention instead: fp32 NCHW tensors in and out.  The bridge is the tensor's quantisation grid:
every tensor produced by a frostnet_b200 module carries ``_frost_qparams = (scale, zero_point)`` (the producer's live
state_dict buffers); the consuming block re-derives the exact uint8 indices from them (a value on the grid survives the
round trip exactly), runs the same kernels as the whole-network engine (``QATEngine._block_forward`` /
``_block_backward``), and returns the dequantised result with its own qparams attached.  A tensor without qparams - the
output of a foreign op - must go through a ``QuantStub`` first, exactly like in the reference.
"""
import ctypes as C

import torch

from . import _lib as L
from . import qat as Q
from .engine import QATEngine, _Layer, _QT, _fq_struct


def attach_qparams(t, scale, zero_point):
    t._frost_qparams = (scale, zero_point)
    return t


def qparams_of(t, who):
    qp = getattr(t, "_frost_qparams", None)
    if qp is None:
        raise RuntimeError(
            "frostnet_b200: %s got a tensor without quantisation parameters.  A prepared Frost module computes on uint8 "
            "indices: feed it the output of another frostnet_b200 module (QuantStub, bottleneck), which carries its "
            "(scale, zero_point) as `_frost_qparams`; tensors produced by other ops must pass a QuantStub first." % who)
    return qp


class BlockEngine(QATEngine):
    """QATEngine over ONE bottleneck: same layer artefacts, same kernels, fp32 NCHW at the boundary."""

    def _discover(self, m, add):
        """`model` is a CascadePreExBottleneck, or a ConvBNReLU / ConvBN wrapper around one fused 1x1 / depthwise conv."""
        from . import frostnet as FN
        self.stem = None
        self._single = None
        if isinstance(m, FN.CascadePreExBottleneck):
            self._add_block(add, "", m)
            return
        if isinstance(m, Q.FrostConvBn2d):                   # a fused conv called directly (it sits in a foreign container)
            seq, mod = "self", m
        else:
            seq = getattr(m, "_seq_name", "conv")           # "conv" (frostnet.py) or "cbr" / "cb" (mobilenetv3.py)
            mod = getattr(m, seq)[0]
        if not isinstance(mod, Q.FrostConvBn2d):
            raise RuntimeError("frostnet_b200: %s is not fused; call fuse_model() + prepare_qat" % type(m).__name__)
        if tuple(mod.dilation) != (1, 1) and not mod.is_depthwise:
            raise RuntimeError("frostnet_b200: stand-alone %s: only depthwise convolutions may be dilated" % type(m).__name__)
        kk = mod.kernel_size[0] * mod.kernel_size[1] * mod.in_channels
        if mod.is_depthwise:
            kind = "dw"
        elif mod.groups == 1 and tuple(mod.kernel_size) == (1, 1) and tuple(mod.stride) == (1, 1):
            kind = "pw"
        elif mod.groups == 1 and mod.kernel_size[0] == mod.kernel_size[1] and kk <= 32 and mod.out_channels <= 32:
            kind = "stem"                      # dense kxk on an image-like input (MobileNetV3's conv1: 3 -> 16, 3x3 / 2)
        else:
            raise RuntimeError("frostnet_b200: stand-alone %s: only 1x1, depthwise 3x3 / 5x5 and stem-sized dense kxk "
                               "(k*k*cin <= 32, cout <= 32) convolutions have kernels" % type(m).__name__)
        self._single = _Layer(seq + ".0", mod, kind)
        if kind == "stem":
            # stand-alone stems take the direct-convolution kernels (frost_stem_conv_forward / frost_stem_wgrad); the im2col
            # tensor-core route needs the QuantStub fused in front, which only the whole-network engine has
            self._single.allow_im2col = False
        self.layers.append(self._single)

    def __deepcopy__(self, memo):
        import copy
        return BlockEngine(copy.deepcopy(self.model, memo))

    # ------------------------------------------------------------------ forward / backward
    def _forward(self, x, qp, save):
        self._refresh_flags()
        m, dev, st = self.model, self.dev, L.stream(self.dev)
        self.generation += 1
        if x.device != dev:
            raise RuntimeError("frostnet_b200: input on %s, module on %s" % (x.device, dev))
        x = x.detach()
        if x.dtype != torch.float32 or not x.is_contiguous():
            x = x.float().contiguous()
        N, Cin, H, W = x.shape
        cin_expected = self._single.cin if self._single is not None else m.in_channels
        if Cin != cin_expected:
            raise RuntimeError("%s expects %d input channels, got %d" % (type(m).__name__, cin_expected, Cin))
        saved = {"gen": self.generation} if save else None
        if self.record_taps:
            self.last_taps = {}
        L.call("frost_weight_prep_multi", self._wdesc_dev[0].data_ptr(), len(self.layers), self._wchunks.data_ptr(),
               self._n_wchunks, self._wscratch.data_ptr(), st)
        L.call("frost_stats_reset", self.stats.data_ptr(), self.n_stat_chan, st)
        self.barriers.zero_()
        # fp32 NCHW on the producer's grid -> uint8 NHWC indices (no observer: the grid is the producer's)
        scale, zp = qp
        xq = torch.empty((N, H, W, Cin), dtype=torch.uint8, device=dev)
        mm_in = torch.empty(2, dtype=torch.float32, device=dev)
        fq = L.FQ(self._dummy_mm.data_ptr(), self._dummy_mm.data_ptr() + 4, scale.data_ptr(), zp.data_ptr())
        L.call("frost_input_quant", x.data_ptr(), N, Cin, H, W, fq, 0, Q.AVERAGING_CONSTANT, xq.data_ptr(), mm_in.data_ptr(),
               self.scratch.data_ptr(), st)
        is_stem = self._single is not None and self._single.kind == "stem"
        q, ld = (xq.view(N * H * W, Cin), Cin) if is_stem else self._alloc_q(N * H * W, Cin)    # the stem kernels read dense NHWC
        if ld != Cin:
            q.zero_()
            q[:, :Cin].copy_(xq.view(N * H * W, Cin))      # rows padded to TMA's 16-byte pitch (boundary plumbing)
        else:
            q = xq.view(N * H * W, Cin)
        t = _QT(q, N, H, W, Cin, scale, zp, mm_in, ld)
        if self._single is not None:
            o = self._conv_bn(self._single, t, m.training, st, saved)
        else:
            o = self._block_forward(self.blocks[0], t, m.training, st, saved)
        y = torch.empty((o.N, o.C, o.H, o.W), dtype=torch.float32, device=dev)
        L.call("frost_dequant_to_nchw", o.q.data_ptr(), o.ld, o.scale.data_ptr(), o.zp.data_ptr(), o.N, o.H, o.W, o.C,
               y.data_ptr(), st)
        if saved is not None:
            saved["out"] = o
            # STE of the boundary quantisation: values whose index leaves [0, 255] pass no gradient (none do when the
            # input really is on the grid)
            idx = torch.round(x * (1.0 / scale)) + zp
            saved["in_mask"] = (idx >= 0) & (idx <= 255)
        return (y, o), saved

    def _backward(self, saved, dy):
        if saved.get("gen") != self.generation:
            raise RuntimeError("frostnet_b200: backward through a bottleneck forward that is no longer the latest one of this "
                               "module (every forward overwrites the per-layer state its backward reads)")
        dev, st = self.dev, L.stream(self.dev)
        which = self._pick_gflat()
        gflat = self.gflat[which]
        gbase = gflat.data_ptr()
        o = saved["out"]
        dy = dy.contiguous().float()
        g = torch.empty((o.M, o.C), dtype=torch.float32, device=dev)
        L.call("frost_nchw_to_nhwc", dy.data_ptr(), o.N, o.C, o.H, o.W, g.data_ptr(), 0, st)
        if self._single is not None and self._single.kind == "stem":
            # the reference's stem sees the image: no gradient flows to it (run() refuses inputs that require one)
            self._conv_bn_bwd(self._single, g, saved, gbase, None, False, st)
            L.call("frost_weight_backward_multi", self._wdesc_dev[which].data_ptr(), len(self.layers),
                   self._wbchunks.data_ptr(), self._n_wbchunks, st)
            return None, [gflat.narrow(0, self.param_off[id(p)], p.numel()).view(p.shape) for p in self.params]
        if self._single is not None:
            xin = saved[self._single.name][0]
            gx = torch.empty((xin.M, xin.C), dtype=torch.float32, device=dev)
            self._conv_bn_bwd(self._single, g, saved, gbase, gx, False, st)
        else:
            xin = saved[".in"]
            gx = self._block_backward(self.blocks[0], g, saved, gbase, st)
        L.call("frost_weight_backward_multi", self._wdesc_dev[which].data_ptr(), len(self.layers),
               self._wbchunks.data_ptr(), self._n_wbchunks, st)
        dx = gx.view(xin.N, xin.H, xin.W, xin.C).permute(0, 3, 1, 2).contiguous() * saved["in_mask"]
        grads = [gflat.narrow(0, self.param_off[id(p)], p.numel()).view(p.shape) for p in self.params]
        return dx, grads

    def forward(self, x, qp, save):
        self._ensure_built(check=False)
        with torch.cuda.device(self.dev):
            return self._forward(x, qp, save)

    # ------------------------------------------------------------------ entry point
    def run(self, x):
        qp = qparams_of(x, "a prepared bottleneck")
        if not x.is_cuda:
            raise RuntimeError("frostnet_b200: the QAT path runs on a CUDA device (B200) only; got a CPU tensor")
        self._ensure_built()
        if not hasattr(self, "_dummy_mm") or self._dummy_mm.device != self.dev:
            self._dummy_mm = torch.zeros(2, dtype=torch.float32, device=self.dev)
        need_grad = torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in self.params))
        if need_grad and x.requires_grad and self._single is not None and self._single.kind == "stem":
            raise RuntimeError("frostnet_b200: a stand-alone stem convolution has no input gradient (the reference feeds it "
                               "the image); detach the input")
        if need_grad:
            y = _BlockFunction.apply(self, x, qp[0], qp[1], *self.params)
        else:
            (y, o), _ = self.forward(x, qp, save=False)
            self._last_out = o
        o = self._last_out
        return attach_qparams(y, o.scale, o.zp)


class _BlockFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, engine, x, scale, zero_point, *params):
        (y, o), saved = engine.forward(x, (scale, zero_point), save=True)
        engine._last_out = o
        ctx.engine, ctx.saved = engine, saved
        return y

    @staticmethod
    def backward(ctx, dy):
        dx, grads = ctx.engine.backward(ctx.saved, dy)
        ctx.saved = None
        return (None, dx, None, None) + tuple(grads)


def run_block(block, x):
    eng = block.__dict__.get("_frost_block_engine")
    if eng is None:
        eng = BlockEngine(block)
        block.__dict__["_frost_block_engine"] = eng
    return eng.run(x)


# ---------------------------------------------------------------------- stand-alone QuantStub
class _FakeQuantFunction(torch.autograd.Function):
    """fake_quantize.py:423-438 on an fp32 tensor: observer (if on) + quantise-dequantise; backward = STE mask."""

    @staticmethod
    def forward(ctx, x, fq, scratch):
        x = x.detach()
        if x.dtype != torch.float32 or not x.is_contiguous():
            x = x.float().contiguous()
        y = torch.empty_like(x)
        mask = torch.empty(x.shape, dtype=torch.uint8, device=x.device)
        with torch.cuda.device(x.device):
            L.call("frost_fq_forward", x.data_ptr(), x.numel(), _fq_struct(fq), Q.ACT_QMIN, Q.ACT_QMAX, 0,
                   1 if fq._observe else 0, Q.AVERAGING_CONSTANT, y.data_ptr(), mask.data_ptr(), None, scratch.data_ptr(),
                   L.stream(x.device))
        ctx.mask = mask
        return y

    @staticmethod
    def backward(ctx, dy):
        dy = dy.contiguous().float()
        dx = torch.empty_like(dy)
        with torch.cuda.device(dy.device):
            L.call("frost_fq_backward", dy.data_ptr(), ctx.mask.data_ptr(), dy.numel(), dx.data_ptr(), L.stream(dy.device))
        return dx, None, None


def run_quant_stub(stub, x):
    fq = stub.activation_post_process
    if not x.is_cuda:
        raise RuntimeError("frostnet_b200: the QAT path runs on a CUDA device (B200) only; got a CPU tensor")
    sc = stub.__dict__.get("_frost_scratch")
    if sc is None or sc.device != x.device:
        sc = torch.zeros(L.FQ_SCRATCH_FLOATS, dtype=torch.float32, device=x.device)
        stub.__dict__["_frost_scratch"] = sc
    y = _FakeQuantFunction.apply(x, fq, sc)
    return attach_qparams(y, fq.scale, fq.zero_point)


# ---------------------------------------------------------------------- stand-alone FloatFunctional.add / .cat
def _to_nhwc_u8(x, qp, scratch, dummy):
    """fp32 NCHW on the grid qp -> (uint8 NHWC indices, dequantised [min, max] of the tensor)"""
    scale, zp = qp
    x = x.detach()
    if x.dtype != torch.float32 or not x.is_contiguous():
        x = x.float().contiguous()
    N, Cc, H, W = x.shape
    q = torch.empty((N, H, W, Cc), dtype=torch.uint8, device=x.device)
    mm = torch.empty(2, dtype=torch.float32, device=x.device)
    fq = L.FQ(dummy.data_ptr(), dummy.data_ptr() + 4, scale.data_ptr(), zp.data_ptr())
    L.call("frost_input_quant", x.data_ptr(), N, Cc, H, W, fq, 0, Q.AVERAGING_CONSTANT, q.data_ptr(), mm.data_ptr(),
           scratch.data_ptr(), L.stream(x.device))
    return q, mm


def _qt(q, qp, mm):
    N, H, W, Cc = q.shape
    return L.QTensor(q.data_ptr(), qp[0].data_ptr(), qp[1].data_ptr(), mm.data_ptr(), Cc, Cc)


def _functional_buffers(ff, dev):
    b = ff.__dict__.get("_frost_buffers")
    if b is None or b[0].device != dev:
        b = (torch.zeros(L.FQ_SCRATCH_FLOATS, dtype=torch.float32, device=dev), torch.zeros(2, dtype=torch.float32, device=dev))
        ff.__dict__["_frost_buffers"] = b
    return b


class _AddFunction(torch.autograd.Function):
    """FloatFunctional.add (functional_modules.py:50-52): torch.add + the module's own observer / fake-quant."""

    @staticmethod
    def forward(ctx, x, y, xs, xz, ys, yz, ff):
        dev = x.device
        fq = ff.activation_post_process
        scratch, dummy = _functional_buffers(ff, dev)
        with torch.cuda.device(dev):
            qa, ma = _to_nhwc_u8(x, (xs, xz), scratch, dummy)
            qb, mb = _to_nhwc_u8(y, (ys, yz), scratch, dummy)
            N, H, W, Cc = qa.shape
            qo = torch.empty_like(qa)
            mo = torch.empty(2, dtype=torch.float32, device=dev)
            ta, tb = _qt(qa, (xs, xz), ma), _qt(qb, (ys, yz), mb)
            L.call("frost_add_forward", ta, tb, qa.numel(), _fq_struct(fq), 1 if fq._observe else 0, Q.AVERAGING_CONSTANT,
                   qo.data_ptr(), Cc, mo.data_ptr(), scratch.data_ptr(), L.stream(dev))
            out = torch.empty((N, Cc, H, W), dtype=torch.float32, device=dev)
            L.call("frost_dequant_to_nchw", qo.data_ptr(), Cc, fq.scale.data_ptr(), fq.zero_point.data_ptr(), N, H, W, Cc,
                   out.data_ptr(), L.stream(dev))
        ctx.saved = (qa, ma, qb, mb, (xs, xz), (ys, yz), fq)
        return out

    @staticmethod
    def backward(ctx, dout):
        qa, ma, qb, mb, qpa, qpb, fq = ctx.saved
        dev = dout.device
        N, H, W, Cc = qa.shape
        with torch.cuda.device(dev):
            st = L.stream(dev)
            g = torch.empty((N * H * W, Cc), dtype=torch.float32, device=dev)
            L.call("frost_nchw_to_nhwc", dout.contiguous().float().data_ptr(), N, Cc, H, W, g.data_ptr(), 0, st)
            dsum = torch.empty_like(g)
            da = torch.empty_like(g)
            L.call("frost_add_backward", g.data_ptr(), _qt(qa, qpa, ma), _qt(qb, qpb, mb), g.numel(), fq.scale.data_ptr(),
                   fq.zero_point.data_ptr(), dsum.data_ptr(), da.data_ptr(), 0, st)
        dx = da.view(N, H, W, Cc).permute(0, 3, 1, 2).contiguous()
        return dx, dx.clone(), None, None, None, None, None


def run_functional_add(ff, x, y):
    xs, xz = qparams_of(x, "a prepared FloatFunctional.add")
    ys, yz = qparams_of(y, "a prepared FloatFunctional.add")
    if x.shape != y.shape or x.dim() != 4:
        raise RuntimeError("frostnet_b200: stand-alone FloatFunctional.add expects two NCHW tensors of the same shape")
    out = _AddFunction.apply(x, y, xs, xz, ys, yz, ff)
    fq = ff.activation_post_process
    return attach_qparams(out, fq.scale, fq.zero_point)


class _CatFunction(torch.autograd.Function):
    """FloatFunctional.cat along channels (functional_modules.py:80-82): torch.cat + observer / fake-quant."""

    @staticmethod
    def forward(ctx, x, y, xs, xz, ys, yz, ff):
        dev = x.device
        fq = ff.activation_post_process
        scratch, dummy = _functional_buffers(ff, dev)
        with torch.cuda.device(dev):
            qa, ma = _to_nhwc_u8(x, (xs, xz), scratch, dummy)
            qb, mb = _to_nhwc_u8(y, (ys, yz), scratch, dummy)
            N, H, W, Ca = qa.shape
            Cb = qb.shape[3]
            qo = torch.empty((N, H, W, Ca + Cb), dtype=torch.uint8, device=dev)
            mo = torch.empty(2, dtype=torch.float32, device=dev)
            L.call("frost_cat_forward", _qt(qa, (xs, xz), ma), _qt(qb, (ys, yz), mb), N * H * W, _fq_struct(fq),
                   1 if fq._observe else 0, Q.AVERAGING_CONSTANT, qo.data_ptr(), Ca + Cb, mo.data_ptr(), L.stream(dev))
            out = torch.empty((N, Ca + Cb, H, W), dtype=torch.float32, device=dev)
            L.call("frost_dequant_to_nchw", qo.data_ptr(), Ca + Cb, fq.scale.data_ptr(), fq.zero_point.data_ptr(), N, H, W, Ca + Cb,
                   out.data_ptr(), L.stream(dev))
        ctx.saved = (qa, ma, qb, mb, (xs, xz), (ys, yz), fq)
        return out

    @staticmethod
    def backward(ctx, dout):
        qa, ma, qb, mb, qpa, qpb, fq = ctx.saved
        dev = dout.device
        N, H, W, Ca = qa.shape
        Cb = qb.shape[3]
        with torch.cuda.device(dev):
            st = L.stream(dev)
            g = torch.empty((N * H * W, Ca + Cb), dtype=torch.float32, device=dev)
            L.call("frost_nchw_to_nhwc", dout.contiguous().float().data_ptr(), N, Ca + Cb, H, W, g.data_ptr(), 0, st)
            da = torch.empty((N * H * W, Ca), dtype=torch.float32, device=dev)
            db = torch.empty((N * H * W, Cb), dtype=torch.float32, device=dev)
            L.call("frost_cat_backward", g.data_ptr(), _qt(qa, qpa, ma), _qt(qb, qpb, mb), N * H * W, fq.scale.data_ptr(),
                   fq.zero_point.data_ptr(), da.data_ptr(), db.data_ptr(), 0, st)
        dx = da.view(N, H, W, Ca).permute(0, 3, 1, 2).contiguous()
        dy = db.view(N, H, W, Cb).permute(0, 3, 1, 2).contiguous()
        return dx, dy, None, None, None, None, None


def run_functional_cat(ff, xs, dim):
    if len(xs) != 2 or dim != 1 or xs[0].dim() != 4:
        raise RuntimeError("frostnet_b200: stand-alone FloatFunctional.cat supports two NCHW tensors along dim 1")
    x, y = xs
    a = qparams_of(x, "a prepared FloatFunctional.cat")
    b = qparams_of(y, "a prepared FloatFunctional.cat")
    out = _CatFunction.apply(x, y, a[0], a[1], b[0], b[1], ff)
    fq = ff.activation_post_process
    return attach_qparams(out, fq.scale, fq.zero_point)
