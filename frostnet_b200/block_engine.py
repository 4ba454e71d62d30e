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
        mod = m.conv[0]
        if not isinstance(mod, Q.FrostConvBn2d):
            raise RuntimeError("frostnet_b200: %s is not fused; call fuse_model() + prepare_qat" % type(m).__name__)
        if mod.is_depthwise:
            kind = "dw"
        elif mod.groups == 1 and tuple(mod.kernel_size) == (1, 1) and tuple(mod.stride) == (1, 1):
            kind = "pw"
        else:
            raise RuntimeError("frostnet_b200: stand-alone %s: only 1x1 and depthwise 3x3 / 5x5 convolutions have kernels "
                               "(dense kxk exists for the 3-channel stem inside FrostNet only)" % type(m).__name__)
        self._single = _Layer("conv.0", mod, kind)
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
        q, ld = self._alloc_q(N * H * W, Cin)
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
