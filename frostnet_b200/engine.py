"""Whole-network QAT executor: FrostNet.forward / backward as a plan of sm_100a kernel launches.

One ``QATEngine`` per prepared model.  ``run(x)`` is what ``FrostNet.forward`` calls after
``prepare_qat``: the whole network is ONE autograd node (``_QATFunction``) whose forward walks
frostnet.py:318-332 / :124-145 and whose backward walks it in reverse (SURVEY.md 8a'), issuing the
C-ABI kernels of include/frost_b200.h on the current CUDA stream.  Module parameters and buffers
(BN running stats, observer min/max, scale, zero_point) are read and updated IN PLACE on the
device - no host synchronisation anywhere in a step.

HBM layout: activations are NHWC uint8 indices (1 B/element), raw conv outputs NHWC int32, the
gradients NHWC fp32; weights are re-quantised to int8 once per step by one multi-tensor launch.
"""
import ctypes as C
import os

import torch

from . import _lib as L
from . import qat as Q


class _QT:
    """A quantised activation tensor: uint8 NHWC indices + pointers to its (scale, zero_point)
    buffers and to the dequantised [min, max] of its current contents."""
    __slots__ = ("q", "N", "H", "W", "C", "scale", "zp", "mm", "ld")

    def __init__(self, q, N, H, W, Cc, scale, zp, mm, ld=None):
        self.q, self.N, self.H, self.W, self.C, self.scale, self.zp, self.mm = q, N, H, W, Cc, scale, zp, mm
        self.ld = Cc if ld is None else ld          # row pitch in bytes (== C: dense NHWC)

    @property
    def M(self):
        return self.N * self.H * self.W

    def c(self):
        return L.QTensor(self.q.data_ptr(), self.scale.data_ptr(), self.zp.data_ptr(), self.mm.data_ptr(), self.C, self.ld)


class _RawInput:
    """fp32 NCHW image fed straight to the stem (feature backbone: no QuantStub); scale == 1."""
    __slots__ = ("x", "N", "H", "W", "C", "scale")

    def __init__(self, x, scale_one):
        self.x, self.scale = x, scale_one
        self.N, self.C, self.H, self.W = x.shape


def _fq_struct(f):
    return L.FQ(f.activation_post_process.min_val.data_ptr(), f.activation_post_process.max_val.data_ptr(),
                f.scale.data_ptr(), f.zero_point.data_ptr())


class _Layer:
    """One fused conv (FrostConvBn2d) or the classifier conv, with its persistent device artefacts."""

    def __init__(self, name, mod, kind):
        self.name, self.mod, self.kind = name, mod, kind       # kind: stem | pw | dw | cls
        w = mod.weight
        self.cout, self.cin_g, self.kh, self.kw = w.shape
        self.has_bn = kind != "cls"
        self.relu = bool(getattr(mod, "relu", False))
        if kind != "cls":
            self.stride, self.pad = mod.stride[0], mod.padding[0]
            self.cin = mod.in_channels
            self.dil = mod.dilation[0]
            # depthwise convs that are dilated or not "same"-padded take the gather kernels (csrc/dw_dilated.cu)
            self.gather = kind == "dw" and (self.dil > 1 or self.pad != (self.kh - 1) // 2)
        else:
            self.stride, self.pad, self.cin, self.dil, self.gather = 1, 0, mod.in_channels, 1, False
        self.layout = {"pw": 0, "cls": 0, "dw": 1, "stem": 2}[kind]
        self.allow_im2col = True               # stand-alone stems (block_engine.py) switch the im2col GEMM route off

    def alloc(self, dev):
        f32 = dict(dtype=torch.float32, device=dev)
        n = self.cout * self.cin_g * self.kh * self.kw
        if self.kind == "cls":
            # the classifier GEMMs read 4 output channels at a time: nclass that is not a multiple of 4 (CIFAR-10)
            # runs on a zero-padded copy of the weight-index / gradient rows
            self.cout_p = (self.cout + 3) // 4 * 4
            n = self.cout_p * self.cin_g * self.kh * self.kw
        self.wq = torch.zeros(n, dtype=torch.int8, device=dev)
        self.wt_bf16 = torch.zeros(n, dtype=torch.bfloat16, device=dev) if self.kind == "pw" else None
        # tensor-core operand bytes of the fused 1x1 kernels: rows padded to a multiple of 16 bytes (TMA's stride rule).
        # The dense kxk stem joins them as an im2col GEMM (K = kh*kw*cin padded to 32) when its patch fits 32 bytes.
        kk = self.cin_g * self.kh * self.kw
        self.im2col = self.kind == "stem" and kk <= 32 and self.allow_im2col
        self.k_mma = self.cin_g if self.kind == "pw" else ((kk + 15) // 16 * 16 if self.im2col else 0)
        self.ldw = (self.k_mma + 15) // 16 * 16
        self.wq_mma = torch.zeros(self.cout * self.ldw, dtype=torch.int8, device=dev) if self.k_mma else None
        if self.im2col:
            n = max(n, self.cout * self.ldw)        # dwq of the im2col path is [cout][ldw]
        self.wmask = torch.zeros(n, dtype=torch.uint8, device=dev)
        self.sf = torch.ones(self.cout, **f32)
        self.rstd = torch.ones(self.cout, **f32)
        self.wsum = torch.zeros(self.cout, dtype=torch.int32, device=dev)
        self.dwq = torch.zeros(n, **f32)
        self.dgamma_bn = torch.zeros(self.cout, **f32)
        self.dsf_bn = torch.zeros(self.cout, **f32)
        self.A = torch.zeros(self.cout, **f32)
        self.B = torch.zeros(self.cout, **f32)
        self.mean_I = torch.zeros(self.cout, **f32)
        self.kfac = torch.zeros(self.cout, **f32)
        self.coef = torch.zeros(3 * self.cout, **f32)
        self.sums = torch.zeros(2 * self.cout, dtype=torch.float64, device=dev)
        self.mm = torch.zeros(2, **f32)


class QATEngine:
    def __init__(self, model):
        self.model = model
        self._built = False
        self.last_taps = None
        self.last_grad_taps = {}
        self.record_taps = False
        self.dropout_mask = None      # tests may inject a keep mask [N,1280]
        # Teacher forcing for parity tests: name -> uint8 NHWC indices that REPLACE the computed output of that
        # node after it has been compared (force_report[name] = (mismatch rate, max |delta index|)).  A quantised
        # network amplifies a single rounding flip layer by layer, so implementations that sum in a different
        # order can only be compared layer by layer from identical inputs.
        self.force = None
        self.force_report = {}
        self.grad_sync = None         # callable(flat_grad) -> None, set by frostnet_b200.parallel
        # every forward overwrites the per-layer persistent device state the backward reads (quantised weights,
        # BN affine / saved statistics, live scale / zero_point buffers): a graph is only valid for backward while
        # no later forward has run.  `generation` stamps the graphs (see backward()).
        self.generation = 0
        self._fingerprint = None
        # 1x1 ConvBn(ReLU): one fused tensor-core launch per pass, int32 accumulators never stored (csrc/pw_fused.cu).
        # False selects the first-generation chain conv -> bn_finalize -> bnq_apply / bn_backward_* (kept as the cross-check).
        self.fused_pw = True
        self.chain_bwd = True       # expand convs: BN-backward apply + dgrad + wgrad in one kernel (csrc/pw_chain.cu)
        # Weight gradients on a side stream next to the data-gradient chain.  Measured on B200 (bs=256): no gain - every
        # tensor-core kernel takes a whole SM per CTA (shared memory) and even the small layers launch 148 CTAs, so the
        # two streams take turns anyway.  Off by default; FROST_OVERLAP_WGRAD=1 re-enables it for measurements.
        self.overlap_wgrad = os.environ.get("FROST_OVERLAP_WGRAD", "0") == "1"
        self._side = None

    def invalidate(self):
        self._built = False

    def _side_stream(self):
        if self._side is None:
            self._side = torch.cuda.Stream(self.dev)
        return self._side

    # A copy (copy.deepcopy(model) for an EMA model, torch.save(model)) must not inherit raw device pointers into
    # the ORIGINAL model's tensors: it gets a fresh, unbuilt engine bound to the copied module tree.
    def __deepcopy__(self, memo):
        import copy
        new = type(self)(copy.deepcopy(self.model, memo))
        new.grad_sync = self.grad_sync
        return new

    def __getstate__(self):
        return {"model": self.model}

    def __setstate__(self, state):
        self.__init__(state["model"])

    def _tensor_fingerprint(self, full=True):
        """data_ptr of every parameter (cheap: ~200 tensors) and, when `full`, of every buffer (~1500)."""
        m = self.model
        fp = tuple(t.data_ptr() for t in m.parameters())
        return fp + tuple(t.data_ptr() for t in m.buffers()) if full else fp

    def _alloc_q(self, M, Cc):
        """uint8 index tensor [M][ld]: rows padded to a multiple of 16 bytes (TMA's stride rule; the padding is
        don't-care) on the fused path, dense for the first-generation kernels."""
        ld = (Cc + 15) // 16 * 16 if self.fused_pw else Cc
        return torch.empty((M, ld), dtype=torch.uint8, device=self.dev), ld

    @staticmethod
    def _im2col_from_indices(xq, ly, zp):
        """[N,H,W,C] uint8 indices -> the rows frost_input_quant_im2col produces (torch ops; parity tests only)."""
        N, H, W, Cc = xq.shape
        x4 = torch.nn.functional.pad(xq.permute(0, 3, 1, 2).float(), (ly.pad, ly.pad, ly.pad, ly.pad), value=float(zp))
        u = torch.nn.functional.unfold(x4, (ly.kh, ly.kw), stride=ly.stride)            # [N, C*kh*kw, L], (c, kh, kw) order
        L_ = u.shape[-1]
        u = u.reshape(N, Cc, ly.kh * ly.kw, L_).permute(0, 3, 2, 1).reshape(N * L_, ly.kh * ly.kw * Cc)   # (kh, kw, c)
        cols = torch.zeros((N * L_, ly.ldw), dtype=torch.uint8, device=xq.device)
        cols[:, :u.shape[1]] = u.to(torch.uint8)
        return cols

    def _maybe_force(self, name, q, Cc=None):
        if self.force is None or name not in self.force:
            return
        if Cc is not None:
            q = q[:, :Cc]                              # the logical [M][C] part of a row-padded tensor
        ref = self.force[name].to(q.device).reshape(q.shape)
        d = (q.int() - ref.int()).abs()
        self.force_report[name] = (float((d > 0).float().mean()), int(d.max()))
        q.copy_(ref)

    # ------------------------------------------------------------------ build
    def _ensure_built(self, check=True):
        if self._built and not check:
            return
        if self._built:
            # p.data = ..., load_state_dict(assign=True), vector_to_parameters ... move tensors without going
            # through Module._apply: the cached pointer tables would silently address freed memory
            # parameters every forward, buffers every 64th: the check sits on the critical path of a step's first launch
            self._fp_tick = (getattr(self, "_fp_tick", 0) + 1) % 64
            if self._fp_tick == 0:
                ok = self._tensor_fingerprint(True) == self._fingerprint
            else:
                ok = tuple(p.data_ptr() for p in self.params) == self._fingerprint[:len(self.params)]
            if ok:
                return
            self._built = False
        m = self.model
        L.load()
        self.features = not hasattr(m, "classifier")     # feature backbone (frostnet_features.py): no head, 4 taps out
        named = dict(m.named_modules())
        self.layers, self.blocks = [], []

        def add(name, kind):
            mod = named[name + ".conv.0"] if kind != "cls" else named[name]
            if kind != "cls" and not isinstance(mod, Q.FrostConvBn2d):
                raise RuntimeError("frostnet_b200: %s is not fused; call model.fuse_model() + prepare_qat" % name)
            ly = _Layer(name + (".conv.0" if kind != "cls" else ""), mod, kind)
            self.layers.append(ly)
            return ly

        self._discover(m, add)
        p0 = next(m.parameters())
        L.require_cuda(p0, "model")
        dev = p0.device
        self.dev = dev
        self._side = None
        for ly in self.layers:
            if ly.kind == "dw" and not (ly.kh in (3, 5) and ly.stride in (1, 2) and ly.cin_g == 1
                                        and 0 <= ly.pad <= ly.dil * (ly.kh - 1) and 1 <= ly.dil <= 16):
                raise RuntimeError("frostnet_b200: unsupported depthwise conv %s" % ly.name)
            if ly.kind != "dw" and ly.dil != 1:
                raise RuntimeError("frostnet_b200: %s: only depthwise convolutions may be dilated" % ly.name)
            ly.alloc(dev)
        # parameter order == model.named_parameters() (the optimizer / checkpoint contract)
        self.params = [p for _, p in m.named_parameters()]
        self.param_off, off = {}, 0
        for p in self.params:
            self.param_off[id(p)] = off
            off += p.numel()
        self.n_param_elems = off
        self.gflat = [torch.zeros(off, dtype=torch.float32, device=dev) for _ in range(2)]
        # channel statistics arena
        tot_c = sum(ly.cout for ly in self.layers if ly.has_bn)
        self.stats = torch.zeros(tot_c * C.sizeof(L.ChanStats), dtype=torch.uint8, device=dev)
        o = 0
        for ly in self.layers:
            if ly.has_bn:
                ly.stats_ptr = self.stats.data_ptr() + o * C.sizeof(L.ChanStats)
                o += ly.cout
        self.n_stat_chan = tot_c
        # one grid-barrier word per layer, zeroed with the statistics at the start of every forward
        self.barriers = torch.zeros(len(self.layers), dtype=torch.int32, device=dev)
        for i, ly in enumerate(self.layers):
            ly.barrier_ptr = self.barriers.data_ptr() + 4 * i
        self.scratch = torch.zeros(L.FQ_SCRATCH_FLOATS, dtype=torch.float32, device=dev)
        self.one = torch.ones(1, dtype=torch.float32, device=dev)
        self._wdesc_dev = [self._build_wdesc(i) for i in range(2)]
        self._flag_epoch = Q.FrostFakeQuantize.flag_epoch
        self._desc_fused = self.fused_pw
        self._wchunks, self._n_wchunks = L.chunk_table([ly.wq.numel() for ly in self.layers], L.WEIGHT_CHUNK, dev)
        self._wbchunks, self._n_wbchunks = L.chunk_table([ly.cout for ly in self.layers], L.WEIGHT_BWD_CHANNELS, dev)
        self._wscratch = torch.tensor([float("inf"), float("-inf")] * len(self.layers), dtype=torch.float32, device=dev)
        self._fingerprint = self._tensor_fingerprint()
        self._built = True

    def _add_block(self, add, p, blk, stage_end=False, stage=0):
        """One CascadePreExBottleneck (frostnet.py:81-145) -> its fused convs in execution order; `p`: module-name prefix."""
        dot = p + "." if p else ""
        b = dict(name=p, mod=blk, squeeze=None, conv1=None, stage_end=stage_end, stage=stage)
        if blk.expand_ratio != 1:
            if blk.block_type == "CAS":
                b["squeeze"] = add(dot + "squeeze_conv", "pw")
            b["conv1"] = add(dot + "conv1", "pw")
        b["conv2"] = add(dot + "conv2", "dw")
        b["reduce"] = add(dot + "reduce_conv", "pw")
        b["skip"] = not blk.reduction
        self.blocks.append(b)
        return b

    def _discover(self, m, add):
        """Whole network: stem, every bottleneck, last_layer + classifier (frostnet.py:318-332)."""
        self.stem = add("conv1", "stem")
        for si, stage in enumerate(m.stages()):
            for bi, blk in enumerate(stage):
                self._add_block(add, "layer%d.%d" % (si + 1, bi), blk, stage_end=(bi == len(stage) - 1), stage=si)
        if not self.features:
            self.last = add("last_layer", "pw")
            self.cls = add("classifier.2", "cls")

    def _build_wdesc(self, which):
        arr = (L.WeightDesc * len(self.layers))()
        g = self.gflat[which]
        for i, ly in enumerate(self.layers):
            d, mod = arr[i], ly.mod
            d.weight = mod.weight.data_ptr()
            if ly.has_bn:
                d.bn_weight = mod.bn.weight.data_ptr()
                d.bn_var = mod.bn.running_var.data_ptr()
                d.bn_eps = mod.bn.eps
            else:
                d.bn_weight, d.bn_var, d.bn_eps = None, None, 0.0
            d.cout, d.cin_g, d.kh, d.kw, d.layout = ly.cout, ly.cin_g, ly.kh, ly.kw, ly.layout
            d.observe = 1 if mod.weight_fake_quant._observe else 0
            d.averaging_const = Q.AVERAGING_CONSTANT
            d.wfq = _fq_struct(mod.weight_fake_quant)
            d.wq, d.wmask, d.sf, d.rstd_run, d.wsum = (ly.wq.data_ptr(), ly.wmask.data_ptr(), ly.sf.data_ptr(),
                                                       ly.rstd.data_ptr(), ly.wsum.data_ptr())
            d.wt_bf16 = ly.wt_bf16.data_ptr() if ly.wt_bf16 is not None else None
            # the stem's GEMM-layout bytes (and its [cout][ldw] gradient) only exist on the fused QuantStub -> im2col path
            use_mma = ly.kind == "pw" or (ly.im2col and self.fused_pw and not self.features)
            d.wq_mma = ly.wq_mma.data_ptr() if (ly.wq_mma is not None and use_mma) else None
            d.ldw = ly.ldw
            d.dwq = ly.dwq.data_ptr()
            d.dweight = g.data_ptr() + 4 * self.param_off[id(mod.weight)]
            if ly.has_bn:
                d.dgamma_bn, d.dsf_bn = ly.dgamma_bn.data_ptr(), ly.dsf_bn.data_ptr()
                d.dgamma = g.data_ptr() + 4 * self.param_off[id(mod.bn.weight)]
            else:
                d.dgamma_bn, d.dsf_bn, d.dgamma = None, None, None
        host = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8)
        return host.to(self.dev)

    def _refresh_flags(self):
        # someone toggled an observer (torch.ao.quantization.disable_observer / load_state_dict): the weight
        # descriptor tables carry the per-layer `observe` flag, re-encode them (rare)
        if self._flag_epoch != Q.FrostFakeQuantize.flag_epoch or self._desc_fused != self.fused_pw:
            self._wdesc_dev = [self._build_wdesc(i) for i in range(2)]
            self._flag_epoch = Q.FrostFakeQuantize.flag_epoch
            self._desc_fused = self.fused_pw

    # ------------------------------------------------------------------ forward pieces
    def _finalize_args(self, ly, xin, M, training, raw):
        mod = ly.mod
        afq, bn = mod.activation_post_process, mod.bn
        a = L.BnFinalizeArgs()
        a.stats, a.C, a.count = ly.stats_ptr, ly.cout, M
        a.stats_format = 1 if raw else 0
        a.x_scale, a.w_scale, a.sf = xin.scale.data_ptr(), mod.weight_fake_quant.scale.data_ptr(), ly.sf.data_ptr()
        a.gamma, a.beta = bn.weight.data_ptr(), bn.bias.data_ptr()
        a.running_mean, a.running_var = bn.running_mean.data_ptr(), bn.running_var.data_ptr()
        a.num_batches_tracked = bn.num_batches_tracked.data_ptr()
        a.momentum = -1.0 if bn.momentum is None else bn.momentum     # None: cumulative average 1/num_batches_tracked
        a.eps = bn.eps
        a.training = 1 if (training and bn.training) else 0
        a.relu = 1 if ly.relu else 0
        a.observe = 1 if afq._observe else 0
        a.averaging_const = Q.AVERAGING_CONSTANT
        a.afq = _fq_struct(afq)
        a.A, a.B, a.mean_I, a.kfac = ly.A.data_ptr(), ly.B.data_ptr(), ly.mean_I.data_ptr(), ly.kfac.data_ptr()
        a.cur_minmax = ly.mm.data_ptr()
        return a

    def _pw_operands(self, ly, xin, M):
        wfq = ly.mod.weight_fake_quant
        op = L.PwOperands()
        op.x, op.M, op.K, op.ldx = xin.q.data_ptr(), M, ly.k_mma, xin.ld
        op.w_mma, op.ldw, op.cout = ly.wq_mma.data_ptr(), ly.ldw, ly.cout
        op.x_zp, op.w_zp, op.wsum = xin.zp.data_ptr(), wfq.zero_point.data_ptr(), ly.wsum.data_ptr()
        return op

    def _conv_bn(self, ly, xin, training, st, saved):
        mod, dev = ly.mod, self.dev
        N = xin.N
        mma = self.fused_pw and (ly.kind == "pw" or (ly.im2col and not isinstance(xin, _RawInput)))
        if ly.kind == "pw" or mma:
            Ho, Wo = xin.H, xin.W              # (the im2col stem's input rows are already per output pixel)
        else:
            Ho = (xin.H + 2 * ly.pad - ly.dil * (ly.kh - 1) - 1) // ly.stride + 1
            Wo = (xin.W + 2 * ly.pad - ly.dil * (ly.kw - 1) - 1) // ly.stride + 1
        M = N * Ho * Wo
        wzp = mod.weight_fake_quant.zero_point
        afq = mod.activation_post_process
        raw = isinstance(xin, _RawInput)          # feature backbone stem: fp32 image, fp32 raw conv output
        a = self._finalize_args(ly, xin, M, training, raw)
        q, ldq = self._alloc_q(M, ly.cout)
        acc = None
        if ly.kind == "pw" and xin.C != ly.cin:
            raise RuntimeError("%s: input has %d channels, expected %d" % (ly.name, xin.C, ly.cin))
        if mma:
            # conv + BN statistics + finalize + observer + quantise: one launch, no int32 accumulator in HBM
            f = L.PwFusedFwdArgs()
            f.op, f.bn = self._pw_operands(ly, xin, M), a
            f.grid_barrier, f.q, f.ldq = ly.barrier_ptr, q.data_ptr(), ldq
            L.call("frost_pw_fused_forward", C.byref(f), st)
        else:
            acc = torch.empty((M, ly.cout), dtype=torch.int32, device=dev)
            if raw:
                L.call("frost_stats_reset_f32", ly.stats_ptr, ly.cout, st)
                L.call("frost_stem_conv_forward_f32", xin.x.data_ptr(), ly.wq.data_ptr(), wzp.data_ptr(), N, xin.H, xin.W, xin.C,
                       ly.cout, ly.kh, ly.stride, ly.pad, acc.data_ptr(), ly.stats_ptr, st)
            elif ly.kind == "pw":
                L.call("frost_pw_conv_forward", xin.q.data_ptr(), xin.zp.data_ptr(), ly.wq.data_ptr(), wzp.data_ptr(),
                       ly.wsum.data_ptr(), M, ly.cin, ly.cout, acc.data_ptr(), ly.stats_ptr, st)
            elif ly.kind == "dw" and ly.gather:
                L.call("frost_dw_conv_forward_dilated", xin.q.data_ptr(), xin.ld, xin.zp.data_ptr(), ly.wq.data_ptr(), wzp.data_ptr(),
                       N, xin.H, xin.W, xin.C, ly.kh, ly.stride, ly.dil, ly.pad, acc.data_ptr(), ly.stats_ptr, st)
            elif ly.kind == "dw":
                L.call("frost_dw_conv_forward", xin.q.data_ptr(), xin.ld, xin.zp.data_ptr(), ly.wq.data_ptr(), wzp.data_ptr(),
                       N, xin.H, xin.W, xin.C, ly.kh, ly.stride, acc.data_ptr(), ly.stats_ptr, st)
            else:
                L.call("frost_stem_conv_forward", xin.q.data_ptr(), xin.zp.data_ptr(), ly.wq.data_ptr(), wzp.data_ptr(),
                       N, xin.H, xin.W, xin.C, ly.cout, ly.kh, ly.stride, ly.pad, acc.data_ptr(), ly.stats_ptr, st)
            L.call("frost_bn_finalize", C.byref(a), st)
            L.call("frost_bnq_apply", acc.data_ptr(), 1 if raw else 0, M, ly.cout, ly.A.data_ptr(), ly.B.data_ptr(), a.relu,
                   afq.scale.data_ptr(), afq.zero_point.data_ptr(), q.data_ptr(), ldq, st)
        self._maybe_force(ly.name, q, ly.cout)
        out = _QT(q, N, Ho, Wo, ly.cout, afq.scale, afq.zero_point, ly.mm, ldq)
        if saved is not None:
            saved[ly.name] = (xin, acc, out, a.training, M)
        if self.record_taps:
            self.last_taps[ly.name + ".out_q"] = q[:, :ly.cout].contiguous()
            if acc is not None:
                self.last_taps[ly.name + ".acc"] = acc
        return out

    def _block_forward(self, b, t, training, st, saved):
        """One bottleneck (frostnet.py:124-145) on the quantised tensor t; returns the block output."""
        dev = self.dev
        blk = b["mod"]
        xin = t
        if saved is not None:
            saved[b["name"] + ".in"] = xin
        if b["conv1"] is not None:
            if b["squeeze"] is not None:
                sq = self._conv_bn(b["squeeze"], xin, training, st, saved)
                cfq = blk.quant_cat.activation_post_process
                qc, ldc = self._alloc_q(xin.M, sq.C + xin.C)
                mmc = torch.empty(2, dtype=torch.float32, device=dev)
                L.call("frost_cat_forward", sq.c(), xin.c(), xin.M, _fq_struct(cfq), 1 if cfq._observe else 0,
                       Q.AVERAGING_CONSTANT, qc.data_ptr(), ldc, mmc.data_ptr(), st)
                self._maybe_force(b["name"] + ".cat", qc, sq.C + xin.C)
                cat = _QT(qc, xin.N, xin.H, xin.W, sq.C + xin.C, cfq.scale, cfq.zero_point, mmc, ldc)
                if saved is not None:
                    saved[b["name"] + ".cat"] = (sq, xin, cat)
                if self.record_taps:
                    self.last_taps[b["name"] + ".cat_q"] = qc[:, :cat.C].contiguous()
                o = cat
            else:
                o = xin
            o = self._conv_bn(b["conv1"], o, training, st, saved)
        else:
            o = xin
        o = self._conv_bn(b["conv2"], o, training, st, saved)
        o = self._conv_bn(b["reduce"], o, training, st, saved)
        if b["skip"]:
            afq = blk.skip_add.activation_post_process
            qa, lda = self._alloc_q(o.M, o.C)
            mma = torch.empty(2, dtype=torch.float32, device=dev)
            L.call("frost_add_forward", xin.c(), o.c(), o.M * o.C, _fq_struct(afq), 1 if afq._observe else 0,
                   Q.AVERAGING_CONSTANT, qa.data_ptr(), lda, mma.data_ptr(), self.scratch.data_ptr(), st)
            self._maybe_force(b["name"] + ".add", qa, o.C)
            s = _QT(qa, o.N, o.H, o.W, o.C, afq.scale, afq.zero_point, mma, lda)
            if saved is not None:
                saved[b["name"] + ".add"] = (xin, o, s)
            if self.record_taps:
                self.last_taps[b["name"] + ".add_q"] = qa[:, :o.C].contiguous()
            o = s
        return o

    def forward(self, x, save):
        self._ensure_built(check=False)          # run() has just verified the pointer tables
        with torch.cuda.device(self.dev):       # kernels launch on the CURRENT device: make it the model's
            return self._forward(x, save)

    def backward(self, saved, dlogits):
        with torch.cuda.device(self.dev):
            return self._backward(saved, dlogits)

    def _forward(self, x, save):
        self._refresh_flags()
        m, dev, st = self.model, self.dev, L.stream(self.dev)
        self.generation += 1
        training = m.training
        if x.device != dev:
            raise RuntimeError("frostnet_b200: input on %s, model on %s" % (x.device, dev))
        x = x.detach()
        if x.dtype != torch.float32 or not x.is_contiguous():
            x = x.float().contiguous()
        N, Cin, H, W = x.shape
        saved = {"gen": self.generation} if save else None
        if self.record_taps:
            self.last_taps = {}
        # all weights: scale_factor, observer, int8 indices - one launch
        L.call("frost_weight_prep_multi", self._wdesc_dev[0].data_ptr(), len(self.layers), self._wchunks.data_ptr(),
               self._n_wchunks, self._wscratch.data_ptr(), st)
        L.call("frost_stats_reset", self.stats.data_ptr(), self.n_stat_chan, st)
        self.barriers.zero_()
        if self.features:
            t = _RawInput(x, self.one)           # frostnet_features.py:342-343: conv1 sees the raw image
        else:
            # QuantStub
            qfq = m.quant.activation_post_process
            stem = self.stem
            use_im2col = self.fused_pw and stem.im2col           # (features: no QuantStub, the stem sees the raw image)
            testing = (self.force is not None and "quant" in self.force) or self.record_taps
            mm_in = torch.empty(2, dtype=torch.float32, device=dev)
            xq = None
            if not use_im2col or testing:
                xq = torch.empty((N, H, W, Cin), dtype=torch.uint8, device=dev)
                L.call("frost_input_quant", x.data_ptr(), N, Cin, H, W, _fq_struct(qfq), 1 if qfq._observe else 0,
                       Q.AVERAGING_CONSTANT, xq.data_ptr(), mm_in.data_ptr(), self.scratch.data_ptr(), st)
                self._maybe_force("quant", xq)
                if self.record_taps:
                    self.last_taps["quant.q"] = xq
            if use_im2col:
                # the stem runs on the fused tensor-core kernels: QuantStub emits its im2col rows directly
                Ho = (H + 2 * stem.pad - stem.kh) // stem.stride + 1
                Wo = (W + 2 * stem.pad - stem.kw) // stem.stride + 1
                if xq is None:
                    cols = torch.empty((N * Ho * Wo, stem.ldw), dtype=torch.uint8, device=dev)
                    L.call("frost_input_quant_im2col", x.data_ptr(), N, Cin, H, W, stem.kh, stem.stride, stem.pad, _fq_struct(qfq),
                           1 if qfq._observe else 0, Q.AVERAGING_CONSTANT, cols.data_ptr(), stem.ldw, mm_in.data_ptr(),
                           self.scratch.data_ptr(), st)
                else:
                    cols = self._im2col_from_indices(xq, stem, int(qfq.zero_point))     # test mode (teacher forcing / taps)
                t = _QT(cols, N, Ho, Wo, stem.k_mma, qfq.scale, qfq.zero_point, mm_in, stem.ldw)
            else:
                t = _QT(xq, N, H, W, Cin, qfq.scale, qfq.zero_point, mm_in)
        feats = []
        t = self._conv_bn(self.stem, t, training, st, saved)
        for b in self.blocks:
            t = self._block_forward(b, t, training, st, saved)
            if self.features and b["stage_end"] and b["stage"] != 3:     # [x1, x2, x3, x5]
                y = torch.empty((t.N, t.C, t.H, t.W), dtype=torch.float32, device=dev)
                L.call("frost_dequant_to_nchw", t.q.data_ptr(), t.ld, t.scale.data_ptr(), t.zp.data_ptr(), t.N, t.H, t.W, t.C,
                       y.data_ptr(), st)
                feats.append(y)
                if saved is not None:
                    saved.setdefault("taps", []).append((b["name"], t))
        if self.features:
            if saved is not None:
                saved["N"] = N
            return feats, saved
        t = self._conv_bn(self.last, t, training, st, saved)
        # head: avg-pool -> dropout -> classifier conv (+bias) -> FQ
        HW = t.H * t.W
        pooled = torch.empty((N, t.C), dtype=torch.float32, device=dev)
        keep, keep_scale = None, 1.0
        p_drop = m.classifier[1].p
        if training and m.classifier[1].training and p_drop > 0:
            keep = self.dropout_mask
            if keep is None:
                keep = torch.empty((N, t.C), dtype=torch.float32, device=dev).bernoulli_(1.0 - p_drop)
            keep = keep.reshape(N, t.C).float().contiguous()
            keep_scale = 1.0 / (1.0 - p_drop)
        L.call("frost_pool_dropout_forward", t.q.data_ptr(), t.scale.data_ptr(), t.zp.data_ptr(), N, HW, t.C,
               L.ptr(keep), keep_scale, pooled.data_ptr(), st)
        cls = self.cls
        cmod = cls.mod
        pre = torch.empty((N, cls.cout), dtype=torch.float32, device=dev)
        L.call("frost_linear_forward", pooled.data_ptr(), cls.wq.data_ptr(), cmod.weight_fake_quant.scale.data_ptr(),
               cmod.weight_fake_quant.zero_point.data_ptr(), L.ptr(cmod.bias), N, cls.cin, cls.cout, pre.data_ptr(), st)
        lfq = cmod.activation_post_process
        logits = torch.empty((N, cls.cout), dtype=torch.float32, device=dev)
        lmask = torch.empty((N, cls.cout), dtype=torch.uint8, device=dev)
        L.call("frost_fq_forward", pre.data_ptr(), pre.numel(), _fq_struct(lfq), Q.ACT_QMIN, Q.ACT_QMAX, 0,
               1 if lfq._observe else 0, Q.AVERAGING_CONSTANT, logits.data_ptr(), lmask.data_ptr(), None,
               self.scratch.data_ptr(), st)
        if self.record_taps:
            self.last_taps["pooled"] = pooled
            self.last_taps["classifier.2.pre"] = pre
        if saved is not None:
            saved["head"] = (t, pooled, keep, keep_scale, lmask, HW)
            saved["N"] = N
        return logits, saved

    # ------------------------------------------------------------------ backward pieces
    def _pick_gflat(self):
        """Use the flat gradient buffer that no live .grad aliases (so autograd can accumulate)."""
        for which in (0, 1):
            g = self.gflat[which]
            lo, hi = g.data_ptr(), g.data_ptr() + 4 * g.numel()
            if not any(p.grad is not None and lo <= p.grad.data_ptr() < hi for p in self.params):
                return which
        raise RuntimeError("frostnet_b200: both flat gradient buffers are aliased by live .grad tensors")

    def _conv_bn_bwd(self, ly, dy, saved, gbase, dx, accumulate, st):
        """dy: fp32 [M, cout] grad wrt the layer's fake-quantised output.  Writes the BN/weight grads and,
        if dx is not None, (accumulates) the grad wrt the layer input into dx [M_in, cin]."""
        xin, acc, out, was_training, M = saved[ly.name]
        mod, dev = ly.mod, self.dev
        # pointwise layers get dz as two bf16 planes (the operand format of the tensor-core dgrad/wgrad);
        # depthwise / stem consumers read fp32.  Expand convs (small K, wide cout) on the fused path never materialise dz:
        # the chained kernel keeps it in shared memory between the BatchNorm-backward epilogue and the dgrad / wgrad MMAs.
        tc_fmt = ly.kind == "pw" or (ly.kind == "stem" and acc is None)      # acc is None <=> the forward took the fused path
        chain = (tc_fmt and acc is None and dx is not None and self.chain_bwd
                 and L.load().frost_pw_chain_supported(ly.cin, ly.cout) != 0)
        if chain:
            dz = dz_lo = None
        elif tc_fmt:
            dz = torch.empty((M, ly.cout), dtype=torch.bfloat16, device=dev)
            dz_lo = torch.empty((M, ly.cout), dtype=torch.bfloat16, device=dev)
        else:
            dz = torch.empty((M, ly.cout), dtype=torch.float32, device=dev)
            dz_lo = None
        a = L.BnBackwardArgs()
        raw = isinstance(xin, _RawInput)
        a.dy, a.acc, a.M, a.C, a.relu = dy.data_ptr(), L.ptr(acc), M, ly.cout, 1 if ly.relu else 0
        a.acc_format = 1 if raw else 0
        a.A, a.B, a.mean_I, a.kfac = ly.A.data_ptr(), ly.B.data_ptr(), ly.mean_I.data_ptr(), ly.kfac.data_ptr()
        a.gamma, a.sf = mod.bn.weight.data_ptr(), ly.sf.data_ptr()
        a.x_scale, a.w_scale = xin.scale.data_ptr(), mod.weight_fake_quant.scale.data_ptr()
        a.out_scale, a.out_zp = out.scale.data_ptr(), out.zp.data_ptr()
        a.eps = mod.bn.eps
        a.sums, a.coef, a.dz = ly.sums.data_ptr(), ly.coef.data_ptr(), L.ptr(dz)
        a.dz_lo, a.dz_format = L.ptr(dz_lo), 1 if tc_fmt else 0
        a.dgamma_bn, a.dsf_bn = ly.dgamma_bn.data_ptr(), ly.dsf_bn.data_ptr()
        a.dbeta = gbase + 4 * self.param_off[id(mod.bn.bias)]
        a.frozen = 0 if was_training else 1        # eval-mode BatchNorm (frozen statistics): still differentiable
        if acc is None:
            # fused 1x1 path: both passes recompute the accumulator tile on the tensor cores from the uint8 rows
            f = L.PwFusedBwdArgs()
            f.op, f.bn = self._pw_operands(ly, xin, M), a
            L.call("frost_pw_fused_bwd_reduce", C.byref(f), st)
            if chain:
                ch = L.PwChainArgs()
                ch.op, ch.bn = f.op, a
                ch.wt_bf16, ch.dx, ch.accumulate, ch.dwq = L.ptr(ly.wt_bf16), L.ptr(dx), 1 if accumulate else 0, ly.dwq.data_ptr()
                L.call("frost_pw_chain_backward", C.byref(ch), st)
                return
            L.call("frost_pw_fused_bwd_apply", C.byref(f), st)
        else:
            L.call("frost_bn_backward_reduce", C.byref(a), st)
            L.call("frost_bn_backward_apply", C.byref(a), st)
        wfq = mod.weight_fake_quant
        # The weight gradient has no consumer until the end of the backward pass: it runs on a side stream next to the
        # data-gradient chain (dgrad -> the next layer's reduce / apply).  Most layers launch fewer CTAs than there are SMs
        # and are latency-bound, so the two streams share the GPU instead of taking turns.
        wst = st
        if self.overlap_wgrad and dx is not None:
            side = self._side_stream()
            side.wait_stream(torch.cuda.current_stream(dev))
            wst = side.cuda_stream
            dz.record_stream(side)                # the caching allocator must not hand dz out again before the side stream is done
            if dz_lo is not None:
                dz_lo.record_stream(side)
        if tc_fmt:
            L.call("frost_pw_wgrad_tc", dz.data_ptr(), dz_lo.data_ptr(), xin.q.data_ptr(), xin.ld, xin.scale.data_ptr(),
                   xin.zp.data_ptr(), M, ly.k_mma, ly.cout, ly.dwq.data_ptr(), wst)
            if dx is not None:
                L.call("frost_pw_dgrad_tc", dz.data_ptr(), dz_lo.data_ptr(), ly.wt_bf16.data_ptr(), wfq.scale.data_ptr(),
                       M, ly.cin, ly.cout, dx.data_ptr(), 1 if accumulate else 0, st)
        elif ly.kind == "dw" and ly.gather:
            L.call("frost_dw_wgrad_dilated", dz.data_ptr(), xin.q.data_ptr(), xin.ld, xin.scale.data_ptr(), xin.zp.data_ptr(), xin.N,
                   xin.H, xin.W, xin.C, ly.kh, ly.stride, ly.dil, ly.pad, ly.dwq.data_ptr(), wst)
            if dx is not None:
                L.call("frost_dw_dgrad_dilated", dz.data_ptr(), ly.wq.data_ptr(), wfq.scale.data_ptr(), wfq.zero_point.data_ptr(),
                       xin.N, xin.H, xin.W, xin.C, ly.kh, ly.stride, ly.dil, ly.pad, dx.data_ptr(), 1 if accumulate else 0, st)
        elif ly.kind == "dw":
            L.call("frost_dw_wgrad", dz.data_ptr(), xin.q.data_ptr(), xin.ld, xin.scale.data_ptr(), xin.zp.data_ptr(), xin.N,
                   xin.H, xin.W, xin.C, ly.kh, ly.stride, ly.dwq.data_ptr(), wst)
            if dx is not None:
                L.call("frost_dw_dgrad", dz.data_ptr(), ly.wq.data_ptr(), wfq.scale.data_ptr(), wfq.zero_point.data_ptr(),
                       xin.N, xin.H, xin.W, xin.C, ly.kh, ly.stride, dx.data_ptr(), 1 if accumulate else 0, st)
        elif raw:
            L.call("frost_stem_wgrad_f32", dz.data_ptr(), xin.x.data_ptr(), xin.N, xin.H, xin.W, xin.C, ly.cout, ly.kh,
                   ly.stride, ly.pad, ly.dwq.data_ptr(), st)
        else:
            L.call("frost_stem_wgrad", dz.data_ptr(), xin.q.data_ptr(), xin.scale.data_ptr(), xin.zp.data_ptr(), xin.N,
                   xin.H, xin.W, xin.C, ly.cout, ly.kh, ly.stride, ly.pad, ly.dwq.data_ptr(), st)

    def _block_backward(self, b, g, saved, gbase, st):
        """Backward of one bottleneck: g = grad wrt the block output [M][C] fp32 (consumed); returns the grad wrt its input."""
        f32 = dict(dtype=torch.float32, device=self.dev)
        dout = g                                   # grad wrt the block output
        conv2, reduce = b["conv2"], b["reduce"]
        xin = saved[b["name"] + ".in"]
        gx = torch.empty((xin.M, xin.C), **f32)    # grad wrt the block input
        gx_written = False
        if b["skip"]:
            xa, ya, sa = saved[b["name"] + ".add"]
            dsum = torch.empty_like(dout)
            L.call("frost_add_backward", dout.data_ptr(), xa.c(), ya.c(), dout.numel(), sa.scale.data_ptr(),
                   sa.zp.data_ptr(), dsum.data_ptr(), gx.data_ptr(), 0, st)
            gx_written = True
            d_reduce = dsum
        else:
            d_reduce = dout
        x_red = saved[reduce.name][0]
        d_dw_out = torch.empty((x_red.M, x_red.C), **f32)
        self._conv_bn_bwd(reduce, d_reduce, saved, gbase, d_dw_out, False, st)
        del d_reduce, dout
        x_dw = saved[conv2.name][0]
        if b["conv1"] is not None:
            d_c1_out = torch.empty((x_dw.M, x_dw.C), **f32)
            self._conv_bn_bwd(conv2, d_dw_out, saved, gbase, d_c1_out, False, st)
            del d_dw_out
            if b["squeeze"] is not None:
                sq, xc, cat = saved[b["name"] + ".cat"]
                d_cat = torch.empty((cat.M, cat.C), **f32)
                self._conv_bn_bwd(b["conv1"], d_c1_out, saved, gbase, d_cat, False, st)
                del d_c1_out
                d_sq = torch.empty((sq.M, sq.C), **f32)
                L.call("frost_cat_backward", d_cat.data_ptr(), sq.c(), xc.c(), cat.M, cat.scale.data_ptr(),
                       cat.zp.data_ptr(), d_sq.data_ptr(), gx.data_ptr(), 1 if gx_written else 0, st)
                gx_written = True
                del d_cat
                self._conv_bn_bwd(b["squeeze"], d_sq, saved, gbase, gx, True, st)
            else:
                self._conv_bn_bwd(b["conv1"], d_c1_out, saved, gbase, gx, gx_written, st)
                gx_written = True
        else:
            self._conv_bn_bwd(conv2, d_dw_out, saved, gbase, gx, gx_written, st)
            gx_written = True
        return gx

    def _backward(self, saved, dlogits):
        """dlogits: grad of the logits (classifier) or the list of 4 NCHW feature-map grads (feature backbone)."""
        if saved.get("gen") != self.generation:
            raise RuntimeError(
                "frostnet_b200: backward through a forward pass that is no longer the latest one - every forward "
                "(also an eval / no_grad one) overwrites the per-layer quantised weights, BN coefficients and "
                "scale/zero_point buffers this graph's backward reads.  Run backward before the next forward.")
        dev, st = self.dev, L.stream(self.dev)
        which = self._pick_gflat()
        gflat = self.gflat[which]
        gbase = gflat.data_ptr()
        f32 = dict(dtype=torch.float32, device=dev)
        N = saved["N"]
        tap_grads = {}
        if self.features:
            for (name, qt), gr in zip(saved["taps"], dlogits):
                if gr is not None:
                    tap_grads[name] = (qt, gr.contiguous().float())
            g = None
        else:
            t_last, pooled, keep, keep_scale, lmask, HW = saved["head"]
            cls, cmod = self.cls, self.cls.mod
            dlogits = dlogits.contiguous().float()
            dpre = torch.empty_like(dlogits)
            L.call("frost_fq_backward", dlogits.data_ptr(), lmask.data_ptr(), dlogits.numel(), dpre.data_ptr(), st)
            dpooled = torch.empty((N, cls.cin), **f32)
            dbias_ptr = gbase + 4 * self.param_off[id(cmod.bias)] if cmod.bias is not None else None
            if cls.cout_p != cls.cout:
                dpre = torch.nn.functional.pad(dpre, (0, cls.cout_p - cls.cout)).contiguous()
                dbias_p = torch.empty(cls.cout_p, **f32)
                dbias_ptr = dbias_p.data_ptr()
            L.call("frost_linear_backward", dpre.data_ptr(), pooled.data_ptr(), cls.wq.data_ptr(),
                   cmod.weight_fake_quant.scale.data_ptr(), cmod.weight_fake_quant.zero_point.data_ptr(), N, cls.cin,
                   cls.cout_p, dpooled.data_ptr(), cls.dwq.data_ptr(), dbias_ptr, st)
            if cls.cout_p != cls.cout and cmod.bias is not None:
                gflat.narrow(0, self.param_off[id(cmod.bias)], cls.cout).copy_(dbias_p[:cls.cout])
            dy = torch.empty((N * HW, t_last.C), **f32)
            L.call("frost_pool_dropout_backward", dpooled.data_ptr(), N, HW, t_last.C, L.ptr(keep), keep_scale,
                   dy.data_ptr(), st)
            xin_last = saved[self.last.name][0]
            g = torch.empty((xin_last.M, xin_last.C), **f32)
            self._conv_bn_bwd(self.last, dy, saved, gbase, g, False, st)
            del dy
        for b in reversed(self.blocks):
            if b["name"] in tap_grads:                 # a returned feature map: its NCHW grad joins the chain
                qt, gr = tap_grads[b["name"]]
                if g is None:
                    g = torch.empty((qt.M, qt.C), **f32)
                    L.call("frost_nchw_to_nhwc", gr.data_ptr(), qt.N, qt.C, qt.H, qt.W, g.data_ptr(), 0, st)
                else:
                    L.call("frost_nchw_to_nhwc", gr.data_ptr(), qt.N, qt.C, qt.H, qt.W, g.data_ptr(), 1, st)
            if g is None:                              # no gradient reaches this block's output
                xo = saved[b["reduce"].name][2] if not b["skip"] else saved[b["name"] + ".add"][2]
                g = torch.zeros((xo.M, xo.C), **f32)
            if self.record_taps:                       # tests: gradient wrt every block output / input (NHWC fp32)
                self.last_grad_taps[b["name"] + ".out"] = g.clone()
            g = self._block_backward(b, g, saved, gbase, st)
            if self.record_taps:
                self.last_grad_taps[b["name"] + ".in"] = g.clone()
        self._conv_bn_bwd(self.stem, g, saved, gbase, None, False, st)
        if self._side is not None:
            torch.cuda.current_stream(dev).wait_stream(self._side)       # every weight gradient is complete
        L.call("frost_weight_backward_multi", self._wdesc_dev[which].data_ptr(), len(self.layers),
               self._wbchunks.data_ptr(), self._n_wbchunks, st)
        if self.grad_sync is not None:
            self.grad_sync(gflat)
        grads = []
        for p in self.params:
            o = self.param_off[id(p)]
            grads.append(gflat.narrow(0, o, p.numel()).view(p.shape))
        return grads

    # ------------------------------------------------------------------ entry point
    def run(self, x):
        self._ensure_built()
        if not x.is_cuda:
            raise RuntimeError("frostnet_b200: the QAT path runs on a CUDA device (B200) only; got a CPU tensor")
        if x.requires_grad and torch.is_grad_enabled():
            raise RuntimeError("frostnet_b200: the input requires grad, but the QAT engine does not propagate a gradient to "
                               "the image (the stem has no dgrad); detach the input")
        need_grad = torch.is_grad_enabled() and any(p.requires_grad for p in self.params)
        if need_grad:
            out = _QATFunction.apply(self, x, *self.params)
            return list(out) if self.features else out
        out, _ = self.forward(x, save=False)
        return out


class _QATFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, engine, x, *params):
        out, saved = engine.forward(x, save=True)
        ctx.engine, ctx.saved = engine, saved
        return tuple(out) if engine.features else out

    @staticmethod
    def backward(ctx, *douts):
        grads = ctx.engine.backward(ctx.saved, list(douts) if ctx.engine.features else douts[0])
        ctx.saved = None
        return (None, None) + tuple(grads)
