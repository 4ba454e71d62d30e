"""Generate golden vectors by running the REAL reference (/root/reference) under torch on CPU.

Runs only in the build container (the GPU box has no /root/reference); its outputs
(``tests/golden/*.pt``) are committed.  It also asserts that ``oracle/frost_oracle.py`` is
bit-identical to the reference on the same inputs, which is what pins the oracle.

    python tests/golden/make_golden.py

Shims: stub ``timm`` (constants + identity register_model),
``torch.quantization.fuse_modules = fuse_modules_qat`` (the reference fuses in train mode,
Classification/train.py:171, which torch>=1.11 only allows through the _qat entry point),
``Tensor.cuda = identity`` (hard-coded .cuda() at optimizer.py:180), and a fake parent package
for frostnet_features.py's ``from ..builder import BACKBONES``.
"""
import importlib.util
import os
import sys
import types
import warnings

import numpy as np
import torch

warnings.filterwarnings("ignore")
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)


def install_shims():
    timm = types.ModuleType("timm")
    data = types.ModuleType("timm.data")
    models = types.ModuleType("timm.models")
    reg = types.ModuleType("timm.models.registry")
    data.IMAGENET_DEFAULT_MEAN = (0.485, 0.456, 0.406)
    data.IMAGENET_DEFAULT_STD = (0.229, 0.224, 0.225)
    data.IMAGENET_INCEPTION_MEAN = (0.5, 0.5, 0.5)
    data.IMAGENET_INCEPTION_STD = (0.5, 0.5, 0.5)
    reg.register_model = lambda f: f
    sys.modules.update({"timm": timm, "timm.data": data, "timm.models": models,
                        "timm.models.registry": reg})
    torch.quantization.fuse_modules = torch.ao.quantization.fuse_modules_qat
    torch.Tensor.cuda = lambda self, *a, **k: self


def load_reference():
    install_shims()
    sys.path.insert(0, REF)
    import frostnet as ref_frostnet
    import optimizer as ref_optimizer
    # frostnet_features needs a parent package with builder.BACKBONES
    pkg = types.ModuleType("refpkg")
    pkg.__path__ = []
    sub = types.ModuleType("refpkg.backbones")
    sub.__path__ = []
    builder = types.ModuleType("refpkg.builder")

    class _Reg:
        def register_module(self):
            return lambda c: c
    builder.BACKBONES = _Reg()
    sys.modules.update({"refpkg": pkg, "refpkg.backbones": sub, "refpkg.builder": builder})
    spec = importlib.util.spec_from_file_location("refpkg.backbones.frostnet_features",
                                                  os.path.join(REF, "frostnet_features.py"))
    feat = importlib.util.module_from_spec(spec)
    sys.modules[spec.name] = feat
    spec.loader.exec_module(feat)
    return ref_frostnet, ref_optimizer, feat


def prepare(model):
    """Classification/train.py:166-173."""
    model.train()
    model.fuse_model()
    model.qconfig = torch.ao.quantization.get_default_qat_qconfig("qnnpack")
    torch.ao.quantization.prepare_qat(model, inplace=True)
    return model


def param_groups(model, weight_decay):
    """Classification/train.py:121-137: one group per tensor, wd by shape."""
    groups = []
    for name, p in model.named_parameters():
        if p.dim() == 4:
            wd = 0.0 if p.shape[1] == 1 else weight_decay
        else:
            wd = weight_decay * 0.01
        groups.append({"params": [p], "weight_decay": wd})
    return groups


# activation-index taps kept in the fixture (weights' indices are kept for every layer)
TAP_KEEP = ("quant_idx", "conv1.", "layer3.1.", "layer5.0.", "last_layer.", "classifier.")


class Args:
    learning_rate = 5e-3
    weight_decay = 1e-5
    nesterov = True
    clip_by = 1e-3
    toss_coin = True
    noise_decay = 1e-2
    amsgrad = False


def main():
    from oracle import frost_oracle as O
    ref_frostnet, ref_optimizer, ref_feat = load_reference()
    out = {}

    # ------------------------------------------------------------------ G1: tiny classifier net
    torch.manual_seed(1882)
    mode, wm, ncls, N, R = "small", 0.35, 16, 4, 64
    model = ref_frostnet.FrostNet(nclass=ncls, mode=mode, width_mult=wm, quantized=True, drop_rate=0.0)
    # de-trivialise BN so that gamma/beta/running stats matter
    g = torch.Generator().manual_seed(7)
    for m in model.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            m.weight.data = 0.5 + torch.rand(m.weight.shape, generator=g)
            m.bias.data = 0.2 * torch.randn(m.bias.shape, generator=g)
            m.running_var.data = 0.5 + torch.rand(m.running_var.shape, generator=g)
            m.running_mean.data = 0.1 * torch.randn(m.running_mean.shape, generator=g)
    prepare(model)
    sd0 = {k: v.clone() for k, v in model.state_dict().items()}
    spec = O.net_spec(mode, wm, ncls)
    onet = O.OracleNet(spec, sd0)
    assert list(dict(model.named_parameters()).keys()) == list(onet.P.keys()), "param order/keys differ"
    xs = [torch.randn(N, 3, R, R, generator=g) * (1.0 + 0.3 * i) for i in range(3)]
    ys = [torch.randint(0, ncls, (N,), generator=g) for _ in range(3)]
    steps = []
    crit = torch.nn.CrossEntropyLoss()
    for i in range(3):
        model.zero_grad()
        logits = model(xs[i])
        loss = crit(logits, ys[i])
        loss.backward()
        onet.record = True
        for p in onet.parameters():
            p.grad = None
        ologits = onet.forward(xs[i], training=True, drop_rate=0.0)
        oloss = crit(ologits, ys[i])
        oloss.backward()
        assert torch.equal(logits, ologits), f"oracle logits differ from reference at step {i}"
        ref_grads = {k: p.grad.clone() for k, p in model.named_parameters()}
        for k, p in onet.named_parameters():
            assert torch.equal(ref_grads[k], p.grad), f"oracle grad {k} differs at step {i}"
        sd_ref = model.state_dict()
        sd_or = onet.state_dict()
        for k in sd_ref:
            assert torch.equal(sd_ref[k], sd_or[k]), f"oracle state {k} differs at step {i}"
        steps.append(dict(logits=logits.detach().clone(), loss=loss.detach().clone(),
                          grads=ref_grads if i == 2 else None,
                          taps={k: (v.clamp(-128, 127).to(torch.int8) if k.endswith("w_idx") else v.clamp(-2, 300).to(torch.int16))
                                for k, v in onet.taps.items()
                                if k.endswith("w_idx") or (k.endswith("idx") and k.startswith(TAP_KEEP))}
                          if i == 2 else None,
                          state={k: v.clone() for k, v in sd_ref.items()
                                 if not (k.endswith("weight") or k.endswith("bias")) or ".bn." in k}))
        # plain SGD between steps so that weights move and EMA paths are exercised
        with torch.no_grad():
            for (k, p), op in zip(model.named_parameters(), onet.parameters()):
                p.add_(p.grad, alpha=-0.05)
                op.add_(op.grad, alpha=-0.05)
    out["net"] = dict(mode=mode, width_mult=wm, nclass=ncls, sd0=sd0, xs=xs, ys=ys, steps=steps)
    print("G1 ok: oracle == reference bit-exact over 3 QAT steps (logits, grads, buffers)")

    # ------------------------------------------------------------------ G2: features backbone
    torch.manual_seed(5)
    fmodel = ref_feat.FrostNet(mode="small", width_mult=0.35, quantized=True)
    fmodel.init_weights("")
    fmodel.train()
    for m in fmodel.modules():
        if type(m).__name__ in ("ConvBNReLU", "ConvBN"):
            m.fuse_model()
    fmodel.qconfig = torch.ao.quantization.get_default_qat_qconfig("qnnpack")
    torch.ao.quantization.prepare_qat(fmodel, inplace=True)
    fsd0 = {k: v.clone() for k, v in fmodel.state_dict().items()}
    fx = torch.randn(2, 3, 64, 64, generator=g).abs()   # no QuantStub: feed a non-negative image
    feats = fmodel(fx)
    fnet = O.OracleNet(O.net_spec("small", 0.35), fsd0, features=True)
    ofeats = fnet.forward(fx, training=True)
    for a, b in zip(feats, ofeats):
        assert torch.equal(a, b)
    out["features"] = dict(mode="small", width_mult=0.35, sd0=fsd0, x=fx,
                           feats=[f.detach().clone() for f in feats])
    print("G2 ok: features oracle == reference", [tuple(f.shape) for f in feats])

    # ------------------------------------------------------------------ G3: GradBoost optimizers
    gb = {}
    for kind, cls, kw in [
        ("QSGD", ref_optimizer.QSGD, dict(lr=5e-3, momentum=0.9, weight_decay=1e-5, nesterov=True)),
        ("QRMS", ref_optimizer.QRMSprop, dict(lr=1e-3, alpha=0.9, momentum=0.9, eps=1e-8, weight_decay=1e-5)),
        ("QRMSc", ref_optimizer.QRMSprop, dict(lr=1e-3, alpha=0.9, momentum=0.0, eps=1e-8, weight_decay=1e-5, centered=True)),
        ("QAdam", ref_optimizer.QAdam, dict(lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-5, amsgrad=True)),
        ("QAdamW", ref_optimizer.QAdamW, dict(lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2)),
    ]:
        gg = torch.Generator().manual_seed(11)
        shapes = [(8, 4, 3, 3), (16,), (5, 7)]
        p0 = [torch.randn(s, generator=gg) for s in shapes]
        params = [torch.nn.Parameter(p.clone()) for p in p0]
        opt = cls(params, clip_by=1e-3, toss_coin=True, noise_decay=1e-2, **kw)
        grads_seq, noises, coins, snaps = [], [], [], []
        # injectable randomness: capture what the reference draws
        nsteps, warm = 6, 3
        for t in range(nsteps):
            if t == warm:
                opt.is_warmup = False
            gs = [torch.randn(s, generator=gg) * (0.1 if t % 2 else 1e-4) for s in shapes]
            grads_seq.append([x.clone() for x in gs])
            for p, x in zip(params, gs):
                p.grad = x.clone()
            np.random.seed(100 + t)
            torch.manual_seed(200 + t)
            opt.step()
            # replay the RNG streams to record the draws in order
            np.random.seed(100 + t)
            torch.manual_seed(200 + t)
            if t >= warm:
                ns, cs = [], []
                for s in shapes:
                    ns.append(torch.from_numpy(np.abs(np.random.laplace(0.0, 1.0, s))).float())
                    cs.append(torch.zeros(s).random_(2))
                noises.append(ns)
                coins.append(cs)
            snaps.append(dict(params=[p.detach().clone() for p in params],
                              grads_after=[p.grad.clone() for p in params],
                              exp_max=[opt.state[p]["exp_max"].clone() for p in params],
                              exp_min=[opt.state[p]["exp_min"].clone() for p in params]))
        final_state = [{k: (v.clone() if torch.is_tensor(v) else v) for k, v in opt.state[p].items()} for p in params]
        # oracle restatement with injected draws must agree bit-exactly
        it = {"n": None, "c": None}
        okind = "QRMS" if kind.startswith("QRMS") else kind
        okw = dict(kw)
        olr = okw.pop("lr")
        oparams = [p.clone() for p in p0]
        oopt = O.GradBoost(okind, oparams, olr, clip_by=1e-3, toss_coin=True, noise_decay=1e-2,
                           noise_fn=lambda shape: next(it["n"]).clone(), coin_fn=lambda shape: next(it["c"]).clone(), **okw)
        for t in range(nsteps):
            if t == warm:
                oopt.is_warmup = False
            if t >= warm:
                it["n"] = iter(noises[t - warm])
                it["c"] = iter(coins[t - warm])
            ogr = [x.clone() for x in grads_seq[t]]
            oopt.step(ogr)
            for a, b in zip(oparams, snaps[t]["params"]):
                assert torch.equal(a, b), f"{kind} oracle params differ at step {t}"
            for a, b in zip(ogr, snaps[t]["grads_after"]):
                assert torch.equal(a, b), f"{kind} oracle grads differ at step {t}"
        gb[kind] = dict(kw=kw, shapes=shapes, p0=p0, grads=grads_seq, noises=noises, coins=coins,
                        warm=warm, snaps=snaps, final_state=final_state)
        print(f"G3 ok: {kind} oracle == reference bit-exact over {nsteps} steps")
    out["gradboost"] = gb

    # ------------------------------------------------------------------ G4: the headline topology (Large 1.0)
    # FrostNet-Large-1.0 is 5.8 M parameters: the fixture keeps the SEED of the state (oracle.fresh_state_dict is a
    # deterministic function of it) instead of 23 MB of weights, the input, and digests of what the reference produced.
    lspec = O.net_spec("large", 1.0, 1000)
    lseed = 1882
    lmodel = ref_frostnet.FrostNet(nclass=1000, mode="large", width_mult=1.0, quantized=True, drop_rate=0.0)
    prepare(lmodel)
    missing, unexpected = lmodel.load_state_dict(O.fresh_state_dict(lspec, seed=lseed), strict=False)
    assert not unexpected and all("fake_quant" in k or "activation_post_process" in k for k in missing), missing
    lsd0 = {k: v.clone() for k, v in lmodel.state_dict().items()}
    lnet = O.OracleNet(lspec, lsd0)
    lnet.record = True
    assert list(dict(lmodel.named_parameters()).keys()) == list(lnet.P.keys())
    assert len(lnet.P) == 209 and sum(p.numel() for p in lnet.parameters()) == 5807056
    lx = torch.randn(2, 3, 64, 64, generator=g)
    ly = torch.randint(0, 1000, (2,), generator=g)
    llogits = lmodel(lx)
    lloss = crit(llogits, ly)
    lloss.backward()
    ologits = lnet.forward(lx, training=True, drop_rate=0.0)
    oloss = crit(ologits, ly)
    oloss.backward()
    assert torch.equal(llogits, ologits), "Large: oracle logits differ from the reference"
    lgrads = {k: p.grad for k, p in lmodel.named_parameters()}
    for k, p in lnet.named_parameters():
        assert torch.equal(lgrads[k], p.grad), "Large: oracle grad %s differs" % k
    lsd_ref, lsd_or = lmodel.state_dict(), lnet.state_dict()
    for k in lsd_ref:
        assert torch.equal(lsd_ref[k], lsd_or[k]), "Large: oracle state %s differs" % k
    out["large"] = dict(
        mode="large", width_mult=1.0, nclass=1000, seed=lseed, x=lx, y=ly, logits=llogits.detach().clone(),
        loss=lloss.detach().clone(),
        grad_norm={k: float(v.double().norm()) for k, v in lgrads.items()},
        grad_head={k: v.flatten()[:8].clone() for k, v in lgrads.items()},
        state={k: v.clone() for k, v in lsd_ref.items()
               if v.numel() == 1 and ("activation_post_process" in k or "fake_quant" in k)},
        bn_digest={k: float(v.double().sum()) for k, v in lsd_ref.items() if "running_" in k},
        w_idx_digest={k: int(v.clamp(-128, 127).long().abs().sum()) for k, v in lnet.taps.items() if k.endswith("w_idx")},
        stem_out_idx=lnet.taps["conv1.conv.0.out_idx"].clamp(0, 255).to(torch.uint8))
    print("G4 ok: Large-1.0 oracle == reference bit-exact (logits, 209 grads, all buffers), N=2 64x64")

    torch.save(out["large"], os.path.join(HERE, "net_large10_digest.pt"))
    torch.save(out["net"], os.path.join(HERE, "net_small035.pt"))
    torch.save(out["features"], os.path.join(HERE, "features_small035.pt"))
    torch.save(out["gradboost"], os.path.join(HERE, "gradboost.pt"))
    for f in ("net_small035.pt", "features_small035.pt", "gradboost.pt", "net_large10_digest.pt"):
        print(f, os.path.getsize(os.path.join(HERE, f)) // 1024, "KiB")


if __name__ == "__main__":
    main()
