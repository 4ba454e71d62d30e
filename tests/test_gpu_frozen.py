"""the frozen-BatchNorm / observer-off mode - late-QAT fine-tuning (Classification/train.py:27-33
defines disable_observer; frostnet_features.py:354-359 _freeze_stages puts every BatchNorm in eval mode while the
backbone keeps training).  BatchNorm then is a fixed per-channel affine, the observers hold their scales, and the fused
1x1 kernels run ONE pass per layer (no statistics phase, no grid barrier).  Forward AND backward against the oracle."""
import pytest
import torch
import torch.nn.functional as Fn

from util import rel_l2
from test_gpu_net import GRAD_REL_L2, LAYER_MISMATCH_RATE, LOGIT_REL_L2, _force_dict, _grad_rel_l2

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _pair(nclass=24, seed=5):
    import frostnet_b200 as F
    from oracle import frost_oracle as O
    spec = O.net_spec("small", 0.35, nclass)
    sd = O.fresh_state_dict(spec, seed=seed)
    g = torch.Generator().manual_seed(3)
    for k in sd:
        if k.endswith("bn.weight"):
            sd[k] = 0.5 + torch.rand(sd[k].shape, generator=g)
        if k.endswith("bn.bias"):
            sd[k] = 0.2 * torch.randn(sd[k].shape, generator=g)
    onet = O.OracleNet(spec, sd)
    model = F.FrostNet(nclass=nclass, mode="small", width_mult=0.35, quantized=True, drop_rate=0.0)
    model.train()
    model.fuse_model()
    F.prepare_qat(model)
    model.to(DEV)
    return model, onet


def test_frozen_bn_and_observers_off_forward_backward_vs_oracle():
    model, onet = _pair()
    g = torch.Generator().manual_seed(11)
    x0 = torch.randn(8, 3, 96, 96, generator=g)
    # one ordinary QAT step initialises running statistics and observers in the ORACLE; the device model takes that state
    onet.forward(x0, training=True, drop_rate=0.0)
    model.load_state_dict(onet.state_dict(), strict=True)
    # freeze: BatchNorm in eval mode, observers off (fake-quant stays on)
    for fq in onet.fq.values():
        fq.observer_enabled = 0
    for m in model.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            m.eval()
    model.apply(torch.ao.quantization.disable_observer)
    sd_frozen = {k: v.clone() for k, v in model.state_dict().items()}
    x = torch.randn(8, 3, 96, 96, generator=g)
    y = torch.randint(0, 24, (8,), generator=g)
    onet.record = True
    for p in onet.parameters():
        p.grad = None
    ologits = onet.forward(x, training=False)
    Fn.cross_entropy(ologits, y).backward()
    eng = model._frost_engine
    eng.force, eng.force_report = _force_dict(onet, torch.device(DEV)), {}
    logits = model(x.to(DEV))
    Fn.cross_entropy(logits, y.to(DEV)).backward()
    rep = eng.force_report
    worst = max(rep.items(), key=lambda kv: kv[1][0])
    assert all(mx <= 1 for _, mx in rep.values()) and worst[1][0] <= LAYER_MISMATCH_RATE, worst
    e_log = rel_l2(logits.detach().cpu(), ologits.detach())
    e_grad, wg = _grad_rel_l2(model, onet)
    print("frozen BN + observers off: worst layer %s %.3g; logits rel-L2 %.3g; grad rel-L2 %.3g (worst %s)" % (
        worst[0], worst[1][0], e_log, e_grad, wg))
    assert e_log < LOGIT_REL_L2 and e_grad < GRAD_REL_L2
    # nothing that is frozen moved
    for k, v in model.state_dict().items():
        if "running_" in k or "num_batches" in k or k.endswith("scale") or k.endswith("zero_point") or k.endswith("min_val") \
                or k.endswith("max_val"):
            if "weight_fake_quant" in k:
                continue                       # weight observers were disabled too; their state must hold as well
            assert torch.equal(v, sd_frozen[k]), k


def test_feature_backbone_trains_after_freeze_stages():
    """frostnet_features.FrostNet._freeze_stages() (mmdet norm_eval): backward through eval-mode BatchNorm."""
    import frostnet_b200 as F
    from frostnet_b200 import frostnet_features as FF
    torch.manual_seed(0)
    model = FF.FrostNet(mode="small", width_mult=0.35, quantized=True)
    model.train()
    model.fuse_model()
    F.prepare_qat(model)
    model.to(DEV)
    x = torch.rand(2, 3, 64, 64, device=DEV)
    for f in model(x):                          # an ordinary step first (running statistics, observers)
        pass
    model._freeze_stages()
    rm = {k: v.clone() for k, v in model.state_dict().items() if "running_" in k}
    feats = model(x)
    sum(f.square().mean() for f in feats).backward()
    grads = [p.grad for p in model.parameters()]
    assert all(g is not None and torch.isfinite(g).all() for g in grads)
    assert sum(float(g.abs().sum()) for g in grads) > 0
    for k, v in model.state_dict().items():
        if k in rm:
            assert torch.equal(v, rm[k]), k     # frozen statistics did not move
