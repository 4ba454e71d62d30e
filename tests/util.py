"""Shared helpers for the parity tests (the oracle is imported here and ONLY in tests/bench/smoke)."""
import os

import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    return torch.load(os.path.join(GOLDEN, name), weights_only=False)


def build_model_from_golden(g, device):
    """frostnet_b200 model in the reference's post-prepare state of the golden fixture."""
    import frostnet_b200 as F
    model = F.FrostNet(nclass=g["nclass"], mode=g["mode"], width_mult=g["width_mult"], quantized=True, drop_rate=0.0)
    model.train()
    model.fuse_model()
    F.prepare_qat(model)
    missing, unexpected = model.load_state_dict(g["sd0"], strict=True)
    assert not missing and not unexpected
    return model.to(device)


def nchw_idx_to_nhwc_u8(idx):
    """oracle tap (unclamped float index, NCHW) -> clamped uint8 NHWC"""
    return idx.clamp(0, 255).to(torch.uint8).permute(0, 2, 3, 1).contiguous()


def rel_l2(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def fill_params_by_name(model):
    """Deterministic parameters that depend only on each parameter's NAME and shape (a generator seeded with crc32(name)):
    the reference network in the build container and our network on the GPU box get identical weights without storing them
    (tests/golden/make_golden_mbv3_net.py).  Conv / Linear weights ~ N(0, 2 / fan_out) (x4 for the tiny SE layers so that the
    gates move), BatchNorm weight 1 + 0.1 N, biases 0.1 N."""
    import zlib
    with torch.no_grad():
        for name, p in model.named_parameters():
            g = torch.Generator().manual_seed(zlib.crc32(name.encode()))
            r = torch.randn(p.shape, generator=g)
            if p.dim() >= 2:
                fan_out = p.shape[0] * (p[0, 0].numel() if p.dim() > 2 else 1)
                std = (2.0 / fan_out) ** 0.5 * (4.0 if ".fc." in name else 1.0)
                p.copy_(r * std)
            elif name.endswith("bn.weight"):
                p.copy_(1.0 + 0.1 * r)
            else:
                p.copy_(0.1 * r)
    return model


def qdigest_compact(sd):
    """state_dict of a converted (int8) model -> comparable plain values: quantized tensors as (shape, qparams, SHA-1 of the int8
    bytes), other tensors whole when small and hashed when large, packed-parameter tuples element by element."""
    import hashlib

    def one(v):
        if isinstance(v, torch.Tensor) and v.is_quantized:
            if v.qscheme() in (torch.per_channel_affine, torch.per_channel_symmetric):
                qp = (v.q_per_channel_scales().tolist(), v.q_per_channel_zero_points().tolist())
            else:
                qp = (float(v.q_scale()), int(v.q_zero_point()))
            return dict(shape=tuple(v.shape), qparams=qp, sha1=hashlib.sha1(v.int_repr().contiguous().numpy().tobytes()).hexdigest())
        if isinstance(v, torch.Tensor):
            if v.numel() <= 4096:
                return v.clone()
            return dict(shape=tuple(v.shape), sha1=hashlib.sha1(v.contiguous().numpy().tobytes()).hexdigest())
        if isinstance(v, (tuple, list)):
            return [one(e) for e in v]
        if isinstance(v, (int, float, type(None), torch.dtype, str)):
            return v
        return repr(v)
    return {k: one(v) for k, v in sd.items()}
