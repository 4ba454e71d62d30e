"""Golden vectors for the int8 inference export: the REAL reference model, three QAT steps, then the
reference's own conversion (Classification/evaluate.py:131, torch.quantization.convert(model.eval())) and its int8 logits.
Runs only in the build container; the output tests/golden/int8_small035.pt is committed.

    python tests/golden/make_golden_int8.py
"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as G  # noqa: E402


def qdigest(sd):
    """state_dict of a converted model -> plain tensors (quantized weights as int_repr + qparams)"""
    out = {}
    for k, v in sd.items():
        if isinstance(v, torch.Tensor) and v.is_quantized:
            out[k] = dict(int_repr=v.int_repr().clone(), scale=float(v.q_scale()), zero_point=int(v.q_zero_point()))
        elif isinstance(v, torch.Tensor):
            out[k] = v.clone()
        elif isinstance(v, (int, float, type(None), torch.dtype)):
            out[k] = v
        else:
            out[k] = repr(v)
    return out


def main():
    ref_frostnet, _, _ = G.load_reference()
    torch.backends.quantized.engine = "qnnpack"          # the engine the qnnpack qconfig is meant for
    torch.manual_seed(1882)
    mode, wm, ncls, N, R = "small", 0.35, 16, 4, 64
    model = ref_frostnet.FrostNet(nclass=ncls, mode=mode, width_mult=wm, quantized=True, drop_rate=0.0)
    g = torch.Generator().manual_seed(11)
    for m in model.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            m.weight.data = 0.5 + torch.rand(m.weight.shape, generator=g)
            m.bias.data = 0.2 * torch.randn(m.bias.shape, generator=g)
            m.running_var.data = 0.5 + torch.rand(m.running_var.shape, generator=g)
            m.running_mean.data = 0.1 * torch.randn(m.running_mean.shape, generator=g)
    G.prepare(model)
    crit = torch.nn.CrossEntropyLoss()
    for i in range(3):                                    # observers, running statistics and weights all move
        x = torch.randn(N, 3, R, R, generator=g)
        y = torch.randint(0, ncls, (N,), generator=g)
        model.zero_grad()
        crit(model(x), y).backward()
        with torch.no_grad():
            for p in model.parameters():
                p.add_(p.grad, alpha=-0.05)
    x_test = torch.randn(N, 3, R, R, generator=g)
    model.eval()
    with torch.no_grad():
        qat_logits = model(x_test).clone()            # NB: observers stay on in eval mode - this forward moves their state
        sd = {k: v.clone() for k, v in model.state_dict().items()}     # the state the conversion starts from
        qmodel = torch.ao.quantization.convert(model, inplace=False)
        int8_logits = qmodel(x_test).clone()
    out = dict(mode=mode, width_mult=wm, nclass=ncls, sd=sd, x=x_test, qat_logits=qat_logits, int8_logits=int8_logits,
               engine="qnnpack", converted=qdigest(qmodel.state_dict()), torch=torch.__version__)
    torch.save(out, os.path.join(HERE, "int8_small035.pt"))
    agree = float((qat_logits - int8_logits).abs().max()), float(qat_logits.abs().max())
    print("int8 golden ok: %d converted entries, |qat - int8|max = %.4g of %.4g" % (len(out["converted"]), *agree))


if __name__ == "__main__":
    main()
