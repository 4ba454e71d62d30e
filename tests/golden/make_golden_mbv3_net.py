"""Golden vectors for the whole MobileNetV3: the REAL reference network
(Classification/models/imagenet/mobilenetv3.py:162-383, mobilenet_v3_small, 10 classes) fused by its own
fuse_model(), prepared with the qnnpack QAT qconfig, two training-mode forward/backward passes at 64x64.  The reference's
forward hard-wires F.dropout(p=0.8) on the last feature map; its random mask cannot be shared with another device's generator,
so the golden run replaces it by the identity (our test sets drop_rate = 0).  Weights are not stored: both sides fill them
from generators seeded by the parameter names (tests/util.py fill_params_by_name, applied AFTER fuse + prepare, where both
trees name their parameters alike); per step the file holds logits, loss, every parameter gradient's norm plus a few whole gradients,
and the non-weight state (observers, BatchNorm statistics).
Runs only in the build container; tests/golden/mbv3_net.pt is committed.

    python tests/golden/make_golden_mbv3_net.py
"""
import importlib.util
import os
import warnings

import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from util import fill_params_by_name  # noqa: E402

warnings.filterwarnings("ignore")
HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/Classification/models/imagenet/mobilenetv3.py"
FULL_GRADS = ("conv1.cb.cb.0.weight", "layer1.0.conv.2.fc.0.weight", "layer3.2.conv.1.cb.0.weight", "layer5.cb.cb.0.bn.weight",
              "classifier.0.fc.2.weight", "classifier.4.bias")


def main():
    torch.quantization.fuse_modules = torch.ao.quantization.fuse_modules_qat
    spec = importlib.util.spec_from_file_location("ref_mbv3", REF)
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    ref.F.dropout = lambda x, p=0.5, training=True, inplace=False: x
    torch.manual_seed(1882)
    net = ref.get_mobilenet_v3("small", 1.0, nclass=10)
    float_keys = list(net.state_dict().keys())
    net.train()
    net.fuse_model()
    net.qconfig = torch.ao.quantization.get_default_qat_qconfig("qnnpack")
    torch.ao.quantization.prepare_qat(net, inplace=True)
    fill_params_by_name(net)                    # after prepare: the parameter names are the final ones on both sides
    sd0_keys = {k: (tuple(v.shape), str(v.dtype)) for k, v in net.state_dict().items()}
    g = torch.Generator().manual_seed(3)
    # step 0 also records the output of every top-level member (teacher forcing: the test feeds each of OUR blocks the
    # reference's input, so that one flipped index does not travel through eleven BatchNorm'd blocks)
    tap_names = (["quant", "conv1"] + ["layer%d.%d" % (li, bi) for li in (1, 2, 3, 4) for bi in range(len(getattr(net, "layer%d" % li)))]
                 + ["layer5"] + ["classifier.%d" % i for i in range(len(net.classifier))])
    mods = dict(net.named_modules())
    taps = {}

    def tap(name):
        def hook(mod, inp, out):
            taps[name] = out.detach().clone()          # (returns None: the output is not replaced)
        return hook
    hooks = [mods[n].register_forward_hook(tap(n)) for n in tap_names]
    steps = []
    for i in range(2):
        net.zero_grad()
        x = torch.randn(4, 3, 64, 64, generator=g)
        t = torch.randint(0, 10, (4,), generator=g)
        logits = net(x)
        for h in hooks:
            h.remove()
        hooks = []
        loss = torch.nn.functional.cross_entropy(logits, t)
        loss.backward()
        grads = {n: p.grad for n, p in net.named_parameters()}
        steps.append(dict(x=x, t=t, logits=logits.detach().clone(), loss=float(loss),
                          grad_norms={n: float(v.norm()) for n, v in grads.items()},
                          grads={n: grads[n].clone() for n in FULL_GRADS},
                          state={k: v.clone() for k, v in net.state_dict().items()
                                 if not (k.endswith(".weight") and v.dim() > 1) and not k.endswith(".bias")}))
        print("step", i, "loss", float(loss), "logits", logits[0, :4].tolist())
    torch.save(dict(float_keys=float_keys, sd0_keys=sd0_keys, steps=steps, tap_names=tap_names, taps=taps, param_names=[n for n, _ in net.named_parameters()],
                    torch=torch.__version__), os.path.join(HERE, "mbv3_net.pt"))
    print("params", sum(p.numel() for p in net.parameters()), "keys", len(sd0_keys))


if __name__ == "__main__":
    main()
