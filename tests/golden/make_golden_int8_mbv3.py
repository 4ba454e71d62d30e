"""Golden vectors for the int8 inference export of MobileNetV3 (SURVEY.md 8f, f1 + f4): the REAL reference network
(Classification/models/imagenet/mobilenetv3.py, mobilenet_v3_small, 10 classes), fused + prepared, name-seeded weights
(tests/util.py fill_params_by_name), two training-mode forward passes so that observers and BatchNorm statistics move, then the
reference's own conversion (Classification/evaluate.py:131, torch.quantization.convert(model.eval())) and its int8 logits.
Stored: the non-weight state the conversion starts from, the input, the int8 logits, and a digest of the converted state_dict
(qparams + SHA-1 of every quantized weight's int8 bytes, small tensors whole).  Runs only in the build container;
tests/golden/int8_mbv3.pt is committed.

    python tests/golden/make_golden_int8_mbv3.py
"""
import hashlib
import importlib.util
import os
import sys
import warnings

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from util import fill_params_by_name, qdigest_compact  # noqa: E402

warnings.filterwarnings("ignore")
HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/Classification/models/imagenet/mobilenetv3.py"


def main():
    torch.quantization.fuse_modules = torch.ao.quantization.fuse_modules_qat
    torch.backends.quantized.engine = "qnnpack"
    spec = importlib.util.spec_from_file_location("ref_mbv3", REF)
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    ref.F.dropout = lambda x, p=0.5, training=True, inplace=False: x
    net = ref.get_mobilenet_v3("small", 1.0, nclass=10)
    net.train()
    net.fuse_model()
    net.qconfig = torch.ao.quantization.get_default_qat_qconfig("qnnpack")
    torch.ao.quantization.prepare_qat(net, inplace=True)
    fill_params_by_name(net)
    g = torch.Generator().manual_seed(21)
    with torch.no_grad():
        for _ in range(2):
            net(torch.randn(4, 3, 64, 64, generator=g))
    x = torch.randn(4, 3, 64, 64, generator=g)
    net.eval()
    with torch.no_grad():
        qat_logits = net(x).clone()                       # observers stay on in eval mode: this forward moves their state
        state = {k: v.clone() for k, v in net.state_dict().items()
                 if not (k.endswith(".weight") and v.dim() > 1) and not k.endswith(".bias")}
        qnet = torch.ao.quantization.convert(net, inplace=False)
        int8_logits = qnet(x).clone()
    out = dict(state=state, x=x, qat_logits=qat_logits, int8_logits=int8_logits, engine="qnnpack",
               converted=qdigest_compact(qnet.state_dict()), torch=torch.__version__)
    torch.save(out, os.path.join(HERE, "int8_mbv3.pt"))
    print("int8 mbv3 golden ok: %d converted entries, |qat - int8|max = %.4g of %.4g"
          % (len(out["converted"]), float((qat_logits - int8_logits).abs().max()), float(qat_logits.abs().max())))
    print(qnet.classifier)


if __name__ == "__main__":
    main()
