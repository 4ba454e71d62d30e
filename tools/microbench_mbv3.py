"""One QAT training step (forward + backward, dropout on) of the reference's MobileNetV3 on the per-module executor, timed
with CUDA events; and the same float network through torch's own fp32 CUDA kernels for scale.

    python tools/microbench_mbv3.py [large|small] [batch]
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import frostnet_b200 as F                      # noqa: E402
from frostnet_b200 import _lib as L            # noqa: E402
from frostnet_b200 import mobilenetv3 as M     # noqa: E402


def timed(step, iters=5, warm=3):
    for _ in range(warm):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        step()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    mode = sys.argv[1] if len(sys.argv) > 1 else "large"
    bs = int(sys.argv[2]) if len(sys.argv) > 2 else 128
    dev = "cuda:0"
    torch.manual_seed(0)
    x = torch.randn(bs, 3, 224, 224, device=dev)
    t = torch.randint(0, 1000, (bs,), device=dev)

    fl = M.get_mobilenet_v3(mode, 1.0).to(dev).train()

    def float_step():
        fl.zero_grad(set_to_none=True)
        torch.nn.functional.cross_entropy(fl(x), t).backward()
    ms_f = timed(float_step)

    net = M.get_mobilenet_v3(mode, 1.0)
    net.train()
    net.fuse_model()
    F.attach_fake_quant(net)
    net.to(dev)

    def qat_step():
        net.zero_grad(set_to_none=True)
        torch.nn.functional.cross_entropy(net(x), t).backward()
    n0 = L.load().frost_launch_count()
    ms_q = timed(qat_step)
    launches = (L.load().frost_launch_count() - n0) / 8
    print("MobileNetV3-%s bs=%d 224x224: QAT step on the per-module executor %.1f ms (%.0f images/s, %.0f native launches/step); "
          "float fp32 step on torch's kernels %.1f ms" % (mode, bs, ms_q, bs / ms_q * 1e3, launches, ms_f))


if __name__ == "__main__":
    main()
