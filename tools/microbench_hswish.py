"""CUDA-event timing of the QAT hard-swish (csrc/hswish.cu) against its algorithmic bytes, next to the same chain done
with one torch op per reference op (what the reference's autograd graph launches).

    python tools/microbench_hswish.py [n_elements]
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import frostnet_b200 as F  # noqa: E402


def timed(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 256 * 16 * 112 * 112
    dev = "cuda:0"
    net = torch.nn.Sequential(F.QuantStub(), F.Hswish(True))
    F.attach_fake_quant(net)
    net.to(dev).train()
    x0 = torch.randn(n, device=dev) * 3
    xq = net[0](x0)
    qp = xq._frost_qparams                    # (a detach() makes a new tensor object: the attribute stays behind)
    xq = xq.detach()
    hs = net[1]

    def fwd():
        x = xq.requires_grad_(True)
        x._frost_qparams = qp
        return hs(x), x
    y, x = fwd()
    dy = torch.randn_like(y)
    us_f = timed(lambda: fwd())
    us_fb = timed(lambda: fwd()[0].backward(dy))
    print("hswish forward   n=%d  %.1f us  %.0f GB/s (algorithmic 10 B/element: 4+1 index pass, 1+4 table pass)" % (n, us_f, 10 * n / us_f / 1e3))
    print("hswish fwd+bwd   n=%d  %.1f us  backward alone %.1f us  %.0f GB/s (9 B/element)" % (n, us_fb, us_fb - us_f, 9 * n / max(us_fb - us_f, 1e-9) / 1e3))
    # the reference's op-per-op graph with torch's own CUDA kernels (fused_moving_avg_obs_fake_quant x2, add, hardtanh, mul, mul)
    fa = torch.ao.quantization.get_default_qat_qconfig("qnnpack").activation().to(dev)
    fb = torch.ao.quantization.get_default_qat_qconfig("qnnpack").activation().to(dev)

    def ref():
        x = xq.detach().requires_grad_(True)
        out = fb(x * fa(torch.nn.functional.relu6(x + 3.0))) * (1 / 6)
        out.backward(dy)
    us_ref = timed(ref)
    print("torch op-per-op fwd+bwd  %.1f us  (%.1fx)" % (us_ref, us_ref / us_fb))


if __name__ == "__main__":
    main()
