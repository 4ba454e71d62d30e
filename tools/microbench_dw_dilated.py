"""Dilated depthwise gather kernels (csrc/dw_dilated.cu) against the tuned dilation-1 kernels on the same shape, CUDA events.

    python tools/microbench_dw_dilated.py [N C H k]
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from frostnet_b200 import _lib as L  # noqa: E402


def timed(fn, iters=20):
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
    a = [int(v) for v in sys.argv[1:5]] if len(sys.argv) >= 5 else [64, 960, 19, 3]
    N, C, H, k = a
    dev = "cuda:0"
    xq = torch.randint(0, 256, (N, H, H, C), dtype=torch.uint8, device=dev)
    wq = torch.randint(-128, 128, (k * k, C), dtype=torch.int8, device=dev)
    zp_a = torch.tensor([117], dtype=torch.int32, device=dev)
    zp_w = torch.tensor([0], dtype=torch.int32, device=dev)
    ws, xs = torch.tensor([0.02], device=dev), torch.tensor([0.05], device=dev)
    acc = torch.empty((N, H, H, C), dtype=torch.int32, device=dev)
    stats = torch.zeros(C * 8, dtype=torch.int32, device=dev)
    dz = torch.randn(N, H, H, C, device=dev)
    dx = torch.empty_like(dz)
    dwq = torch.empty((k * k, C), device=dev)
    st = L.stream(xq.device)
    L.call("frost_stats_reset", stats.data_ptr(), C, st)
    rows = []
    for name, d in (("dilation 1 (tuned kernels)", 1), ("dilation 2 (gather kernels)", 2)):
        if d == 1:
            f = lambda: L.call("frost_dw_conv_forward", xq.data_ptr(), C, zp_a.data_ptr(), wq.data_ptr(), zp_w.data_ptr(), N, H, H, C, k, 1,
                               acc.data_ptr(), stats.data_ptr(), st)
            g = lambda: L.call("frost_dw_dgrad", dz.data_ptr(), wq.data_ptr(), ws.data_ptr(), zp_w.data_ptr(), N, H, H, C, k, 1, dx.data_ptr(), 0, st)
            w = lambda: L.call("frost_dw_wgrad", dz.data_ptr(), xq.data_ptr(), C, xs.data_ptr(), zp_a.data_ptr(), N, H, H, C, k, 1, dwq.data_ptr(), st)
        else:
            f = lambda: L.call("frost_dw_conv_forward_dilated", xq.data_ptr(), C, zp_a.data_ptr(), wq.data_ptr(), zp_w.data_ptr(), N, H, H, C, k,
                               1, d, d * (k - 1) // 2, acc.data_ptr(), stats.data_ptr(), st)
            g = lambda: L.call("frost_dw_dgrad_dilated", dz.data_ptr(), wq.data_ptr(), ws.data_ptr(), zp_w.data_ptr(), N, H, H, C, k, 1, d,
                               d * (k - 1) // 2, dx.data_ptr(), 0, st)
            w = lambda: L.call("frost_dw_wgrad_dilated", dz.data_ptr(), xq.data_ptr(), C, xs.data_ptr(), zp_a.data_ptr(), N, H, H, C, k, 1, d,
                               d * (k - 1) // 2, dwq.data_ptr(), st)
        rows.append((name, timed(f), timed(g), timed(w)))
    for name, tf, tg, tw in rows:
        print("N=%d C=%d %dx%d k=%d  %-30s forward %7.1f us  dgrad %7.1f us  wgrad %7.1f us" % (N, C, H, H, k, name, tf, tg, tw))


if __name__ == "__main__":
    main()
