"""Probe: how fast is the TMA-tiled depthwise dgrad on SMALL planes when its boxes are contiguous?  C = 32 makes every pixel
one 128-byte line; N is scaled so that the element count equals a wide FrostNet layer (14x14x624, 7x7x1440 at bs=256).
knob 2 = gather kernel, knob 1 = tiles.   python tools/probes/dw_dgrad_contig_probe.py"""
import sys
import torch
sys.path.insert(0, ".")
from frostnet_b200 import _lib as L  # noqa: E402

dev = "cuda:0"
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, iters=7):
    ts = []
    fn()
    for _ in range(iters):
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    return sorted(ts)[len(ts) // 2]


for (H, W, k, mult) in [(14, 14, 5, 624 // 32), (7, 7, 5, 45), (28, 28, 3, 5), (56, 56, 3, 2)]:
    N, C = 256 * mult, 32
    dz = torch.randn(N, H, W, C, device=dev)
    wq = torch.randint(-128, 128, (k * k, C), dtype=torch.int8, device=dev)
    sw = torch.tensor([0.01], device=dev)
    zw = torch.zeros(1, dtype=torch.int32, device=dev)
    dx = torch.empty(N, H, W, C, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    out = []
    for knob in (2, 1):
        assert L.load().frost_set_tunable(12, knob) == 0
        t = timeit(lambda: L.call("frost_dw_dgrad", dz.data_ptr(), wq.data_ptr(), sw.data_ptr(), zw.data_ptr(), N, H, W, C, k, 1,
                                  dx.data_ptr(), 0, st))
        out.append("%s %.1f us (%.0f GB/s)" % ("gather" if knob == 2 else "tiles", t, 8 * dz.numel() / t / 1e3))
    L.load().frost_set_tunable(12, 0)
    print("%dx%d k%d, %d images x 32 ch (%.1f M elements): %s" % (H, W, k, N, dz.numel() / 1e6, " | ".join(out)))
