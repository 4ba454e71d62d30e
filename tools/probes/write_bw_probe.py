"""Write-only / read-only / copy DRAM bandwidth of the box (torch ops, CUDA events): the denominators for a kernel whose
traffic is mostly stores (the quantise phase of the fused forward writes 6x what it reads)."""
import torch

dev = "cuda:0"
n = 1 << 30
a = torch.empty(n, dtype=torch.uint8, device=dev)
b = torch.empty(n, dtype=torch.uint8, device=dev)
af = a.view(torch.float32)


def timed(fn, bytes_, name, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / iters
    print("%-28s %8.1f us  %6.0f GB/s" % (name, us, bytes_ / us / 1e3))


timed(lambda: a.zero_(), n, "memset 1 GiB (write only)")
timed(lambda: af.fill_(1.5), n, "fill fp32 1 GiB (write only)")
timed(lambda: b.copy_(a), 2 * n, "copy 1 GiB (read + write)")
timed(lambda: af.sum(), n, "sum fp32 1 GiB (read only)")
half = n // 2
timed(lambda: a[:half].zero_(), half, "memset 512 MiB")
