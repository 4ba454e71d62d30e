"""Micro-benchmark of frost_bn_backward / frost_bnq_apply on one layer shape (CUDA events, L2 flushed)."""
import ctypes as C
import sys
import torch
sys.path.insert(0, ".")
from frostnet_b200 import _lib as L

dev = "cuda:0"
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, iters=10):
    ts = []
    for _ in range(iters):
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


for (M, Cc) in [(50176, 312), (12544, 1440), (200704, 168), (802816, 96), (3211264, 96)]:
    st = torch.cuda.current_stream().cuda_stream
    dy = torch.randn(M, Cc, device=dev)
    acc = torch.randint(-5000, 5000, (M, Cc), dtype=torch.int32, device=dev)
    f = lambda n=Cc, v=1.0: torch.full((n,), v, device=dev)
    A, B, meanI, kf, gamma, sf = f(v=1e-3), f(v=0.1), f(v=0.0), f(v=1e-3), f(), f()
    one = torch.ones(1, device=dev); zp = torch.zeros(1, dtype=torch.int32, device=dev); sc = torch.full((1,), 0.05, device=dev)
    sums = torch.zeros(2 * Cc, dtype=torch.float64, device=dev); coef = torch.zeros(3 * Cc, device=dev)
    dz = torch.empty(M, Cc, device=dev); dzl = torch.empty(M, Cc, dtype=torch.bfloat16, device=dev)
    o = [torch.zeros(Cc, device=dev) for _ in range(3)]
    b = L.BnBackwardArgs()
    b.dy, b.acc, b.M, b.C, b.relu = dy.data_ptr(), acc.data_ptr(), M, Cc, 1
    b.A, b.B, b.mean_I, b.kfac = A.data_ptr(), B.data_ptr(), meanI.data_ptr(), kf.data_ptr()
    b.gamma, b.sf, b.x_scale, b.w_scale = gamma.data_ptr(), sf.data_ptr(), one.data_ptr(), one.data_ptr()
    b.out_scale, b.out_zp, b.eps = sc.data_ptr(), zp.data_ptr(), 1e-5
    b.sums, b.coef, b.dz, b.dz_lo, b.dz_format = sums.data_ptr(), coef.data_ptr(), dz.data_ptr(), dzl.data_ptr(), 0
    b.dgamma_bn, b.dbeta, b.dsf_bn = o[0].data_ptr(), o[1].data_ptr(), o[2].data_ptr()
    t_bn = timeit(lambda: L.call("frost_bn_backward", C.byref(b), st))
    q = torch.empty(M, Cc, dtype=torch.uint8, device=dev)
    t_q = timeit(lambda: L.call("frost_bnq_apply", acc.data_ptr(), 0, M, Cc, A.data_ptr(), B.data_ptr(), 1, sc.data_ptr(), zp.data_ptr(), q.data_ptr(), st))
    t_copy = timeit(lambda: dz.copy_(dy))
    n = M * Cc
    print("M=%d C=%d  elems %.1fM | bn_backward %.1f us (%.0f GB/s of 20B/elt) | bnq %.1f us (%.0f GB/s of 5B/elt) | torch copy %.1f us (%.0f GB/s)" % (
        M, Cc, n / 1e6, t_bn, 20 * n / t_bn / 1e3, t_q, 5 * n / t_q / 1e3, t_copy, 8 * n / t_copy / 1e3))
