"""Per-op micro-benchmark on FrostNet-L bs=256 layer shapes (CUDA events on the launch stream, L2 flushed
between iterations), with sweeps of the FROST_TUNE_* launch-shape knobs.

usage (GPU box):  python tools/microbench_ops.py [--quick] > gpurun_out/microbench.txt
Prints one line per (op, shape, knob value): median us, algorithmic GB/s.
"""
import ctypes as C
import sys

import torch

sys.path.insert(0, ".")
from frostnet_b200 import _lib as L  # noqa: E402

dev = "cuda:0"
N = 256
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
T_DW_FWD, T_DW_WGRAD, T_BN_RED, T_BN_CGB, T_STEM_FWD, T_STEM_WGRAD, T_DW_DGRAD, T_PDL, T_BN_RED_UNROLL, T_BN_APPLY_UNROLL, T_BNQ_UNROLL, T_DW_FWD_TILED, T_DW_DGRAD_TILED = range(13)
QUICK = "--quick" in sys.argv


def st():
    return torch.cuda.current_stream().cuda_stream


def timeit(fn, iters=7):
    ts = []
    fn()
    for _ in range(iters):
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


def tune(which, value):
    rc = L.load().frost_set_tunable(which, value)
    assert rc == 0


def sweep(name, shape, fn, nbytes, knob, values):
    out = []
    for v in values:
        tune(knob, v)
        t = timeit(fn)
        out.append("%s=%d: %.1f us (%.0f GB/s)" % ("knob", v, t, nbytes / t / 1e3))
    tune(knob, 0)
    print("%-16s %-28s %s" % (name, shape, " | ".join(out)), flush=True)


def i32(v):
    return torch.tensor([v], dtype=torch.int32, device=dev)


def f32(v):
    return torch.tensor([v], dtype=torch.float32, device=dev)


def bench_dw(H, W, Cc, k, s):
    pad = (k - 1) // 2
    Ho, Wo = (H + 2 * pad - k) // s + 1, (W + 2 * pad - k) // s + 1
    xq = torch.randint(0, 256, (N, H, W, Cc), dtype=torch.uint8, device=dev)
    wq = torch.randint(-128, 128, (k * k, Cc), dtype=torch.int8, device=dev)
    za, zw = i32(3), i32(0)
    acc = torch.empty(N, Ho, Wo, Cc, dtype=torch.int32, device=dev)
    stats = torch.zeros(Cc * 32, dtype=torch.uint8, device=dev)
    L.call("frost_stats_reset", stats.data_ptr(), Cc, st())
    shape = "dw %dx%d C=%d k%d s%d" % (H, W, Cc, k, s)
    nin, nout = N * H * W * Cc, N * Ho * Wo * Cc
    sweep("dw_fwd", shape, lambda: L.call("frost_dw_conv_forward", xq.data_ptr(), Cc, za.data_ptr(), wq.data_ptr(), zw.data_ptr(),
                                          N, H, W, Cc, k, s, acc.data_ptr(), stats.data_ptr(), st()),
          nin + 4 * nout, T_DW_FWD_TILED, [1, 2])
    dz = torch.randn(N, Ho, Wo, Cc, device=dev)
    dwq = torch.empty(k * k, Cc, device=dev)
    sa = f32(0.02)
    sweep("dw_wgrad", shape, lambda: L.call("frost_dw_wgrad", dz.data_ptr(), xq.data_ptr(), Cc, sa.data_ptr(), za.data_ptr(),
                                            N, H, W, Cc, k, s, dwq.data_ptr(), st()),
          nin + 4 * nout, T_DW_WGRAD, [3, 4, 6])
    dx = torch.empty(N, H, W, Cc, device=dev)
    sw = f32(0.01)
    sweep("dw_dgrad", shape, lambda: L.call("frost_dw_dgrad", dz.data_ptr(), wq.data_ptr(), sw.data_ptr(), zw.data_ptr(),
                                            N, H, W, Cc, k, s, dx.data_ptr(), 0, st()),
          4 * nin + 4 * nout, T_DW_DGRAD_TILED, [2, 1])            # 2: gather kernel, 1: TMA-filled shared-memory tiles (stride 1; used where C == 32)


def bench_bn(M, Cc):
    dy = torch.randn(M, Cc, device=dev)
    acc = torch.randint(-5000, 5000, (M, Cc), dtype=torch.int32, device=dev)
    f = lambda v=1.0: torch.full((Cc,), v, device=dev)  # noqa: E731
    A, B, meanI, kf, gamma, sf = f(1e-3), f(0.1), f(0.0), f(1e-3), f(), f()
    one, zp, sc = f32(1.0), i32(0), f32(0.05)
    sums = torch.zeros(2 * Cc, dtype=torch.float64, device=dev)
    coef = torch.zeros(3 * Cc, device=dev)
    dz = torch.empty(M, Cc, device=dev)
    dzl = torch.empty(M, Cc, dtype=torch.bfloat16, device=dev)
    o = [torch.zeros(Cc, device=dev) for _ in range(3)]
    b = L.BnBackwardArgs()
    b.dy, b.acc, b.M, b.C, b.relu = dy.data_ptr(), acc.data_ptr(), M, Cc, 1
    b.A, b.B, b.mean_I, b.kfac = A.data_ptr(), B.data_ptr(), meanI.data_ptr(), kf.data_ptr()
    b.gamma, b.sf, b.x_scale, b.w_scale = gamma.data_ptr(), sf.data_ptr(), one.data_ptr(), one.data_ptr()
    b.out_scale, b.out_zp, b.eps = sc.data_ptr(), zp.data_ptr(), 1e-5
    b.sums, b.coef, b.dz, b.dz_lo, b.dz_format = sums.data_ptr(), coef.data_ptr(), dz.data_ptr(), dzl.data_ptr(), 1
    b.dgamma_bn, b.dbeta, b.dsf_bn = o[0].data_ptr(), o[1].data_ptr(), o[2].data_ptr()
    shape = "bn M=%d C=%d" % (M, Cc)
    n = M * Cc
    red = lambda: L.call("frost_bn_backward_reduce", C.byref(b), st())  # noqa: E731
    for unroll in (4, 8):
        tune(T_BN_RED_UNROLL, unroll)
        sweep("bn_red U=%d" % unroll, shape, red, 8 * n, T_BN_RED, [2, 3, 4])
    tune(T_BN_RED_UNROLL, 0)
    sweep("bn_apply", shape, lambda: L.call("frost_bn_backward_apply", C.byref(b), st()), 12 * n, T_BN_APPLY_UNROLL, [1, 2, 4])
    q = torch.empty(M, Cc, dtype=torch.uint8, device=dev)
    sweep("bnq_apply", shape, lambda: L.call("frost_bnq_apply", acc.data_ptr(), 0, M, Cc, A.data_ptr(), B.data_ptr(), 1, sc.data_ptr(),
                                             zp.data_ptr(), q.data_ptr(), Cc, st()), 5 * n, T_BNQ_UNROLL, [2, 4, 8])


def bench_stem():
    H = W = 224
    xq = torch.randint(0, 256, (N, H, W, 3), dtype=torch.uint8, device=dev)
    wq = torch.randint(-128, 128, (32, 3, 3, 3), dtype=torch.int8, device=dev)
    za, zw = i32(114), i32(0)
    acc = torch.empty(N, 112, 112, 32, dtype=torch.int32, device=dev)
    stats = torch.zeros(32 * 32, dtype=torch.uint8, device=dev)
    L.call("frost_stats_reset", stats.data_ptr(), 32, st())
    nb = N * H * W * 3 + 4 * N * 112 * 112 * 32
    sweep("stem_fwd", "224x224x3 -> 32", lambda: L.call("frost_stem_conv_forward", xq.data_ptr(), za.data_ptr(), wq.data_ptr(),
                                                        zw.data_ptr(), N, H, W, 3, 32, 3, 2, 1, acc.data_ptr(), stats.data_ptr(), st()),
          nb, T_STEM_FWD, [4, 8, 16])
    dz = torch.randn(N, 112, 112, 32, device=dev)
    dwq = torch.empty(32, 3, 3, 3, device=dev)
    sa = f32(0.02)
    sweep("stem_wgrad", "224x224x3 -> 32", lambda: L.call("frost_stem_wgrad", dz.data_ptr(), xq.data_ptr(), sa.data_ptr(), za.data_ptr(),
                                                          N, H, W, 3, 32, 3, 2, 1, dwq.data_ptr(), st()),
          nb, T_STEM_WGRAD, [3, 4, 5])


def bench_pw(M, K, cout):
    xq = torch.randint(0, 256, (M, K), dtype=torch.uint8, device=dev)
    wq = torch.randint(-128, 128, (cout, K), dtype=torch.int8, device=dev)
    za, zw = i32(3), i32(0)
    wsum = wq.int().sum(1).int().contiguous()
    acc = torch.empty(M, cout, dtype=torch.int32, device=dev)
    stats = torch.zeros(cout * 32, dtype=torch.uint8, device=dev)
    L.call("frost_stats_reset", stats.data_ptr(), cout, st())
    t = timeit(lambda: L.call("frost_pw_conv_forward", xq.data_ptr(), za.data_ptr(), wq.data_ptr(), zw.data_ptr(), wsum.data_ptr(),
                              M, K, cout, acc.data_ptr(), stats.data_ptr(), st()))
    nb = M * K + 4 * M * cout
    print("%-16s %-28s %.1f us (%.0f GB/s)" % ("pw_fwd", "M=%d K=%d cout=%d" % (M, K, cout), t, nb / t / 1e3), flush=True)
    hi = torch.randn(M, cout, device=dev).to(torch.bfloat16)
    lo = torch.zeros(M, cout, dtype=torch.bfloat16, device=dev)
    wt = torch.randint(-128, 128, (K, cout), device=dev).to(torch.bfloat16)
    sw = f32(0.01)
    dx = torch.empty(M, K, device=dev)
    t = timeit(lambda: L.call("frost_pw_dgrad_tc", hi.data_ptr(), lo.data_ptr(), wt.data_ptr(), sw.data_ptr(), M, K, cout,
                              dx.data_ptr(), 0, st()))
    nb = 4 * M * cout + 4 * M * K
    print("%-16s %-28s %.1f us (%.0f GB/s)" % ("pw_dgrad", "M=%d K=%d cout=%d" % (M, K, cout), t, nb / t / 1e3), flush=True)
    dwq = torch.empty(cout, K, device=dev)
    sa = f32(0.02)
    t = timeit(lambda: L.call("frost_pw_wgrad_tc", hi.data_ptr(), lo.data_ptr(), xq.data_ptr(), K, sa.data_ptr(), za.data_ptr(), M, K,
                              cout, dwq.data_ptr(), st()))
    nb = 4 * M * cout + M * K
    print("%-16s %-28s %.1f us (%.0f GB/s)" % ("pw_wgrad", "M=%d K=%d cout=%d" % (M, K, cout), t, nb / t / 1e3), flush=True)


if __name__ == "__main__":
    only = [a for a in sys.argv[1:] if not a.startswith("--")]
    want = lambda k: not only or k in only  # noqa: E731
    ca = torch.empty(1 << 28, dtype=torch.float32, device=dev)
    cb = torch.empty(1 << 28, dtype=torch.float32, device=dev)
    print("torch copy 1 GiB: %.0f GB/s" % (2 * (1 << 30) / timeit(lambda: cb.copy_(ca)) / 1e3))
    del ca, cb
    if want("stem"):
        bench_stem()
    if want("dw"):
        for shp in [(112, 112, 32, 3, 1), (112, 112, 96, 3, 2), (56, 56, 72, 3, 1), (56, 56, 144, 5, 2), (28, 28, 168, 3, 1),
                    (28, 28, 336, 5, 2), (14, 14, 624, 5, 1), (14, 14, 360, 3, 1), (14, 14, 864, 5, 2), (7, 7, 1440, 5, 1)]:
            bench_dw(*shp)
            torch.cuda.empty_cache()
    if want("bn"):
        for shp in [(3211264, 96), (3211264, 16), (802816, 72), (200704, 168), (50176, 360), (12544, 1440), (12544, 192)]:
            bench_bn(*shp)
            torch.cuda.empty_cache()
    if want("pw"):
        for shp in [(3211264, 16, 96), (3211264, 32, 16), (802816, 96, 24), (802816, 24, 144), (200704, 56, 168), (50176, 360, 96),
                    (50176, 120, 360), (12544, 240, 1440), (12544, 1440, 192), (12544, 320, 1280)]:
            bench_pw(*shp)
            torch.cuda.empty_cache()
