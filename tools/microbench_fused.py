"""Time (CUDA events) the fused 1x1 kernels on one layer shape; used under ncu for the per-kernel captures in profiles/.

    python tools/microbench_fused.py [M K cout] [--iters N]
"""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from frostnet_b200 import _lib as L  # noqa: E402
import test_gpu_fused as T          # noqa: E402


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    iters = 5
    if "--iters" in sys.argv:
        iters = int(sys.argv[sys.argv.index("--iters") + 1])
    M, K, cout = (int(v) for v in args[:3]) if len(args) >= 3 else (3211264, 16, 96)
    dev = "cuda:0"
    ly = T.Layer(K, cout, 0, True, 1)
    ldx, ldq = T._r16(K), T._r16(cout)
    x = torch.randint(0, 256, (M, ldx), dtype=torch.uint8, device=dev)
    x_zp = torch.tensor([7], dtype=torch.int32, device=dev)
    q = torch.empty(M, ldq, dtype=torch.uint8, device=dev)
    dy = torch.randn(M, cout, device=dev)
    hi, lo = (torch.empty(M, cout, dtype=torch.bfloat16, device=dev) for _ in range(2))
    st = torch.cuda.current_stream().cuda_stream
    bar = torch.zeros(1, dtype=torch.int32, device=dev)

    def fwd():
        f = L.PwFusedFwdArgs()
        f.op, f.bn = ly.operands(x, M, ldx, x_zp), ly.fin_args(M, True, True)
        bar.zero_()
        f.grid_barrier, f.q, f.ldq = bar.data_ptr(), q.data_ptr(), ldq
        L.call("frost_pw_fused_forward", C.byref(f), st)

    def bwd(name):
        def run():
            fb = L.PwFusedBwdArgs()
            fb.op, fb.bn = ly.operands(x, M, ldx, x_zp), ly.bwd_args(M, dy, None, hi, lo, False)
            L.call(name, C.byref(fb), st)
        return run

    wt = ly.wq.float().t().contiguous().to(torch.bfloat16)
    dx = torch.empty(M, K, device=dev)
    dwq = torch.empty(cout, K, device=dev)

    def chain():
        ch = L.PwChainArgs()
        ch.op, ch.bn = ly.operands(x, M, ldx, x_zp), ly.bwd_args(M, dy, None, hi, lo, False)
        ch.wt_bf16, ch.dx, ch.accumulate, ch.dwq = wt.data_ptr(), dx.data_ptr(), 0, dwq.data_ptr()
        L.call("frost_pw_chain_backward", C.byref(ch), st)

    def dgrad():
        L.call("frost_pw_dgrad_tc", hi.data_ptr(), lo.data_ptr(), wt.data_ptr(), ly.w_scale.data_ptr(), M, K, cout, dx.data_ptr(), 0, st)

    def wgrad():
        L.call("frost_pw_wgrad_tc", hi.data_ptr(), lo.data_ptr(), x.data_ptr(), ldx, ly.x_scale.data_ptr(), x_zp.data_ptr(), M, K, cout,
               dwq.data_ptr(), st)

    cases = [("fused_forward", fwd, 2 * M * K + M * cout),
             ("fused_bwd_reduce", bwd("frost_pw_fused_bwd_reduce"), M * K + 4 * M * cout),
             ("fused_bwd_apply", bwd("frost_pw_fused_bwd_apply"), M * K + 8 * M * cout),
             ("dgrad_tc", dgrad, 4 * M * cout + 4 * M * K), ("wgrad_tc", wgrad, 4 * M * cout + M * K)]
    if L.load().frost_pw_chain_supported(K, cout):
        cases.append(("chain_backward", chain, M * K + 4 * M * cout + 4 * M * K))
    for name, fn, byt in cases:
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / iters
        print("%-18s M=%d K=%d cout=%d  %.1f us  %.0f GB/s  %.0f elems/us" % (name, M, K, cout, us, byt / us / 1e3, M * cout / us), flush=True)


if __name__ == "__main__":
    main()
