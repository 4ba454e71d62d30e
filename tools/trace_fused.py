"""Timeline of ONE fused 1x1 launch (SM-clock stamps written by CTA (0,0), frost_debug_set_trace): where the fixed
cost of a small layer goes.

    python -m frostnet_b200.build --force --trace      # the stamps are compiled out of the normal library
    python tools/trace_fused.py [M K cout]

stamps: 0 entry | 1 barriers + TMEM alloc issued | 2 setup sync passed (predecessor complete) | 3 first TMA issued |
4 weights landed | 5 first activation tile landed | 6 first accumulator ready | 7 statistics loop done | 8 at the grid barrier |
9 grid barrier passed | 10 BN / qparams finalised | 11 first accumulator of phase B ready | 12 role done | 18 after teardown sync
backward: 13 coefficients loaded | 14 mask interval found | 15 first dy chunk requested | 16 first accumulator ready | 17 tiles done
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
    M, K, cout = (int(v) for v in args[:3]) if len(args) >= 3 else (12544, 192, 48)
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
    stamps = torch.zeros(32, dtype=torch.int64, device=dev)

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

    mhz = 1965.0
    runs = [("forward", fwd), ("bwd_reduce", bwd("frost_pw_fused_bwd_reduce")), ("bwd_apply", bwd("frost_pw_fused_bwd_apply"))]
    for name, fn in runs:
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        L.call("frost_debug_set_trace", stamps.data_ptr())
        stamps.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        # back to back, like inside a step: the second launch is the traced steady-state one
        fn()
        stamps.zero_()
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        L.call("frost_debug_set_trace", None)
        s = stamps.cpu().tolist()
        t0 = s[0]
        line = " ".join("%d:%.2f" % (i, (v - t0) / mhz) for i, v in enumerate(s) if v)
        print("%-11s M=%d K=%d cout=%d  event %.1f us | us since entry: %s" % (name, M, K, cout, e0.elapsed_time(e1) * 1e3, line))


if __name__ == "__main__":
    main()
