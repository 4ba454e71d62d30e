"""One depthwise layer shape through the three depthwise kernels (for ncu captures): python tools/ncu_dw_one.py H W C k s"""
import sys
sys.argv, args = sys.argv[:1], [int(v) for v in sys.argv[1:6]]
sys.path.insert(0, "tools")
import microbench_ops as mb  # noqa: E402
mb.bench_dw(*args)
