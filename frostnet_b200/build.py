"""Build libfrost_b200.so (hand-written sm_100a CUDA behind the C ABI of include/frost_b200.h).

In-tree build with nvcc only (no cmake, no torch headers): the product is a plain C-ABI shared
library that links the CUDA runtime statically, so it loads anywhere a driver exists and travels to
the GPU box with the repo snapshot.
"""
import glob
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_DIR = os.path.join(HERE, "_lib")
LIB_PATH = os.path.join(LIB_DIR, "libfrost_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a"]
CFLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "--use_fast_math=false"]
EXTRA = os.environ.get("FROST_NVCC_FLAGS", "").split()      # experiments only (e.g. -DFROST_NO_NC)
if "--trace" in sys.argv:                                   # launch-timeline stamps in the fused kernels (tools/trace_fused.py)
    EXTRA.append("-DFROST_TRACE")


def _sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def _fingerprint():
    h = hashlib.sha256()
    h.update(" ".join(EXTRA).encode())
    for f in _sources() + sorted(glob.glob(os.path.join(CSRC, "*.cuh"))) + [
            os.path.join(os.path.dirname(HERE), "include", "frost_b200.h")]:
        with open(f, "rb") as fh:
            h.update(f.encode())
            h.update(fh.read())
    return h.hexdigest()


def build_library(force=False, verbose=False):
    """Compile every .cu under csrc/ for sm_100a and link libfrost_b200.so.  Returns its path."""
    os.makedirs(LIB_DIR, exist_ok=True)
    stamp = os.path.join(LIB_DIR, "build.stamp")
    fp = _fingerprint()
    if not force and os.path.exists(LIB_PATH) and os.path.exists(stamp) and open(stamp).read() == fp:
        return LIB_PATH
    objs = []
    procs = []
    for src in _sources():
        obj = os.path.join(LIB_DIR, os.path.basename(src)[:-3] + ".o")
        cmd = [NVCC] + ARCH_FLAGS + [f for f in CFLAGS if not f.startswith("--use_fast_math")] + EXTRA + ["-c", src, "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
            print(" ".join(cmd))
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
        objs.append(obj)
    for src, p in procs:
        out, _ = p.communicate()
        if verbose and out:
            print(out.decode())
        if p.returncode != 0:
            raise RuntimeError("nvcc failed on %s:\n%s" % (src, out.decode()))
    cmd = [NVCC] + ARCH_FLAGS + ["-shared", "-o", LIB_PATH] + objs + ["-cudart", "static"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n" + r.stdout.decode())
    with open(stamp, "w") as fh:
        fh.write(fp)
    return LIB_PATH


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
