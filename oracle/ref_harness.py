"""TEST / BENCH INFRASTRUCTURE ONLY - runs the UNMODIFIED reference files on the host CPU.

The reference (clovaai/frostnet) is pure Python wiring of torch: three files carry the whole hot path
(`frostnet.py`, `frostnet_features.py`, `optimizer.py`).  There is nothing to compile and no
setup.py to pip-install, so "installing the reference" means placing those three files, byte for byte, under
`baseline/_ref/` (git-ignored, NOT gpurun-ignored: it travels to the GPU box like a built .so).  `install()` does
that from `/root/reference` when it exists (the build container); on the GPU box the copies are simply used.

Nothing in `frostnet_b200/` imports this module.  Callers: `bench.py --impl reference` (the reference arm and the
`cpu_baseline` leg, in a process of its own because the shims below patch torch globally) and
`tests/golden/make_golden.py` style checks.

Shims (all three verified there):
  * a stub `timm` package (4 mean/std constants + an identity `register_model`; frostnet.py:4-5);
  * `torch.quantization.fuse_modules = fuse_modules_qat` (the reference fuses in train mode,
    Classification/train.py:171, which torch >= 1.11 only allows through the _qat entry point);
  * `torch.Tensor.cuda = identity` (hard-coded `.cuda()` at optimizer.py:180) - CPU runs only.
"""
import hashlib
import os
import shutil
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SRC = "/root/reference"
REF_DIR = os.path.join(ROOT, "baseline", "_ref")
FILES = ("frostnet.py", "frostnet_features.py", "optimizer.py")


def install(verbose=False):
    """Copy the three reference files into baseline/_ref (no-op without /root/reference).  Returns True if present."""
    if os.path.isdir(REF_SRC):
        os.makedirs(REF_DIR, exist_ok=True)
        for f in FILES:
            src, dst = os.path.join(REF_SRC, f), os.path.join(REF_DIR, f)
            if not os.path.exists(dst) or open(src, "rb").read() != open(dst, "rb").read():
                shutil.copyfile(src, dst)
        with open(os.path.join(REF_DIR, "SHA256SUMS"), "w") as fh:
            for f in FILES:
                fh.write("%s  %s\n" % (hashlib.sha256(open(os.path.join(REF_DIR, f), "rb").read()).hexdigest(), f))
        if verbose:
            print("reference files installed under", REF_DIR)
    return available()


def available():
    return all(os.path.exists(os.path.join(REF_DIR, f)) for f in FILES)


def install_shims():
    import torch
    timm = types.ModuleType("timm")
    data = types.ModuleType("timm.data")
    models = types.ModuleType("timm.models")
    reg = types.ModuleType("timm.models.registry")
    data.IMAGENET_DEFAULT_MEAN = (0.485, 0.456, 0.406)
    data.IMAGENET_DEFAULT_STD = (0.229, 0.224, 0.225)
    data.IMAGENET_INCEPTION_MEAN = (0.5, 0.5, 0.5)
    data.IMAGENET_INCEPTION_STD = (0.5, 0.5, 0.5)
    reg.register_model = lambda f: f
    sys.modules.update({"timm": timm, "timm.data": data, "timm.models": models, "timm.models.registry": reg})
    torch.quantization.fuse_modules = torch.ao.quantization.fuse_modules_qat
    torch.Tensor.cuda = lambda self, *a, **k: self


def load():
    """-> (frostnet module, optimizer module) of the unmodified reference."""
    if not available():
        raise RuntimeError("baseline/_ref is empty: run `python -m oracle.ref_harness` where /root/reference exists")
    install_shims()
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    import importlib
    fn = importlib.import_module("frostnet")
    op = importlib.import_module("optimizer")
    if os.path.dirname(os.path.abspath(fn.__file__)) != REF_DIR:
        raise RuntimeError("`frostnet` resolved to %s, not the reference copy" % fn.__file__)
    return fn, op


class TrainArgs:
    """Classification/setting/train.json"""
    learning_rate, weight_decay, nesterov, clip_by, toss_coin, noise_decay, amsgrad = 5e-3, 1e-5, True, 1e-3, True, 1e-2, False
    momentum = 0.9


def param_groups(model, weight_decay):
    """Classification/train.py:121-137 - one group per tensor, weight decay by shape."""
    groups = []
    for _, p in model.named_parameters():
        if p.dim() == 4:
            wd = 0.0 if p.shape[1] == 1 else weight_decay
        else:
            wd = weight_decay * 0.01
        groups.append({"params": [p], "weight_decay": wd})
    return groups


def qat_images_per_s(batch, steps, warmup, threads=None, fp_warmup_steps=1, fp_batch=8):
    """images/s of the reference's own QAT training step (helper_functions.py:139-143: zero_grad -> model(x) -> CE ->
    backward -> QSGD.step) for frostnet_quant_large_1_0 on the host CPU, after the StatAssist recipe
    (Classification/train.py:149-173: FP warm-up -> is_warmup=False -> fuse -> qnnpack QAT qconfig -> prepare_qat)."""
    import time
    import numpy as np
    import torch
    fn, op = load()
    threads = threads or os.cpu_count() or 1
    torch.set_num_threads(threads)
    torch.manual_seed(1882)
    np.random.seed(1882)
    model = fn.frostnet_quant_large_1_0()
    model.train()
    opt = op.get_optimizer("QSGD", param_groups(model, TrainArgs.weight_decay), TrainArgs)
    crit = torch.nn.CrossEntropyLoss()
    x = torch.randn(batch, 3, 224, 224)
    y = torch.randint(0, 1000, (batch,))

    def step(xb, yb):
        opt.zero_grad()
        loss = crit(model(xb), yb)
        loss.backward()
        opt.step()
        return float(loss)

    for _ in range(fp_warmup_steps):
        step(x[:fp_batch], y[:fp_batch])
    opt.is_warmup = False
    model.fuse_model()
    model.qconfig = torch.ao.quantization.get_default_qat_qconfig("qnnpack")
    torch.ao.quantization.prepare_qat(model, inplace=True)
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        loss = step(x, y)
        if it >= warmup:
            times.append(time.perf_counter() - t0)
    assert loss == loss
    return batch * len(times) / sum(times), threads, sum(times) / len(times)


def c1_small_fp32_ms_per_image(threads=None, iters=30, warmup=5):
    """BASELINE configs[0]: FrostNet-Small fp32 forward, bs=1, 224x224, CPU (Classification/evaluate.py path before
    the QAT conversion: plain float eval-mode forward).  Median ms per image."""
    import statistics
    import time
    import torch
    fn, _ = load()
    threads = threads or os.cpu_count() or 1
    torch.set_num_threads(threads)
    torch.manual_seed(1882)
    model = fn.frostnet_small_1_0().eval()
    x = torch.randn(1, 3, 224, 224)
    ts = []
    with torch.no_grad():
        for i in range(warmup + iters):
            t0 = time.perf_counter()
            model(x)
            if i >= warmup:
                ts.append((time.perf_counter() - t0) * 1e3)
    return statistics.median(ts), threads


if __name__ == "__main__":
    print("installed" if install(verbose=True) else "reference files not available")
