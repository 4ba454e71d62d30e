"""CPU: host logic of the reference-facing surface and the C-ABI library (no compute calls)."""
import ctypes
import os
import re

import pytest
import torch

from util import build_model_from_golden, load_golden

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_builds_loads_and_exports_every_declared_symbol():
    from frostnet_b200 import _lib, build
    path = build.build_library()
    assert os.path.exists(path)
    lib = ctypes.CDLL(path)
    header = open(os.path.join(ROOT, "include", "frost_b200.h")).read()
    declared = sorted(set(re.findall(r"^(?:int|int64_t|const char\*)\s+(frost_\w+)\s*\(", header, flags=re.M)))
    assert len(declared) >= 25
    for name in declared:
        assert hasattr(lib, name), "libfrost_b200.so does not export %s" % name
    assert sorted(declared) == _lib.EXPORTED_SYMBOLS, "ctypes binding and header disagree"
    lib.frost_abi_version.restype = ctypes.c_int
    assert lib.frost_abi_version() == 2
    _lib.load()
    assert _lib.launch_count() == 0            # nothing launched on a CPU-only box


def test_struct_layouts_match_header_sizes():
    from frostnet_b200 import _lib
    assert ctypes.sizeof(_lib.ChanStats) == 32
    assert ctypes.sizeof(_lib.FQ) == 32
    assert ctypes.sizeof(_lib.OptChunk) == 8
    assert ctypes.sizeof(_lib.OptTensor) % 8 == 0


def test_state_dict_layout_matches_reference_after_prepare():
    g = load_golden("net_small035.pt")
    model = build_model_from_golden(g, torch.device("cpu"))
    sd = model.state_dict()
    assert list(sd.keys()) == list(g["sd0"].keys())
    for k, v in g["sd0"].items():
        assert sd[k].dtype == v.dtype and sd[k].shape == v.shape, k
        assert torch.equal(sd[k], v), k


def test_parameter_identity_survives_fuse_and_prepare():
    import frostnet_b200 as F
    model = F.frostnet_quant_small_0_35()
    ids = {n: id(p) for n, p in model.named_parameters()}
    before = [id(p) for p in model.parameters()]
    model.train()
    model.fuse_model()
    F.prepare_qat(model)
    after = [id(p) for p in model.parameters()]
    assert before == after                                   # S1: optimizer state carries over
    assert len(ids) == len(list(model.named_parameters()))
    depthwise = [n for n, p in model.named_parameters() if p.dim() == 4 and p.shape[1] == 1]
    assert len(depthwise) == 14                              # train.py:129-137 identifies dw by shape[1]==1


def test_large_structure_known_answers():
    import frostnet_b200 as F
    from frostnet_b200 import qat
    model = F.frostnet_quant_large_1_0()
    assert sum(p.numel() for p in model.parameters()) == 5807056
    model.fuse_model()
    F.prepare_qat(model)
    assert len(list(model.parameters())) == 209
    n_fq = sum(1 for m in model.modules() if isinstance(m, qat.FrostFakeQuantize))
    assert n_fq == 177                                       # K8
    n_cbr = sum(1 for m in model.modules() if isinstance(m, qat.FrostConvBn2d) and m.relu)
    n_cb = sum(1 for m in model.modules() if isinstance(m, qat.FrostConvBn2d) and not m.relu)
    assert (n_cbr, n_cb) == (51, 18)


def test_factories_and_errors():
    import frostnet_b200 as F
    names = [n for n in dir(F) if n.startswith("frostnet_") and n != "frostnet_features"]
    assert len(names) == 30
    assert sum(p.numel() for p in F.frostnet_small_1_0().parameters()) == 4858776
    assert sum(p.numel() for p in F.frostnet_base_1_0().parameters()) == 5001712
    with pytest.raises(ValueError):
        F.FrostNet(mode="huge")
    with pytest.raises(ValueError):
        F.prepare_qat(F.frostnet_small_0_35())               # quantized=False
    with pytest.raises(ValueError):
        F.QSGD([torch.nn.Parameter(torch.zeros(1))], lr=-1.0)
    with pytest.raises(ValueError):
        F.QSGD([torch.nn.Parameter(torch.zeros(1))], lr=0.1, nesterov=True)
    with pytest.raises(ValueError):
        F.QAdam([torch.nn.Parameter(torch.zeros(1))], betas=(1.0, 0.9))


def test_float_model_matches_reference_style_forward():
    """Before prepare the model is a plain float network (FP warm-up phase of StatAssist)."""
    import frostnet_b200 as F
    torch.manual_seed(0)
    model = F.frostnet_quant_small_0_35(drop_rate=0.0).eval()
    x = torch.randn(2, 3, 64, 64)
    a = model(x)
    model.fuse_model()
    b = model(x)
    assert a.shape == (2, 1000)
    assert torch.allclose(a, b, atol=1e-5)


def test_qat_path_has_no_cpu_fallback():
    g = load_golden("net_small035.pt")
    model = build_model_from_golden(g, torch.device("cpu"))
    with pytest.raises(RuntimeError):
        model(g["xs"][0])
    blk = model.layer1[0]
    with pytest.raises(RuntimeError):
        blk(torch.zeros(1, blk.in_channels, 8, 8))


def test_optimizer_surface():
    import frostnet_b200 as F

    class Args:
        learning_rate, weight_decay, nesterov, clip_by, toss_coin, noise_decay, amsgrad = 5e-3, 1e-5, True, 1e-3, True, 1e-2, False
    p = [torch.nn.Parameter(torch.zeros(3))]
    for name, cls in (("QSGD", F.QSGD), ("QRMS", F.QRMSprop), ("QAdam", F.QAdam), ("QAdamW", F.QAdamW)):
        opt = F.get_optimizer(name, p, Args)
        assert isinstance(opt, cls) and opt.is_warmup is True
        assert opt.param_groups[0]["lr"] == 5e-3
    assert isinstance(F.get_optimizer("SGD", p, Args), torch.optim.SGD)
    opt = F.get_optimizer("QSGD", p, Args)
    p[0].grad = torch.ones(3)
    with pytest.raises(RuntimeError):
        opt.step()                                            # CPU tensors: no fallback


def test_feature_backbone_surface_and_state_dict():
    import frostnet_b200 as F
    from frostnet_b200 import frostnet_features as FF
    g = load_golden("features_small035.pt")
    model = FF.FrostNet(mode=g["mode"], width_mult=g["width_mult"], quantized=True)
    model.init_weights('')
    feats = model(torch.randn(1, 3, 64, 64))
    assert [tuple(f.shape[1:]) for f in feats] == [tuple(f.shape[1:]) for f in g["feats"]]
    model.train()
    model.fuse_model()
    F.prepare_qat(model)
    sd = model.state_dict()
    assert list(sd.keys()) == list(g["sd0"].keys())
    for k, v in g["sd0"].items():
        assert sd[k].dtype == v.dtype and sd[k].shape == v.shape, k
    model.load_state_dict(g["sd0"], strict=True)
    with pytest.raises(RuntimeError):
        model(g["x"])                                   # prepared model on CPU tensors: no fallback
    large = FF.FrostNet(mode="large", width_mult=1.0)
    assert [f.shape[1] for f in large(torch.randn(1, 3, 64, 64))] == [24, 40, 96, 320]
    with pytest.raises(ValueError):
        FF.FrostNet(mode="tiny")


def test_observer_toggle_protocol():
    """torch.ao.quantization.disable_observer / enable_observer reach the fake-quant state modules."""
    import torch.ao.quantization as taq
    from frostnet_b200 import qat
    g = load_golden("net_small035.pt")
    model = build_model_from_golden(g, torch.device("cpu"))
    model.apply(taq.disable_observer)
    fqs = [m for m in model.modules() if isinstance(m, qat.FrostFakeQuantize)]
    assert fqs and all(not f._observe and int(f.observer_enabled) == 0 for f in fqs)
    model.apply(taq.enable_observer)
    assert all(f._observe and int(f.observer_enabled) == 1 for f in fqs)
    with pytest.raises(RuntimeError):
        model.apply(taq.disable_fake_quant)              # the engine computes on indices: cannot be disabled


def test_launch_knobs_roundtrip_without_a_gpu():
    """frost_set_tunable / frost_get_tunable are host-only: defaults, overrides, reset with 0, error reporting."""
    from frostnet_b200 import _lib
    lib = _lib.load()
    header = open(os.path.join(ROOT, "include", "frost_b200.h")).read()
    count = int(re.search(r"FROST_TUNE_COUNT\s*=\s*(\d+)", header).group(1))
    defaults = [lib.frost_get_tunable(k) for k in range(count)]
    assert all(v > 0 for v in defaults)
    try:
        for k in range(count):
            assert lib.frost_set_tunable(k, 7) == 0 and lib.frost_get_tunable(k) == 7
            assert lib.frost_set_tunable(k, 0) == 0 and lib.frost_get_tunable(k) == defaults[k]
        assert lib.frost_set_tunable(count, 1) != 0
        assert b"unknown knob" in lib.frost_last_error()
        assert lib.frost_set_tunable(0, -3) != 0
        assert lib.frost_get_tunable(-1) < 0
    finally:
        for k in range(count):
            lib.frost_set_tunable(k, 0)


def test_sass_obeys_the_pdl_load_rule_and_uses_the_tensor_cores():
    """Static guard for the rule of csrc/common.cuh: kernels launched with programmatic dependent launch must not
    read through the non-coherent path (LDG.E...CONSTANT = ld.global.nc), and the griddepcontrol pair is present.
    Also the evidence that the 1x1 convs run on tcgen05 (UTCIMMA = kind::i8, UTCHMMA = kind::f16)."""
    import shutil
    import subprocess
    from frostnet_b200 import build
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        import pytest
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([cuobjdump, "-sass", build.build_library()], capture_output=True, text=True).stdout
    assert "sm_100a" in sass
    assert not re.search(r"LDG\.E[.\w]*\.CONSTANT", sass), "ld.global.nc in a PDL library (see common.cuh)"
    assert sass.count("ACQBULK") >= 20 and sass.count("PREEXIT") >= 20      # griddepcontrol.wait / launch_dependents
    assert "UTCIMMA" in sass and "UTCHMMA" in sass
    assert "LDGSTS.E.BYPASS.128" in sass                                    # 16-byte cp.async.cg (L2 only)
