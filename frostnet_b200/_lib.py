"""ctypes binding of libfrost_b200.so (the C ABI declared in include/frost_b200.h).

The product path has NO fallback: if the library is missing or a call fails, a RuntimeError is
raised.  Nothing here imports ``oracle/``.
"""
import ctypes as C
import os

import torch

from . import build as _build

c_p = C.c_void_p


class FQ(C.Structure):
    _fields_ = [("min_val", c_p), ("max_val", c_p), ("scale", c_p), ("zero_point", c_p)]


class ChanStats(C.Structure):
    _fields_ = [("sum", C.c_longlong), ("sq_lo", C.c_ulonglong), ("sq_hi", C.c_ulonglong),
                ("min", C.c_int32), ("max", C.c_int32)]


class WeightDesc(C.Structure):
    _fields_ = [("weight", c_p), ("bn_weight", c_p), ("bn_var", c_p), ("bn_eps", C.c_float),
                ("cout", C.c_int32), ("cin_g", C.c_int32), ("kh", C.c_int32), ("kw", C.c_int32),
                ("layout", C.c_int32), ("observe", C.c_int32), ("averaging_const", C.c_float),
                ("wfq", FQ), ("wq", c_p), ("wq_mma", c_p), ("ldw", C.c_int32), ("wt_bf16", c_p), ("wmask", c_p), ("sf", c_p), ("rstd_run", c_p), ("wsum", c_p),
                ("dwq", c_p), ("dgamma_bn", c_p), ("dsf_bn", c_p), ("dweight", c_p), ("dgamma", c_p)]


class BnFinalizeArgs(C.Structure):
    _fields_ = [("stats", c_p), ("C", C.c_int32), ("count", C.c_int64), ("x_scale", c_p), ("w_scale", c_p),
                ("sf", c_p), ("gamma", c_p), ("beta", c_p), ("running_mean", c_p), ("running_var", c_p),
                ("num_batches_tracked", c_p), ("momentum", C.c_float), ("eps", C.c_float),
                ("stats_format", C.c_int32), ("training", C.c_int32), ("relu", C.c_int32), ("observe", C.c_int32),
                ("averaging_const", C.c_float), ("afq", FQ), ("A", c_p), ("B", c_p), ("mean_I", c_p),
                ("kfac", c_p), ("cur_minmax", c_p)]


class BnBackwardArgs(C.Structure):
    _fields_ = [("dy", c_p), ("acc", c_p), ("acc_format", C.c_int32), ("M", C.c_int64), ("C", C.c_int32), ("relu", C.c_int32),
                ("A", c_p), ("B", c_p), ("mean_I", c_p), ("kfac", c_p), ("gamma", c_p), ("sf", c_p),
                ("x_scale", c_p), ("w_scale", c_p), ("out_scale", c_p), ("out_zp", c_p), ("eps", C.c_float),
                ("sums", c_p), ("coef", c_p), ("dz", c_p), ("dz_lo", c_p), ("dz_format", C.c_int32), ("dgamma_bn", c_p), ("dbeta", c_p), ("dsf_bn", c_p),
                ("frozen", C.c_int32)]


class PwOperands(C.Structure):
    _fields_ = [("x", c_p), ("M", C.c_int64), ("K", C.c_int32), ("ldx", C.c_int32), ("w_mma", c_p), ("ldw", C.c_int32),
                ("cout", C.c_int32), ("x_zp", c_p), ("w_zp", c_p), ("wsum", c_p)]


class PwFusedFwdArgs(C.Structure):
    _fields_ = [("op", PwOperands), ("bn", BnFinalizeArgs), ("grid_barrier", c_p), ("q", c_p), ("ldq", C.c_int32)]


class PwFusedBwdArgs(C.Structure):
    _fields_ = [("op", PwOperands), ("bn", BnBackwardArgs)]


class PwChainArgs(C.Structure):
    _fields_ = [("op", PwOperands), ("bn", BnBackwardArgs), ("wt_bf16", c_p), ("dx", c_p), ("accumulate", C.c_int32), ("dwq", c_p)]


class QTensor(C.Structure):
    _fields_ = [("q", c_p), ("scale", c_p), ("zp", c_p), ("cur_minmax", c_p), ("C", C.c_int32), ("ld", C.c_int32)]


class OptTensor(C.Structure):
    _fields_ = [("p", c_p), ("g", c_p), ("exp_min", c_p), ("exp_max", c_p), ("coin_toss", c_p),
                ("buf0", c_p), ("buf1", c_p), ("buf2", c_p), ("noise", c_p), ("coin", c_p),
                ("n", C.c_int64), ("lr", C.c_float), ("weight_decay", C.c_float), ("step", C.c_int32),
                ("restart_step", C.c_int32), ("first_momentum", C.c_int32), ("pad_", C.c_int32)]


class OptHyper(C.Structure):
    _fields_ = [("kind", C.c_int32), ("is_warmup", C.c_int32), ("toss_coin", C.c_int32), ("nesterov", C.c_int32),
                ("centered", C.c_int32), ("amsgrad", C.c_int32), ("momentum", C.c_float), ("dampening", C.c_float),
                ("beta", C.c_float), ("beta1", C.c_float), ("beta2", C.c_float), ("eps", C.c_float),
                ("alpha", C.c_float), ("clip_by", C.c_float), ("noise_decay", C.c_float),
                ("grad_scale", C.c_float), ("seed", C.c_uint64)]


class OptChunk(C.Structure):
    _fields_ = [("tensor", C.c_int32), ("chunk", C.c_int32)]


OPT_CHUNK = 2048
WEIGHT_CHUNK = 8192
WEIGHT_BWD_CHANNELS = 8
FQ_SCRATCH_FLOATS = 2048
OPT_KINDS = {"QSGD": 0, "QRMS": 1, "QAdam": 2, "QAdamW": 3}

i32, i64, f32 = C.c_int, C.c_int64, C.c_float

# name -> argtypes ; every function returns int.  Mirrors include/frost_b200.h one to one.
_SIGNATURES = {
    "frost_stats_reset": [c_p, i64, c_p],
    "frost_stats_reset_f32": [c_p, i64, c_p],
    "frost_stem_conv_forward_f32": [c_p, c_p, c_p, i32, i32, i32, i32, i32, i32, i32, i32, c_p, c_p, c_p],
    "frost_stem_wgrad_f32": [c_p, c_p, i32, i32, i32, i32, i32, i32, i32, i32, c_p, c_p],
    "frost_dequant_to_nchw": [c_p, i32, c_p, c_p, i32, i32, i32, i32, c_p, c_p],
    "frost_nchw_to_nhwc": [c_p, i32, i32, i32, i32, c_p, i32, c_p],
    "frost_fq_forward": [c_p, i64, FQ, i32, i32, i32, i32, f32, c_p, c_p, c_p, c_p, c_p],
    "frost_fq_backward": [c_p, c_p, i64, c_p, c_p],
    "frost_input_quant": [c_p, i32, i32, i32, i32, FQ, i32, f32, c_p, c_p, c_p, c_p],
    "frost_input_quant_im2col": [c_p, i32, i32, i32, i32, i32, i32, i32, FQ, i32, f32, c_p, i32, c_p, c_p, c_p],
    "frost_weight_prep_multi": [c_p, i32, c_p, i32, c_p, c_p],
    "frost_weight_backward_multi": [c_p, i32, c_p, i32, c_p],
    "frost_pw_conv_forward": [c_p, c_p, c_p, c_p, c_p, i64, i32, i32, c_p, c_p, c_p],
    "frost_pw_conv_forward_simt": [c_p, c_p, c_p, c_p, c_p, i64, i32, i32, c_p, c_p, c_p],
    "frost_dw_conv_forward": [c_p, i32, c_p, c_p, c_p, i32, i32, i32, i32, i32, i32, c_p, c_p, c_p],
    "frost_stem_conv_forward": [c_p, c_p, c_p, c_p, i32, i32, i32, i32, i32, i32, i32, i32, c_p, c_p, c_p],
    "frost_bn_finalize": [C.POINTER(BnFinalizeArgs), c_p],
    "frost_bnq_apply": [c_p, i32, i64, i32, c_p, c_p, i32, c_p, c_p, c_p, i32, c_p],
    "frost_bn_backward": [C.POINTER(BnBackwardArgs), c_p],
    "frost_bn_backward_reduce": [C.POINTER(BnBackwardArgs), c_p],
    "frost_bn_backward_apply": [C.POINTER(BnBackwardArgs), c_p],
    "frost_pw_fused_forward": [C.POINTER(PwFusedFwdArgs), c_p],
    "frost_pw_fused_bwd_reduce": [C.POINTER(PwFusedBwdArgs), c_p],
    "frost_pw_fused_bwd_apply": [C.POINTER(PwFusedBwdArgs), c_p],
    "frost_pw_chain_backward": [C.POINTER(PwChainArgs), c_p],
    "frost_cat_forward": [QTensor, QTensor, i64, FQ, i32, f32, c_p, i32, c_p, c_p],
    "frost_cat_backward": [c_p, QTensor, QTensor, i64, c_p, c_p, c_p, c_p, i32, c_p],
    "frost_add_forward": [QTensor, QTensor, i64, FQ, i32, f32, c_p, i32, c_p, c_p, c_p],
    "frost_add_backward": [c_p, QTensor, QTensor, i64, c_p, c_p, c_p, c_p, i32, c_p],
    "frost_axpy": [c_p, c_p, i64, c_p],
    "frost_pool_dropout_forward": [c_p, c_p, c_p, i32, i32, i32, c_p, f32, c_p, c_p],
    "frost_pool_dropout_backward": [c_p, i32, i32, i32, c_p, f32, c_p, c_p],
    "frost_linear_forward": [c_p, c_p, c_p, c_p, c_p, i32, i32, i32, c_p, c_p],
    "frost_linear_backward": [c_p, c_p, c_p, c_p, c_p, i32, i32, i32, c_p, c_p, c_p, c_p],
    "frost_pw_dgrad": [c_p, c_p, c_p, c_p, i64, i32, i32, c_p, i32, c_p],
    "frost_pw_dgrad_tc": [c_p, c_p, c_p, c_p, i64, i32, i32, c_p, i32, c_p],
    "frost_pw_wgrad": [c_p, c_p, c_p, c_p, i64, i32, i32, c_p, c_p],
    "frost_pw_wgrad_tc": [c_p, c_p, c_p, i32, c_p, c_p, i64, i32, i32, c_p, c_p],
    "frost_dw_dgrad": [c_p, c_p, c_p, c_p, i32, i32, i32, i32, i32, i32, c_p, i32, c_p],
    "frost_dw_wgrad": [c_p, c_p, i32, c_p, c_p, i32, i32, i32, i32, i32, i32, c_p, c_p],
    "frost_dw_conv_forward_dilated": [c_p, i32, c_p, c_p, c_p, i32, i32, i32, i32, i32, i32, i32, i32, c_p, c_p, c_p],
    "frost_dw_dgrad_dilated": [c_p, c_p, c_p, c_p, i32, i32, i32, i32, i32, i32, i32, i32, c_p, i32, c_p],
    "frost_dw_wgrad_dilated": [c_p, c_p, i32, c_p, c_p, i32, i32, i32, i32, i32, i32, i32, i32, c_p, c_p],
    "frost_stem_wgrad": [c_p, c_p, c_p, c_p, i32, i32, i32, i32, i32, i32, i32, i32, c_p, c_p],
    "frost_gradboost_multi": [c_p, i32, c_p, i32, C.POINTER(OptHyper), c_p],
    "frost_set_tunable": [i32, i32],
    "frost_get_tunable": [i32],
    "frost_debug_set_trace": [c_p],
    "frost_hswish_forward": [c_p, i64, c_p, c_p, FQ, i32, FQ, i32, f32, c_p, c_p, c_p, c_p, c_p, c_p],
    "frost_hswish_backward": [c_p, c_p, i64, c_p, c_p, c_p],
    "frost_hsigmoid_forward": [c_p, i64, c_p, c_p, FQ, i32, f32, c_p, c_p, c_p, c_p, c_p, c_p],
    "frost_relu_forward": [c_p, i64, c_p, c_p, c_p],
    "frost_multibox_match": [c_p, c_p, c_p, i32, i32, c_p, i32, f32, f32, f32, c_p, c_p, c_p, c_p, c_p],
    "frost_bcast_mul_forward": [c_p, c_p, i64, i32, c_p, c_p],
    "frost_bcast_mul_backward": [c_p, c_p, c_p, i64, i32, c_p, c_p, c_p],
}
EXPORTED_SYMBOLS = sorted(list(_SIGNATURES) + ["frost_abi_version", "frost_last_error", "frost_launch_count",
                           "frost_pw_chain_supported", "frost_hswish_workspace_floats"])

_lib = None


def lib_path():
    return _build.LIB_PATH


def load():
    """Load (building first if the .so is absent and nvcc is available).  Raises if impossible."""
    global _lib
    if _lib is not None:
        return _lib
    path = _build.LIB_PATH
    if not os.path.exists(path):
        path = _build.build_library()
    lib = C.CDLL(path)
    lib.frost_abi_version.restype = C.c_int
    lib.frost_last_error.restype = C.c_char_p
    lib.frost_launch_count.restype = C.c_int64
    lib.frost_pw_chain_supported.argtypes = [C.c_int, C.c_int]
    lib.frost_pw_chain_supported.restype = C.c_int
    for name, argtypes in _SIGNATURES.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = C.c_int
    if lib.frost_abi_version() != 2:
        raise RuntimeError("libfrost_b200.so ABI mismatch; rebuild with python -m frostnet_b200.build --force")
    # measurement aid: FROST_TUNE="knob=value,knob=value" presets the launch-shape knobs (include/frost_b200.h)
    for item in filter(None, os.environ.get("FROST_TUNE", "").split(",")):
        k, v = item.split("=")
        if lib.frost_set_tunable(int(k), int(v)) != 0:
            raise RuntimeError("FROST_TUNE: " + lib.frost_last_error().decode())
    _lib = lib
    return lib


def call(name, *args):
    """Invoke a C-ABI entry point; raise RuntimeError with frost_last_error() on failure."""
    lib = load()
    rc = getattr(lib, name)(*args)
    if rc != 0:
        raise RuntimeError("%s failed (%d): %s" % (name, rc, lib.frost_last_error().decode()))


def chunk_table(counts, per_chunk, device):
    """Device int32 [n,2] work list {tensor, chunk} covering counts[i] items in slices of per_chunk."""
    rows = []
    for t, n in enumerate(counts):
        rows.extend((t, c) for c in range((int(n) + per_chunk - 1) // per_chunk))
    tab = torch.tensor(rows, dtype=torch.int32).reshape(-1, 2)
    return tab.to(device), len(rows)


def launch_count():
    return int(load().frost_launch_count())


def ptr(t):
    """Raw device pointer of a tensor (None -> NULL)."""
    if t is None:
        return None
    return t.data_ptr()


def stream(device=None):
    return torch.cuda.current_stream(device).cuda_stream


def require_cuda(t, what="tensor"):
    if not t.is_cuda:
        raise RuntimeError("frostnet_b200: %s must live on a CUDA device (B200); there is no CPU path" % what)
