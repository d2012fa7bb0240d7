"""ctypes binding of include/capdec.h.  PyTorch only owns the device memory; every computation happens in
libcapdec.so.  There is no CPU fallback: a missing library, a missing GPU or a non-sm_100 GPU raises."""
from __future__ import annotations

import ctypes as C
import os

from . import _build

PARAM_NAMES = [
    "fc_v_w", "fc_v_b", "fc_k_w", "fc_k_b", "lstm_w", "lstm_b", "fc_xs_w", "fc_xs_b", "fc_hs_w", "fc_hs_b",
    "fc_p_w", "fc_p_b", "fc_g_w", "fc_g_b", "fc_s_w", "fc_s_b", "fc_a_w", "fc_a_b", "fc_o_w", "fc_o_b",
    "fc_e_w", "fc_e_b", "word_embedding", "out_fc_bias",
]

ADAPTOR_NAMES = ["fc_img_w", "fc_img_b", "fc_glob_w", "fc_glob_b"]     # model:197, 199

EXPORTS = [
    "capdec_last_error", "capdec_version", "capdec_create", "capdec_destroy", "capdec_workspace_bytes",
    "capdec_launch_count", "capdec_bind_params", "capdec_train_fwd", "capdec_train_bwd", "capdec_decode",
    "capdec_gemm", "capdec_attention_fwd", "capdec_attention_bwd", "capdec_lstm_cell_fwd", "capdec_beam_step",
    "capdec_profile", "capdec_profile_report", "capdec_set_grad_ready_events",
    "capdec_bind_adaptor", "capdec_img2feature", "capdec_img2feature_bwd", "capdec_adam_step", "capdec_decode_trace",
    "capdec_status", "capdec_cast_bf16",
]

DTYPE_F32, DTYPE_BF16 = 0, 1
SOFTMAX_FEATURE, SOFTMAX_POSITION = 0, 1
FLAG_STEPWISE_RECURRENCE = 1
GEMM_AUTO, GEMM_SIMT, GEMM_PAIR, GEMM_SINGLE = 0, 1, 2, 3


class CapdecConfig(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "hidden_dim", "embedding_size", "vocab_size", "num_positions", "sentence_length", "infer_max_length",
        "padding_idx", "start_idx", "stop_idx", "dtype", "softmax_axis", "flags")]


class CapdecParams(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in PARAM_NAMES]


class CapdecGrads(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in PARAM_NAMES]


class CapdecAdaptorParams(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ADAPTOR_NAMES]


class CapdecAdaptorGrads(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ADAPTOR_NAMES]


class CapdecError(RuntimeError):
    pass


_lib = None


def library_path() -> str:
    return _build.LIB


def load_library(build_if_missing: bool = True):
    """dlopen libcapdec.so (building it in-tree first if needed) and declare the signatures."""
    global _lib
    if _lib is not None:
        return _lib
    path = _build.LIB
    if not os.path.exists(path):
        if not build_if_missing:
            raise CapdecError("libcapdec.so is missing (%s); run `python -m myimagecaptioningmodel_b200._build`" % path)
        _build.build()
    lib = C.CDLL(path)
    vp, i32, i64, f32 = C.c_void_p, C.c_int32, C.c_int64, C.c_float
    lib.capdec_last_error.restype = C.c_char_p
    lib.capdec_last_error.argtypes = []
    lib.capdec_version.restype = C.c_int
    lib.capdec_create.argtypes = [C.POINTER(CapdecConfig), C.c_int, C.POINTER(vp)]
    lib.capdec_destroy.argtypes = [vp]
    lib.capdec_workspace_bytes.restype = C.c_size_t
    lib.capdec_workspace_bytes.argtypes = [vp]
    lib.capdec_launch_count.restype = C.c_int64
    lib.capdec_launch_count.argtypes = [vp]
    lib.capdec_set_grad_ready_events.argtypes = [vp, C.POINTER(vp), i32]
    lib.capdec_profile.argtypes = [vp, i32]
    lib.capdec_profile_report.argtypes = [vp, C.c_char_p, C.c_size_t]
    lib.capdec_bind_params.argtypes = [vp, C.POINTER(CapdecParams), vp]
    lib.capdec_train_fwd.argtypes = [vp, vp, vp, vp, vp, i32, vp, vp, vp]
    lib.capdec_train_bwd.argtypes = [vp, f32, vp, vp, vp, C.POINTER(CapdecGrads), vp]
    lib.capdec_decode.argtypes = [vp, vp, vp, i32, i32, i32, vp, vp, vp, vp]
    lib.capdec_decode_trace.argtypes = [vp, vp, vp]
    lib.capdec_status.argtypes = [vp, vp]
    lib.capdec_cast_bf16.argtypes = [i32, vp, vp, i64, vp]
    lib.capdec_bind_adaptor.argtypes = [vp, i32, C.POINTER(CapdecAdaptorParams), vp]
    lib.capdec_img2feature.argtypes = [vp, vp, i32, vp, vp, vp]
    lib.capdec_img2feature_bwd.argtypes = [vp, vp, C.POINTER(CapdecAdaptorGrads), vp]
    lib.capdec_adam_step.argtypes = [vp, vp, vp, vp, i64] + [f32] * 7 + [vp]
    lib.capdec_gemm.argtypes = [i32, i32, vp, i64, vp, i64, vp, i64, i32, i32, i32, vp, vp, i64, vp, i64, i32, i32,
                                i32, i32, vp]
    lib.capdec_attention_fwd.argtypes = [i32, i32, i32, i32, i32, i64, i64, i32] + [vp] * 12
    lib.capdec_attention_bwd.argtypes = [i32, i32, i32, i32, i32, i64, i64, i32] + [vp] * 16
    lib.capdec_lstm_cell_fwd.argtypes = [i32, i32, i32, vp, vp, vp, vp, vp]
    lib.capdec_beam_step.argtypes = [vp, i64, i32, i32, i32, i32, i32, i32, i32, i32] + [vp] * 9
    for name in EXPORTS:
        fn = getattr(lib, name)
        if name not in ("capdec_last_error", "capdec_workspace_bytes", "capdec_launch_count", "capdec_version"):
            fn.restype = C.c_int
    _lib = lib
    return lib


def check(rc: int, what: str = "capdec"):
    if rc != 0:
        msg = load_library().capdec_last_error()
        raise CapdecError("%s failed: %s" % (what, msg.decode("utf-8", "replace") if msg else "unknown error"))


# Version counters of flat parameter buffers, keyed by device pointer.  Whoever writes parameters in place (optim.Adam,
# ParameterStore.load / init_reference, checkpoint.load_params) bumps the buffer's version; the engine re-runs
# capdec_bind_params / capdec_bind_adaptor whenever the version it bound last is stale (include/capdec.h: "call again
# after every optimizer update" -- enforced here instead of being left to the caller).
_param_versions = {}


def bump_version(flat) -> None:
    k = int(flat.data_ptr())
    _param_versions[k] = _param_versions.get(k, 0) + 1


def version_of(flat) -> int:
    return _param_versions.get(int(flat.data_ptr()), 0)


def ptr(t):
    """device pointer of a torch tensor (or None -> NULL)"""
    return None if t is None else C.c_void_p(t.data_ptr())


def stream_ptr(stream=None):
    import torch
    s = torch.cuda.current_stream() if stream is None else stream
    return C.c_void_p(s.cuda_stream)
