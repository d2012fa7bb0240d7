"""helpers shared by the -m gpu parity tests"""
import ctypes as C

import torch

from myimagecaptioningmodel_b200 import _capi


def rel_err(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


def max_err(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).abs().max() / (b.abs().max() + 1e-30))


def vp(t):
    return None if t is None else C.c_void_p(t.data_ptr())


from myimagecaptioningmodel_b200.ops import gemm  # noqa: E402,F401
