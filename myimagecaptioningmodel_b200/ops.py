"""Kernel-level entry points of the C-ABI on torch tensors (used by the parity tests and bench.py's roofline leg)."""
from __future__ import annotations

import ctypes as C

import torch

from . import _capi


def _vp(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def gemm(A, B, M, N, K, *, mn_major=False, a_mn=False, b_mn=False, lda=None, ldb=None, C_out=None, ldc=None, bias=None, addend=None,
         dact_y=None, act=0, out_f32=False, atomic=False, force_simt=False, variant=None, sync=True):
    """capdec_gemm: C[M,N] = act(A.B^T + bias + addend) on CUDA tensors (A, B bf16 -> tcgen05, f32 -> SIMT).
    variant: None / 'auto' = the library's choice, 'simt', 'pair' (cta_group::2 kernel), 'single' (one-CTA kernel)."""
    lib = _capi.load_library()
    dtype = _capi.DTYPE_BF16 if A.dtype == torch.bfloat16 else _capi.DTYPE_F32
    lda = A.stride(0) if lda is None else lda
    ldb = B.stride(0) if ldb is None else ldb
    if C_out is None:
        cdt = torch.float32 if (out_f32 or atomic) else A.dtype
        C_out = torch.zeros(M, N, dtype=cdt, device=A.device)
    ldc = C_out.stride(0) if ldc is None else ldc
    majors = 3 if mn_major else ((1 if a_mn else 0) | (2 if b_mn else 0))
    var = {None: _capi.GEMM_AUTO, "auto": _capi.GEMM_AUTO, "simt": _capi.GEMM_SIMT, "pair": _capi.GEMM_PAIR,
           "single": _capi.GEMM_SINGLE}[variant]
    if force_simt:
        var = _capi.GEMM_SIMT
    _capi.check(lib.capdec_gemm(dtype, majors, _vp(A), lda, _vp(B), ldb, _vp(C_out), ldc, M, N, K,
                                _vp(bias), _vp(addend), 0 if addend is None else addend.stride(0), _vp(dact_y),
                                0 if dact_y is None else dact_y.stride(0), act, 1 if out_f32 else 0, 1 if atomic else 0,
                                var, _capi.stream_ptr()), "capdec_gemm")
    if sync:
        torch.cuda.synchronize()
    return C_out
