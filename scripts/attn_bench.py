"""Attention kernel micro-benchmark through the C-ABI (CUDA-event timing): train layout (B=256, T=34) and decode
layout (512 images x beam 3), bf16, position soft-max."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from myimagecaptioningmodel_b200 import _capi  # noqa: E402

lib = _capi.load_library()
H, k = 1024, 49
P = k + 1
vp = lambda t: None if t is None else C.c_void_p(t.data_ptr())


def bench(name, n_img, q, r0, qs, bwd=True, reps=10):
    rows = n_img * q
    g = torch.Generator(device="cuda").manual_seed(1)
    rnd = lambda *sh: torch.randn(*sh, generator=g, device="cuda").to(torch.bfloat16)
    keys, vals = rnd(n_img, k, H), torch.tanh(rnd(n_img, k, H).float()).to(torch.bfloat16)
    he, se, sn, ph, dctx = rnd(rows, H), rnd(rows, H), rnd(rows, H), rnd(rows, H), rnd(rows, H)
    w = torch.randn(H, generator=g, device="cuda") / 32
    b = torch.zeros(1, device="cuda")
    alpha = torch.zeros(rows, P, device="cuda")
    actx = torch.zeros(rows, H, device="cuda", dtype=torch.bfloat16)
    de = torch.zeros(rows, P, device="cuda")
    dS = torch.zeros_like(actx); dhe = torch.zeros_like(actx); dse = torch.zeros_like(actx)
    dKV = torch.zeros(n_img * k, 2 * H, device="cuda", dtype=torch.bfloat16)
    dw = torch.zeros(H, device="cuda"); db = torch.zeros(1, device="cuda")

    def fwd():
        _capi.check(lib.capdec_attention_fwd(1, H, k, n_img, q, r0, qs, 1, vp(keys), vp(vals), vp(he), vp(se), vp(sn), vp(ph),
                                             vp(w), vp(b), vp(alpha), vp(actx), None, _capi.stream_ptr()))

    def bwdf():
        _capi.check(lib.capdec_attention_bwd(1, H, k, n_img, q, r0, qs, 1, vp(keys), vp(vals), vp(he), vp(se), vp(sn), vp(w),
                                             vp(alpha), vp(dctx), vp(de), vp(dS), vp(dKV), vp(dhe), vp(dse), vp(dw), vp(db),
                                             _capi.stream_ptr()))
    for fn, tag in ((fwd, "fwd"),) + (((bwdf, "bwd"),) if bwd else ()):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / reps * 1e3
        algo = rows * 211144 if tag == "fwd" else rows * 211144 * 1.0
        print("%-7s %s  rows=%6d  %8.1f us   %.2f us/row-wave   %6.0f GB/s algorithmic (per-step figure)" % (
            name, tag, rows, us, us / rows * 148, algo / us / 1e3), flush=True)


bench("train", 256, 34, 1, 256)
bench("decode", 512, 3, 3, 1, bwd=False, reps=20)
bench("decode1", 512, 1, 1, 1, bwd=False, reps=20)
