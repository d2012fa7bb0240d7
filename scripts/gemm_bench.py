"""GEMM micro-benchmark through capdec_gemm (CUDA-event timing), shapes of the decoder train step."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from myimagecaptioningmodel_b200 import _capi  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--shapes", default="head,vocab,wgrad,rec,kv,dh")
ap.add_argument("--reps", type=int, default=20)
ap.add_argument("--profile", action="store_true", help="wrap the timed launches in cudaProfilerStart/Stop")
args = ap.parse_args()
lib = _capi.load_library()
SHAPES = {  # name: (M, N, K, majors, out_f32, atomic, act)
    "head": (8704, 1024, 1024, 2, 0, 0, 1),
    "headk": (8704, 1024, 1024, 0, 0, 0, 0),
    "vocab": (8704, 12295, 256, 0, 1, 0, 0),
    "wgrad": (1024, 1024, 8704, 3, 1, 1, 0),
    "rec": (256, 4096, 1024, 2, 1, 0, 0),
    "kv": (12544, 1024, 1024, 2, 0, 0, 1),
    "dh": (256, 1024, 4096, 0, 1, 1, 0),
    "xe": (8704, 5120, 256, 2, 0, 0, 0),
    "big": (8192, 8192, 8192, 0, 0, 0, 0),
}
for name in args.shapes.split(","):
    M, N, K, majors, out_f32, atomic, act = SHAPES[name]
    a_mn, b_mn = majors & 1, majors & 2
    Np = (N + 127) // 128 * 128
    A = (torch.randn((K, M) if a_mn else (M, K), device="cuda") * 0.5).to(torch.bfloat16)
    B = (torch.randn((K, Np) if b_mn else (N, K), device="cuda") * 0.5).to(torch.bfloat16)
    C = torch.zeros(M, Np, device="cuda", dtype=torch.float32 if (out_f32 or atomic) else torch.bfloat16)
    # capdec_gemm has no way to say "the bias is readable past N" (GemmEpi::aux_padded, set by the decoder schedule), so for
    # N % 8 != 0 a bias would push the launch onto the scalar-store fallback (3.6 ms instead of ~140 us for the vocabulary
    # shape) -- time those shapes without a bias
    bias = torch.zeros(Np, device="cuda") if N % 8 == 0 else None

    def run():
        _capi.check(lib.capdec_gemm(_capi.DTYPE_BF16, majors, A.data_ptr(), A.stride(0), B.data_ptr(), B.stride(0),
                                    C.data_ptr(), Np, M, N, K, bias.data_ptr() if bias is not None else None, None, 0, None, 0, act, out_f32, atomic, 0,
                                    _capi.stream_ptr()), "gemm")
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    if args.profile:
        torch.cuda.profiler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.reps):
        run()
    e1.record()
    torch.cuda.synchronize()
    if args.profile:
        torch.cuda.profiler.stop()
    ms = e0.elapsed_time(e1) / args.reps
    print("%-6s M=%5d N=%5d K=%5d majors=%d  %8.1f us  %7.1f TFLOP/s  BN=%s" % (
        name, M, N, K, majors, ms * 1e3, 2.0 * M * N * K / ms / 1e9, os.environ.get("CAPDEC_GEMM_BN", "auto")), flush=True)
