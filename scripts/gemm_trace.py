"""Per-CTA timeline of one tcgen05 GEMM launch (debug hook capdec_debug_gemm_trace)."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from myimagecaptioningmodel_b200 import _capi  # noqa: E402

lib = _capi.load_library()
lib.capdec_debug_gemm_trace.argtypes = [C.c_void_p]
lib.capdec_debug_gemm_trace.restype = None
shape = sys.argv[1] if len(sys.argv) > 1 else "head"
SH = {"head": (8704, 1024, 1024, 2, 0, 0, 1), "vocab": (8704, 12295, 256, 0, 1, 0, 0), "rec": (256, 4096, 1024, 2, 1, 0, 0),
      "wgrad": (1024, 1024, 8704, 3, 1, 1, 0), "big": (8192, 8192, 8192, 0, 0, 0, 0)}
M, N, K, majors, out_f32, atomic, act = SH[shape]
a_mn, b_mn = majors & 1, majors & 2
Np = (N + 127) // 128 * 128
A = (torch.randn((K, M) if a_mn else (M, K), device="cuda") * 0.5).to(torch.bfloat16)
B = (torch.randn((K, Np) if b_mn else (N, K), device="cuda") * 0.5).to(torch.bfloat16)
Cc = torch.zeros(M, Np, device="cuda", dtype=torch.float32 if (out_f32 or atomic) else torch.bfloat16)
bias = torch.zeros(Np, device="cuda")
trace = torch.zeros(148 * 32, dtype=torch.int64, device="cuda")


def run():
    _capi.check(lib.capdec_gemm(_capi.DTYPE_BF16, majors, A.data_ptr(), A.stride(0), B.data_ptr(), B.stride(0), Cc.data_ptr(),
                                Np, M, N, K, bias.data_ptr(), None, 0, None, 0, act, out_f32, atomic, 0, _capi.stream_ptr()))


for _ in range(3):
    run()
torch.cuda.synchronize()
lib.capdec_debug_gemm_trace(C.c_void_p(trace.data_ptr()))
run()
torch.cuda.synchronize()
lib.capdec_debug_gemm_trace(None)
t = trace.cpu().view(148, 32)
t0 = int(t[:, 0][t[:, 0] > 0].min())
names = {0: "entry", 1: "setup done", 2: "1st TMA issued", 3: "producer done", 4: "1st data landed", 8: "mma0 issued", 9: "mma1 issued",
         10: "mma2 issued", 11: "mma3 issued", 12: "acc0 full", 13: "acc1 full", 14: "acc2 full", 15: "acc3 full",
         16: "epi0 done", 17: "epi1 done", 18: "epi2 done", 19: "epi3 done", 20: "exit",
         21: "p1 start", 22: "p1 tmem loaded", 23: "p1 buffer free", 24: "p1 math+sts done", 25: "p1 fenced", 26: "p1 store issued"}
print("==== %s M=%d N=%d K=%d" % (shape, M, N, K))
for cta in (0, 1, 73, 147):
    print("CTA %d:" % cta, "  ".join("%s=%.2f" % (names[s], (int(t[cta, s]) - t0) / 1e3) for s in sorted(names) if t[cta, s] > 0))
allv = t[:, 20][t[:, 20] > 0]
print("kernel span (first entry -> last exit): %.2f us over %d CTAs" % ((int(allv.max()) - t0) / 1e3, len(allv)))
for s in sorted(names):
    col = t[:, s][t[:, s] > 0]
    if len(col):
        print("  %-16s min %7.2f  median %7.2f  max %7.2f us" % (names[s], (int(col.min()) - t0) / 1e3,
                                                                (int(col.median()) - t0) / 1e3, (int(col.max()) - t0) / 1e3))
