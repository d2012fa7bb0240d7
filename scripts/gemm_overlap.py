"""Do two independent head GEMMs overlap when they are issued on two streams?  
8 products C_i = A_i . W (M rows, N = K = 1024, bf16) timed with CUDA events: all on one stream, alternating on two
streams (fork / join by events), and as the single-CTA variant; M = live rows of the bench batch and neighbours."""
import sys

import torch

sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
from myimagecaptioningmodel_b200.ops import gemm  # noqa: E402

dev = torch.device("cuda:0")
N = K = 1024
W = (torch.randn(K, N, device=dev) * 0.03).bfloat16()
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run(M, mode, variant, n=8, reps=20):
    A = [(torch.randn(M, K, device=dev) * 0.5).bfloat16() for _ in range(n)]
    Cs = [torch.empty(M, N, dtype=torch.bfloat16, device=dev) for _ in range(n)]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def body():
        if mode == "one":
            with torch.cuda.stream(s1):
                for i in range(n):
                    gemm(A[i], W, M, N, K, b_mn=True, C_out=Cs[i], variant=variant, sync=False)
        else:
            s2.wait_stream(s1)
            for i in range(n):
                with torch.cuda.stream(s1 if i % 2 == 0 else s2):
                    gemm(A[i], W, M, N, K, b_mn=True, C_out=Cs[i], variant=variant, sync=False)
            s1.wait_stream(s2)

    for _ in range(3):
        body()
    torch.cuda.synchronize()
    with torch.cuda.stream(s1):
        e0.record()
    for _ in range(reps):
        body()
    with torch.cuda.stream(s1):
        e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / reps / n


print("us per GEMM (N = K = 1024), 8 independent products")
print("%6s %-7s %10s %10s" % ("M", "kernel", "one stream", "two streams"))
for M in (2368, 4736, 4944, 8704, 9472):
    for variant in ("pair", "single"):
        a = run(M, "one", variant)
        b = run(M, "two", variant)
        print("%6d %-7s %10.2f %10.2f" % (M, variant, a, b), flush=True)
