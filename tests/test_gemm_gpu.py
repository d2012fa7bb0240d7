"""-m gpu: the tcgen05/TMEM/TMA GEMM (and the SIMT validation GEMM) against a plain PyTorch fp32 reference of the
same contraction on the same bf16-rounded inputs.  Tolerance: fp32 accumulation of exact bf16 products, so the only
difference is summation order -> rel 1e-5 for fp32 outputs; bf16 outputs add one rounding (2^-9 relative)."""
import pytest
import torch

from gpu_util import gemm, max_err, rel_err

pytestmark = pytest.mark.gpu

F32_TOL = 2e-5
BF16_TOL = 6e-3


def _rand(shape, dtype, seed):
    gen = torch.Generator(device="cuda").manual_seed(seed)
    return (torch.randn(shape, generator=gen, device="cuda", dtype=torch.float32) * 0.5).to(dtype)


KK_SHAPES = [
    (128, 64, 64), (128, 128, 128), (128, 256, 64), (256, 4096, 1024), (1000, 1024, 1024), (200, 72, 96),
    (130, 333, 256), (8, 40, 32), (256, 1024, 4096), (8704, 256, 1024), (1300, 12295, 256),
]


@pytest.mark.parametrize("M,N,K", KK_SHAPES)
@pytest.mark.parametrize("simt", [False, True])
def test_gemm_k_major(M, N, K, simt):
    if simt and M * N * K > 2e10:
        pytest.skip("SIMT path is a validation path; skip the largest shapes")
    A = _rand((M, K), torch.bfloat16, 1)
    B = _rand((N, K), torch.bfloat16, 2)
    ref = A.float() @ B.float().t()
    ldc = (N + 7) // 8 * 8
    Cbuf = torch.zeros(M, ldc, dtype=torch.float32, device="cuda")
    out = gemm(A, B, M, N, K, C_out=Cbuf, ldc=ldc, out_f32=True, force_simt=simt)[:, :N]
    assert rel_err(out, ref) < F32_TOL, (rel_err(out, ref), max_err(out, ref))
    if ldc > N:
        assert float(Cbuf[:, N:].abs().max()) == 0.0          # columns past N are never written


@pytest.mark.parametrize("act", [0, 1, 2])
def test_gemm_epilogues(act):
    M, N, K = 300, 520, 192
    A, B = _rand((M, K), torch.bfloat16, 3), _rand((N, K), torch.bfloat16, 4)
    bias = _rand((N,), torch.float32, 5)
    addend = _rand((M, N), torch.bfloat16, 6)
    y = torch.tanh(_rand((M, N), torch.float32, 7)).to(torch.bfloat16)
    ref = A.float() @ B.float().t() + bias + addend.float()
    if act == 1:
        ref = torch.tanh(ref)
    if act == 2:
        ref = ref * (1 - y.float() ** 2)
    for simt in (False, True):
        out = gemm(A, B, M, N, K, bias=bias, addend=addend, dact_y=y if act == 2 else None, act=act, force_simt=simt)
        assert out.dtype == torch.bfloat16
        assert rel_err(out.float(), ref) < BF16_TOL, (simt, rel_err(out.float(), ref))
        out32 = gemm(A, B, M, N, K, bias=bias, addend=addend, dact_y=y if act == 2 else None, act=act, out_f32=True,
                     force_simt=simt)
        tol = 2e-3 if act == 1 else F32_TOL          # tanh.approx on the tensor-core path
        assert rel_err(out32, ref) < tol, (simt, act, rel_err(out32, ref))


def test_gemm_strided_operands():
    """sub-block views: lda / ldb / ldc larger than the logical widths (how the schedule slices dG, dKV, ...)"""
    M, N, K = 260, 200, 160
    Abig, Bbig = _rand((M, 5 * K), torch.bfloat16, 8), _rand((N, 3 * K), torch.bfloat16, 9)
    A, B = Abig[:, 2 * K:3 * K], Bbig[:, K:2 * K]
    ref = A.float() @ B.float().t()
    Cbig = torch.full((M, 2 * N + 8), 7.0, dtype=torch.float32, device="cuda")
    gemm(A, B, M, N, K, lda=5 * K, ldb=3 * K, C_out=Cbig[:, 8:], ldc=2 * N + 8, out_f32=True)
    assert rel_err(Cbig[:, 8:8 + N], ref) < F32_TOL
    assert float((Cbig[:, :8] - 7).abs().max()) == 0 and float((Cbig[:, 8 + N:] - 7).abs().max()) == 0


MM_SHAPES = [(128, 64, 64), (256, 128, 512), (1024, 1024, 8704), (100, 72, 200), (256, 4096, 300), (12295, 256, 1088),
             (64, 1024, 256)]


@pytest.mark.parametrize("M,N,K", MM_SHAPES)
@pytest.mark.parametrize("simt", [False, True])
def test_gemm_mn_major_weight_grad_form(M, N, K, simt):
    """dW[in,out] += X[rows,in]^T . dY[rows,out]  (contraction over the row index; fp32 atomics, split-K)"""
    Mp, Np = (M + 7) // 8 * 8, (N + 7) // 8 * 8
    X = _rand((K, Mp), torch.bfloat16, 10)
    dY = _rand((K, Np), torch.bfloat16, 11)
    ref = X[:, :M].float().t() @ dY[:, :N].float()
    out = torch.zeros(M, N, dtype=torch.float32, device="cuda")
    gemm(X, dY, M, N, K, mn_major=True, lda=Mp, ldb=Np, C_out=out, ldc=N, atomic=True, force_simt=simt)
    assert rel_err(out, ref) < F32_TOL, (rel_err(out, ref), max_err(out, ref))
    gemm(X, dY, M, N, K, mn_major=True, lda=Mp, ldb=Np, C_out=out, ldc=N, atomic=True, force_simt=simt)   # accumulates
    assert rel_err(out, 2 * ref) < F32_TOL


@pytest.mark.parametrize("M,N,K", [(300, 200, 160), (8704, 1024, 1024), (256, 4096, 1024), (520, 12295, 256), (16, 40, 72)])
@pytest.mark.parametrize("simt", [False, True])
def test_gemm_mixed_majors(M, N, K, simt):
    """forward fc form: A = X [M,K] K-major, B = W [K,N] (PaddlePaddle [in,out]) MN-major; and the mirrored form"""
    Np, Mp = (N + 7) // 8 * 8, (M + 7) // 8 * 8
    X = _rand((M, K), torch.bfloat16, 20)
    W = _rand((K, Np), torch.bfloat16, 21)
    ref = X.float() @ W[:, :N].float()
    out = torch.zeros(M, Np, dtype=torch.float32, device="cuda")
    gemm(X, W, M, N, K, b_mn=True, ldb=Np, C_out=out, ldc=Np, out_f32=True, force_simt=simt)
    assert rel_err(out[:, :N], ref) < F32_TOL, rel_err(out[:, :N], ref)
    Xt = _rand((K, Mp), torch.bfloat16, 22)
    Wt = _rand((N, K), torch.bfloat16, 23)
    ref2 = Xt[:, :M].float().t() @ Wt.float().t()
    out2 = torch.zeros(M, Np, dtype=torch.float32, device="cuda")
    gemm(Xt, Wt, M, N, K, a_mn=True, lda=Mp, C_out=out2, ldc=Np, out_f32=True, force_simt=simt)
    assert rel_err(out2[:, :N], ref2) < F32_TOL, rel_err(out2[:, :N], ref2)


def test_gemm_many_tiles_per_cta():
    """persistent scheduling: more work items than SMs, odd tile counts, both accumulator buffers reused many times"""
    M, N, K = 128 * 37 + 5, 64 * 11 + 3, 200
    A, B = _rand((M, K), torch.bfloat16, 30), _rand((N, K), torch.bfloat16, 31)
    ref = A.float() @ B.float().t()
    ldc = (N + 7) // 8 * 8
    out = torch.zeros(M, ldc, dtype=torch.float32, device="cuda")
    gemm(A, B, M, N, K, C_out=out, ldc=ldc, out_f32=True)
    assert rel_err(out[:, :N], ref) < F32_TOL


def test_gemm_fp32_validation_path():
    M, N, K = 150, 130, 77
    A, B = _rand((M, K), torch.float32, 12), _rand((N, K), torch.float32, 13)
    out = gemm(A, B, M, N, K)
    assert out.dtype == torch.float32
    assert rel_err(out, A.double() @ B.double().t()) < 1e-6
    X, dY = _rand((K, M), torch.float32, 14), _rand((K, N), torch.float32, 15)
    out = gemm(X, dY, M, N, K, mn_major=True, atomic=True)
    assert rel_err(out, X.double().t() @ dY.double()) < 1e-6


def test_gemm_linearity_at_full_size():
    """size-independent property at the BASELINE shape of the head GEMMs (M = 256*34): C(A1 + A2) = C(A1) + C(A2)
    when A1, A2, A1 + A2 are exactly representable in bf16."""
    M, N, K = 8704, 1024, 1024
    gen = torch.Generator(device="cuda").manual_seed(3)
    A1 = torch.randint(-4, 5, (M, K), generator=gen, device="cuda").to(torch.bfloat16)
    A2 = torch.randint(-4, 5, (M, K), generator=gen, device="cuda").to(torch.bfloat16)
    B = torch.randint(-2, 3, (N, K), generator=gen, device="cuda").to(torch.bfloat16)
    c1 = gemm(A1, B, M, N, K, out_f32=True)
    c2 = gemm(A2, B, M, N, K, out_f32=True)
    c12 = gemm((A1.float() + A2.float()).to(torch.bfloat16), B, M, N, K, out_f32=True)
    assert torch.equal(c1 + c2, c12)                       # small integers: exact in fp32
    assert torch.equal(c1, A1.float() @ B.float().t())


@pytest.mark.parametrize("mt", ["single", "pair"])
@pytest.mark.parametrize("majors", [0, 2, 3])
@pytest.mark.parametrize("M,N,K", [(304, 520, 192), (264, 136, 64), (8704, 1024, 1024), (1024, 1024, 8704), (136, 72, 40),
                                   (640, 12295, 256)])
def test_gemm_both_tensor_core_kernels(M, N, K, majors, mt):
    """The single-CTA kernel (cta_group::1, 128-row tiles) and the CTA-pair kernel (tcgen05 cta_group::2, 256 x 256 tiles,
    wherever eligible) forced through capdec_gemm's `variant`, on ragged and full-size shapes, every operand-major form,
    bf16 output with the fused epilogue and fp32 split-K reduction."""
    a_mn, b_mn = bool(majors & 1), bool(majors & 2)
    if b_mn and N % 8:
        pytest.skip("an MN-major operand needs a 16-byte row pitch")
    A = _rand((K, M) if a_mn else (M, K), torch.bfloat16, 11)
    B = _rand((K, N) if b_mn else (N, K), torch.bfloat16, 12)
    Af = A.float().t() if a_mn else A.float()
    Bf = B.float() if b_mn else B.float().t()
    ref = Af @ Bf
    ldc = (N + 7) // 8 * 8
    if majors == 3:   # weight-gradient form: fp32 accumulation into C (split-K, TMA reduce-add)
        Cbuf = torch.ones(M, ldc, dtype=torch.float32, device="cuda")
        out = gemm(A, B, M, N, K, a_mn=True, b_mn=True, C_out=Cbuf, ldc=ldc, atomic=True, variant=mt)[:, :N]
        assert rel_err(out - 1.0, ref) < 5e-5, rel_err(out - 1.0, ref)
    else:
        bias = _rand((ldc,), torch.float32, 13)
        addend = _rand((M, ldc), torch.bfloat16, 14)
        Cbuf = torch.zeros(M, ldc, dtype=torch.bfloat16, device="cuda")
        out = gemm(A, B, M, N, K, a_mn=a_mn, b_mn=b_mn, C_out=Cbuf, ldc=ldc, bias=bias, addend=addend,
                   act=1, variant=mt)[:, :N]
        ref2 = torch.tanh(ref + bias[:N] + addend[:, :N].float())
        assert rel_err(out.float(), ref2) < BF16_TOL, rel_err(out.float(), ref2)


@pytest.mark.parametrize("M,N,K,b_mn", [(4096, 4000, 256, False), (4096, 4104, 448, True), (8704, 12295, 256, False)])
def test_gemm_short_k_many_tiles(M, N, K, b_mn):
    """short-K problems with several output tiles per CTA pair take the pair kernel's variant with two staging tiles per
    epilogue warp (4 pipeline stages): fp32 and bf16 outputs, ragged right edge, relu / drelu epilogues"""
    A = _rand((M, K), torch.bfloat16, 11)
    B = _rand((K, N + (-N) % 8) if b_mn else (N, K), torch.bfloat16, 12)
    Bm = B[:, :N].float() if b_mn else B.float().t()
    ref = A.float() @ Bm
    ldc = (N + 7) // 8 * 8
    Cbuf = torch.zeros(M, ldc, dtype=torch.float32, device="cuda")
    out = gemm(A, B, M, N, K, b_mn=b_mn, C_out=Cbuf, ldc=ldc, out_f32=True)[:, :N]
    assert rel_err(out, ref) < F32_TOL, rel_err(out, ref)
    if ldc > N:
        assert float(Cbuf[:, N:].abs().max()) == 0.0
    if M * N <= 2e7:
        Cb = torch.zeros(M, ldc, dtype=torch.bfloat16, device="cuda")
        outb = gemm(A, B, M, N, K, b_mn=b_mn, C_out=Cb, ldc=ldc, act=5)[:, :N]          # ACT_RELU
        assert rel_err(outb.float(), torch.relu(ref)) < BF16_TOL
        y = torch.relu(_rand((M, ldc), torch.float32, 13)).to(torch.bfloat16)
        outd = gemm(A, B, M, N, K, b_mn=b_mn, C_out=Cb, ldc=ldc, dact_y=y, act=6)[:, :N]   # ACT_DRELU
        assert rel_err(outd.float(), ref * (y[:, :N].float() > 0)) < BF16_TOL
