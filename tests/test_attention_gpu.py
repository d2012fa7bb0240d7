"""-m gpu: fused additive-attention kernels against a PyTorch restatement of model:103-113 (autograd for the
backward).  fp32 storage: rel 1e-4 (only tanhf / expf differences); bf16 storage: rel 2e-2."""
import ctypes as C

import pytest
import torch

from gpu_util import rel_err, vp
from myimagecaptioningmodel_b200 import _capi

pytestmark = pytest.mark.gpu


def _ref(keys, vals, he, se, s, w, b, axis):
    """keys/vals [n,k,H]; he/se/s [n,q,H] -> alpha [n,q,k+1], ctx [n,q,H]  (all float64)"""
    feat = torch.cat([keys.unsqueeze(1).expand(-1, he.shape[1], -1, -1), se.unsqueeze(2)], dim=2)   # [n,q,k+1,H]
    z = torch.tanh(feat + he.unsqueeze(2))
    e = z @ w + b
    alpha = torch.softmax(e, dim=-1) if axis == "position" else torch.ones_like(e)
    v = torch.cat([vals.unsqueeze(1).expand(-1, he.shape[1], -1, -1), s.unsqueeze(2)], dim=2)
    ctx = (v * alpha.unsqueeze(-1)).mean(dim=2)
    return alpha, ctx


def _make(n, q, k, H, seed):
    gen = torch.Generator().manual_seed(seed)
    r = lambda *sh: torch.randn(*sh, generator=gen, dtype=torch.float64)
    return dict(keys=r(n, k, H), vals=torch.tanh(r(n, k, H)), he=r(n, q, H), se=r(n, q, H), s=0.5 * r(n, q, H),
                ph=torch.tanh(r(n, q, H)), w=r(H) / H ** 0.5, b=r(1), dctx=r(n, q, H))


def _rows(x, layout):
    """[n,q,H] -> [rows,H] in the kernel's row order: 'train' = time-major (q outer), 'decode' = image-major"""
    return (x.transpose(0, 1) if layout == "train" else x).reshape(-1, x.shape[-1]).contiguous()


def _unrows(x, n, q, layout):
    return x.reshape(q, n, -1).transpose(0, 1) if layout == "train" else x.reshape(n, q, -1)


CASES = [  # n_img, q, k, H, layout
    (5, 34, 49, 1024, "train"), (7, 3, 49, 1024, "decode"), (3, 6, 5, 64, "train"), (4, 1, 49, 512, "decode"),
    (2, 4, 16, 256, "train"), (3, 2, 7, 40, "decode"),
    # training-shaped tensor-core path (q >= 8, H in {256, 512, 1024}): ragged position blocks, k up to 63
    (6, 34, 49, 1024, "train"), (3, 12, 20, 256, "train"), (2, 9, 63, 512, "decode"), (4, 17, 8, 256, "train"),
    (3, 8, 47, 1024, "train"),
]


@pytest.mark.parametrize("n,q,k,H,layout", CASES)
@pytest.mark.parametrize("axis", ["position", "feature"])
@pytest.mark.parametrize("dtype", ["fp32", "bf16"])
def test_attention_forward_backward(n, q, k, H, layout, axis, dtype):
    lib = _capi.load_library()
    d = _make(n, q, k, H, 1000 * n + q)
    tdt = torch.float32 if dtype == "fp32" else torch.bfloat16
    cdt = _capi.DTYPE_F32 if dtype == "fp32" else _capi.DTYPE_BF16
    dev = lambda x: x.to(tdt).cuda().contiguous()
    keys, vals = dev(d["keys"]), dev(d["vals"])
    he, se, s, ph, dctx = (dev(_rows(d[n_], layout)) for n_ in ("he", "se", "s", "ph", "dctx"))
    w, b = d["w"].float().cuda(), d["b"].float().cuda()
    # reference on the storage-rounded inputs
    R = {k_: v.to(tdt).double() for k_, v in d.items()}
    for k_ in ("he", "se", "s", "keys", "vals"):
        R[k_].requires_grad_(True)
    wr, br = d["w"].float().double().requires_grad_(True), d["b"].float().double().requires_grad_(True)
    alpha_ref, ctx_ref = _ref(R["keys"], R["vals"], R["he"], R["se"], R["s"], wr, br, axis)
    (ctx_ref * R["dctx"]).sum().backward()

    rows = n * q
    P = k + 1
    r0, qs = (1, n) if layout == "train" else (q, 1)
    ax = _capi.SOFTMAX_POSITION if axis == "position" else _capi.SOFTMAX_FEATURE
    alpha = torch.zeros(rows, P, dtype=torch.float32, device="cuda")
    actx = torch.zeros(rows, H, dtype=tdt, device="cuda")
    ctx = torch.zeros(rows, H, dtype=tdt, device="cuda")
    _capi.check(lib.capdec_attention_fwd(cdt, H, k, n, q, r0, qs, ax, vp(keys), vp(vals), vp(he), vp(se), vp(s), vp(ph),
                                         vp(w), vp(b), vp(alpha), vp(actx), vp(ctx), _capi.stream_ptr()))
    torch.cuda.synchronize()
    tol = 1e-4 if dtype == "fp32" else 2e-2
    assert rel_err(_unrows(alpha, n, q, layout), alpha_ref.detach()) < tol
    assert rel_err(_unrows(ctx.float(), n, q, layout), ctx_ref.detach()) < tol
    assert rel_err(_unrows(actx.float(), n, q, layout), (ctx_ref.detach() + R["ph"])) < tol
    if axis == "feature":
        assert torch.equal(alpha, torch.ones_like(alpha))            # alpha == 1 exactly

    de = torch.zeros(rows, P, dtype=torch.float32, device="cuda")
    dS = torch.zeros(rows, H, dtype=tdt, device="cuda")
    dKV = torch.zeros(n * k, 2 * H, dtype=tdt, device="cuda")
    dhe = torch.full((rows, H), 9.0, dtype=tdt, device="cuda")
    dse = torch.full((rows, H), 9.0, dtype=tdt, device="cuda")
    dw = torch.zeros(H, dtype=torch.float32, device="cuda")
    db = torch.zeros(1, dtype=torch.float32, device="cuda")
    _capi.check(lib.capdec_attention_bwd(cdt, H, k, n, q, r0, qs, ax, vp(keys), vp(vals), vp(he), vp(se), vp(s), vp(w),
                                         vp(alpha), vp(dctx), vp(de), vp(dS), vp(dKV), vp(dhe), vp(dse), vp(dw), vp(db),
                                         _capi.stream_ptr()))
    torch.cuda.synchronize()
    dV_pre_ref = (R["vals"].grad * (1 - R["vals"].detach() ** 2)).reshape(n * k, H)
    assert rel_err(dKV[:, :H].float(), dV_pre_ref) < tol
    assert rel_err(_unrows(dS.float(), n, q, layout), R["s"].grad) < tol
    if axis == "feature":
        for t in (dKV[:, H:], dhe, dse, dw, db, de):
            assert float(t.float().abs().max()) == 0.0                 # dead gradients are exactly zero
    else:
        assert rel_err(dKV[:, H:].float(), R["keys"].grad.reshape(n * k, H)) < tol
        assert rel_err(_unrows(dhe.float(), n, q, layout), R["he"].grad) < tol
        assert rel_err(_unrows(dse.float(), n, q, layout), R["se"].grad) < tol
        assert rel_err(dw, wr.grad) < tol
        assert abs(float(db) - float(br.grad)) < 1e-3 * max(1.0, float(wr.grad.abs().max()))
