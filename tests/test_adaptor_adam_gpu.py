"""-m gpu: the callers either side of the decoder through the C-ABI against oracle/train_ref.py
  * capdec_img2feature / capdec_img2feature_bwd (model:193-199) alone and in front of the train / decode schedules,
  * capdec_adam_step (PaddlePaddle 1.8 rule, train.py:26-31) on flat buffers, and a few complete optimisation steps.
Tolerances as everywhere: fp32 storage rel 1e-3, bf16 storage rel 2e-2 (Frobenius-relative per tensor); Adam is fp32
arithmetic in both implementations: rel 1e-5 on the update."""
import numpy as np
import pytest
import torch

import myimagecaptioningmodel_b200 as M
from gpu_util import rel_err
from oracle import decoder_ref as R
from oracle import synth
from oracle import train_ref as TR

pytestmark = pytest.mark.gpu

SMALL = dict(hidden_dim=64, embedding_size=32, vocab_size=101, num_positions=5, sentence_length=8, infer_max_length=9)
MID = dict(hidden_dim=256, embedding_size=64, vocab_size=1000, num_positions=49, sentence_length=12, infer_max_length=12)
TOL = {"fp32": 1e-3, "bf16": 2e-2}


def _setup(dims, X, dtype, axis="position", B=6, seed=5):
    ocfg = R.DecoderConfig(softmax_axis=axis, **dims)
    params = R.init_params(ocfg, seed, torch.float64, bias_scale=0.1)
    ap = TR.init_adaptor_params(ocfg, X, seed, torch.float64, bias_scale=0.1)
    x = TR.make_encoder_map(ocfg, B, X, seed).double()
    cap = synth.make_captions(ocfg, B, seed)
    eng = M.CaptionDecoderEngine(M.DecoderConfig(dtype=dtype, softmax_axis=axis, **dims), seed=None)
    eng.store.load(params)
    eng.attach_adaptor(X, seed=None).load(ap)
    return ocfg, params, ap, x, cap, eng


@pytest.mark.parametrize("dtype", ["fp32", "bf16"])
@pytest.mark.parametrize("dims,X,B", [(SMALL, 24, 3), (MID, 160, 6), (MID, 72, 1), (MID, 128, 5)])
def test_img2feature_forward(dtype, dims, X, B):
    ocfg, params, ap, x, cap, eng = _setup(dims, X, dtype, B=B)
    P_ref, g_ref = TR.img2feature(ap, x)
    P, g = eng.img2feature(x.float().cuda().contiguous(), want_outputs=True)
    torch.cuda.synchronize()
    assert P.shape == P_ref.shape and g.shape == g_ref.shape
    assert rel_err(P, P_ref) < TOL[dtype] and rel_err(g, g_ref) < TOL[dtype]
    assert float(P.min()) >= 0.0 and float(g.min()) >= 0.0
    # 4-D encoder output [B, X, 7, 7]-style input is reshaped like model:195
    if dims["num_positions"] == 49:
        P4, _ = eng.img2feature(x.float().cuda().reshape(B, X, 7, 7), want_outputs=True)
        assert torch.equal(P4, P)


@pytest.mark.parametrize("dtype", ["fp32", "bf16"])
@pytest.mark.parametrize("dims,X,B", [(SMALL, 24, 4), (MID, 160, 6), (MID, 128, 6)])
def test_train_from_encoder_map(dtype, dims, X, B):
    """adaptor + decoder + loss from the encoder map: loss, logits, every decoder / adaptor gradient and d(encoder map)"""
    ocfg, params, ap, x, cap, eng = _setup(dims, X, dtype, B=B)
    loss_ref, logits_ref, grads_ref, _ = TR.loss_and_grads_from_map(params, ap, ocfg, x, cap)
    xd, capd = x.float().cuda().contiguous(), cap.cuda()
    loss_buf, logits = eng.train_forward_map(xd, capd, want_logits=True)
    dmap = eng.train_backward_map()
    torch.cuda.synchronize()
    tol = TOL[dtype]
    assert abs(float(loss_buf[2]) - float(loss_ref)) < tol * max(1.0, abs(float(loss_ref)))
    assert rel_err(logits, logits_ref) < tol
    errs = {n: rel_err(eng.adaptor.grads[n], grads_ref[n]) for n in TR.ADAPTOR_NAMES}
    errs["enc_map"] = rel_err(dmap, grads_ref["enc_map"])
    for n in ("lstm_w", "fc_v_w", "fc_k_w", "word_embedding", "fc_o_w"):
        errs[n] = rel_err(eng.store.grads[n], grads_ref[n])
    bad = {k: v for k, v in errs.items() if not v < tol}
    assert not bad, bad
    # the staged path and the explicit-feature path are the same computation
    P, g = eng.img2feature(xd, want_outputs=True)
    lb2, logits2 = eng.train_forward(g, P, capd, want_logits=True)
    torch.cuda.synchronize()
    assert rel_err(logits2, logits) < (1e-6 if dtype == "fp32" else 1e-2)


def test_staged_features_errors_and_decode():
    ocfg, params, ap, x, cap, eng = _setup(MID, 160, "fp32", B=5)
    xd = x.float().cuda().contiguous()
    with pytest.raises(M.CapdecError):                      # nothing staged yet
        eng._decode(None, None, 5, 1, False, True)
    P, g = eng.img2feature(xd, want_outputs=True)
    ids_a, sc_a, f32_a = eng.decode_map(xd, beam=3, finish_on_stop=True)
    ids_b, sc_b, f32_b = eng.decode(g, P, beam=3, finish_on_stop=True)
    torch.cuda.synchronize()
    assert torch.equal(ids_a, ids_b) and torch.equal(f32_a, f32_b)
    P_ref, g_ref = TR.img2feature(ap, x)
    ref_ids = R.greedy_decode(params, ocfg, g_ref, P_ref)
    got = eng.decode_map(xd, beam=1)[2]
    assert (got.cpu() == ref_ids).float().mean() > 0.95
    with pytest.raises(M.CapdecError):                      # staged for 5 images, asked for 4
        eng._decode(None, None, 4, 1, False, True)
    with pytest.raises(M.CapdecError):                      # adaptor backward without a staged forward/backward
        eng.img2feature(xd)
        eng.train_backward_map()
    with pytest.raises(ValueError):
        eng.img2feature(xd[:, :8])
    model = M.ImageCaptionModel(eng)
    emb, glob = model._img2feature(xd)
    assert rel_err(emb, P_ref) < 1e-3 and rel_err(glob, g_ref) < 1e-3
    assert torch.equal(model.eval_network_from_map(xd), got)


def test_autograd_from_encoder_map():
    ocfg, params, ap, x, cap, eng = _setup(SMALL, 24, "fp32", B=4)
    _, _, grads_ref, _ = TR.loss_and_grads_from_map(params, ap, ocfg, x, cap)
    model = M.ImageCaptionModel(eng)
    xd = x.float().cuda().contiguous().requires_grad_(True)
    loss = model.training_network_from_map(xd, cap.cuda())
    (2.0 * loss).backward()
    torch.cuda.synchronize()
    assert rel_err(xd.grad, 2.0 * grads_ref["enc_map"]) < 1e-3
    assert rel_err(eng.adaptor.flat.grad[:24 * 64].view(24, 64), 2.0 * grads_ref["fc_img_w"]) < 1e-3


# ---- Adam -----------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n,offset,clip", [(1 << 20, 0, 0.0), (1000003, 0, 0.0), (4099, 1, 0.0), (3, 0, 0.0), (70001, 0, 0.05)])
def test_adam_step_matches_paddle18_rule(n, offset, clip):
    gen = torch.Generator().manual_seed(n)
    base = [torch.randn(n + 8, generator=gen) for _ in range(2)] + [torch.rand(n + 8, generator=gen) * 1e-3 for _ in range(2)]
    p0, g0, m0, v0 = [t[offset:offset + n].clone() for t in base]
    dev = [t.cuda()[offset:offset + n] for t in base]           # offset 1: unaligned views -> scalar kernel
    pn, gn, mn, vn = [t.numpy().copy() for t in (p0, g0, m0, v0)]
    lib = M.load_library()
    from myimagecaptioningmodel_b200._capi import check, ptr, stream_ptr
    b1p, b2p = np.float32(0.9), np.float32(0.999)
    for step in range(3):
        check(lib.capdec_adam_step(ptr(dev[0]), ptr(dev[1]), ptr(dev[2]), ptr(dev[3]), n, 1e-3, 0.9, 0.999, 1e-8,
                                   float(b1p), float(b2p), clip, stream_ptr()), "capdec_adam_step")
        b1p, b2p = TR.adam_step_numpy(pn, gn, mn, vn, 1e-3, beta1_pow=b1p, beta2_pow=b2p, clip=clip)
    torch.cuda.synchronize()
    assert np.abs(dev[0].cpu().numpy() - pn).max() < 1e-6
    assert rel_err(dev[2], torch.from_numpy(mn)) < 1e-6 and rel_err(dev[3], torch.from_numpy(vn)) < 1e-6
    assert torch.equal(dev[1].cpu(), g0)                        # the gradient is read-only


def test_adam_rejects_bad_arguments():
    lib = M.load_library()
    t = torch.zeros(8, device="cuda")
    from myimagecaptioningmodel_b200._capi import ptr, stream_ptr
    assert lib.capdec_adam_step(ptr(t), None, ptr(t), ptr(t), 8, 1e-3, 0.9, 0.999, 1e-8, 0.9, 0.999, 0.0, stream_ptr()) != 0
    assert lib.capdec_adam_step(ptr(t), ptr(t), ptr(t), ptr(t), 8, 1e-3, 0.9, 0.999, 1e-8, 1.0, 0.999, 0.0, stream_ptr()) != 0
    assert lib.capdec_adam_step(None, None, None, None, 0, 1e-3, 0.9, 0.999, 1e-8, 0.9, 0.999, 0.0, stream_ptr()) == 0
    with pytest.raises(ValueError):
        M.Adam([(t, torch.zeros(4, device="cuda"))])


def test_three_optimisation_steps_follow_the_oracle():
    """train.py:132-139 in miniature (fp32 storage): forward, backward, Adam on decoder + adaptor, re-bind, repeat; the loss
    trajectory and the final parameters follow the oracle driven by the NumPy Adam."""
    ocfg, params, ap, x, cap, eng = _setup(SMALL, 24, "fp32", B=4)
    opt, _ = M.get_optimizer(eng, {"learning_rate": 1e-2})
    assert len(opt.buffers) == 2
    pn = {k: v.clone() for k, v in params.items()}
    an = {k: v.clone() for k, v in ap.items()}
    state = {k: (np.zeros(v.shape, np.float32), np.zeros(v.shape, np.float32)) for k, v in {**pn, **an}.items()}
    b1p = b2p = None
    xd, capd = x.float().cuda().contiguous(), cap.cuda()
    losses, losses_ref = [], []
    for step in range(3):
        eng.bind_params()
        eng.bind_adaptor()
        lb, _ = eng.train_forward_map(xd, capd)
        eng.train_backward_map(want_dmap=False)
        losses.append(float(lb[2]))
        assert opt.step() == 1e-2
        lref, _, gref, _ = TR.loss_and_grads_from_map(pn, an, ocfg, x, cap)
        losses_ref.append(float(lref))
        nb1 = nb2 = None
        for k in list(pn) + list(an):
            tgt = pn if k in pn else an
            arr = tgt[k].float().numpy().copy()
            nb1, nb2 = TR.adam_step_numpy(arr, gref[k].float().numpy(), state[k][0], state[k][1], 1e-2,
                                          beta1_pow=b1p, beta2_pow=b2p)
            tgt[k] = torch.from_numpy(arr).double()
        b1p, b2p = nb1, nb2
    torch.cuda.synchronize()
    assert losses_ref[2] < losses_ref[0]                        # it does optimise
    for a, b in zip(losses, losses_ref):
        assert abs(a - b) < 2e-3 * abs(b), (losses, losses_ref)
    # sign-like first steps amplify 1e-6 gradient noise where |g| ~ eps; compare where the update is well conditioned
    for k in ("fc_o_w", "lstm_w", "fc_img_w"):
        got = (eng.store.params[k] if k in eng.store.params else eng.adaptor.params[k]).cpu().double()
        want = pn[k] if k in pn else an[k]
        assert float((got - want).abs().mean()) < 1e-3, k
    assert opt.steps == 3 and abs(float(opt.beta1_pow) - 0.9 ** 4) < 1e-6


def test_grouped_update_beside_the_backward_pass_equals_a_plain_step():
    """DataParallelTrainer(optimizer=...) applies Adam gradient group by gradient group on the side stream as the groups
    become final.  The update is element-wise, so replaying a plain whole-buffer Adam step from the parameters before the
    call and the gradients the call left behind must give the same parameters and moments bit for bit -- which also shows
    that no group was updated before its gradient was final."""
    cfg = M.DecoderConfig(dtype="bf16", softmax_axis="position", **MID)
    ocfg = R.DecoderConfig(softmax_axis="position", **MID)
    x = TR.make_encoder_map(ocfg, 8, 128, 3).cuda()
    g, P, cap = synth.make_batch(ocfg, 8, 3)
    g, P, cap = g.cuda(), P.cuda(), cap.cuda()
    eng = M.CaptionDecoderEngine(cfg, seed=4, bias_scale=0.1)
    eng.attach_adaptor(128, seed=4, bias_scale=0.1)
    opt = M.get_optimizer(eng, {"learning_rate": 1e-3, "gradient_clip": 0.5})[0]
    tr = M.DataParallelTrainer(eng)
    stores = (eng.store, eng.adaptor)
    for step in range(3):
        eng.bind_params()
        eng.bind_adaptor()
        before = [(st.flat.clone(), m.clone(), v.clone()) for st, m, v in zip(stores, opt.moment1, opt.moment2)]
        b1p, b2p = float(opt.beta1_pow), float(opt.beta2_pow)
        tr.forward_backward_map(x, cap, optimizer=opt)
        torch.cuda.synchronize()
        assert opt.steps == step + 1
        for (p0, m0, v0), st, m, v in zip(before, stores, opt.moment1, opt.moment2):
            replay = M.Adam([(p0, st.flat_grad)], 1e-3, gradient_clip=0.5)
            replay.moment1[0].copy_(m0)
            replay.moment2[0].copy_(v0)
            replay.beta1_pow, replay.beta2_pow = np.float32(b1p), np.float32(b2p)
            replay.step()
            torch.cuda.synchronize()
            assert float(st.flat_grad.abs().max()) > 0
            assert torch.equal(p0, st.flat.detach()) and torch.equal(replay.moment1[0], m) and torch.equal(replay.moment2[0], v)
    # decoder only, with a one-buffer optimizer
    o0 = M.Adam([(eng.store.flat, eng.store.flat_grad)], 1e-3)
    p0 = eng.store.flat.detach().clone()
    eng.bind_params()
    tr.forward_backward(g, P, cap, optimizer=o0)
    replay = M.Adam([(p0, eng.store.flat_grad)], 1e-3)
    replay.step()
    torch.cuda.synchronize()
    assert torch.equal(p0, eng.store.flat.detach()) and torch.equal(replay.moment2[0], o0.moment2[0]) and o0.steps == 1
    with pytest.raises(ValueError):        # an optimizer over the wrong buffers
        tr.forward_backward(g, P, cap, optimizer=opt)


def test_golden_fixture_of_the_widened_rows(golden_dir):
    """the CUDA path against tests/golden/train_small.npz (adaptor + decoder loss / gradients, two clipped Adam steps)"""
    import os
    z = np.load(os.path.join(golden_dir, "train_small.npz"))
    for dtype in ("bf16", "fp32"):
        eng = M.CaptionDecoderEngine(M.DecoderConfig(dtype=dtype, softmax_axis="position", **SMALL), seed=None)
        eng.store.load({n: z["param_" + n] for n in M.PARAM_NAMES})
        eng.attach_adaptor(24, seed=None).load({n: z["param_" + n] for n in TR.ADAPTOR_NAMES})
        x, cap = torch.from_numpy(z["enc_map"]).cuda(), torch.from_numpy(z["caption"]).cuda()
        P, g = eng.img2feature(x, want_outputs=True)
        lb, logits = eng.train_forward_map(x, cap, want_logits=True)
        dmap = eng.train_backward_map()
        torch.cuda.synchronize()
        tol = TOL[dtype]
        assert rel_err(P, torch.from_numpy(z["P"])) < tol and rel_err(g, torch.from_numpy(z["g"])) < tol
        assert abs(float(lb[2]) - float(z["loss"])) < tol * float(z["loss"])
        assert rel_err(logits, torch.from_numpy(z["logits"])) < tol
        if dtype == "bf16":
            # The gradients of this tiny problem (5 captions, 24 target tokens, H = 64, weight gradients contracted over
            # 5 rows) are sums of a handful of bf16-rounded, partly cancelling terms: measured 4e-2 (fc_img_w) to 1e-1
            # (fc_glob_w) against the fp64 fixture.  bf16 gradient parity at 2e-2 is asserted on the larger shapes of
            # test_train_from_encoder_map; the fixture pins the gradients in fp32 storage (1e-3) only.
            continue
        for n in TR.ADAPTOR_NAMES:
            assert rel_err(eng.adaptor.grads[n], torch.from_numpy(z["grad_" + n])) < tol, (dtype, n)
        assert rel_err(dmap, torch.from_numpy(z["grad_enc_map"])) < tol
        assert rel_err(eng.store.grads["lstm_w"], torch.from_numpy(z["grad_lstm_w"])) < tol
    # Adam: the fixture's two steps (its own gradient, then half of it; clip-by-value 0.01), fp32 arithmetic on both sides
    p = torch.from_numpy(z["param_fc_img_w"]).cuda().contiguous()
    gr = torch.from_numpy(z["grad_fc_img_w"]).cuda().contiguous()
    opt = M.Adam([(p.view(-1), gr.view(-1))], 1e-3, gradient_clip=0.01)
    opt.step()
    gr.mul_(0.5)
    opt.step()
    torch.cuda.synchronize()
    assert float((p.cpu() - torch.from_numpy(z["adam_p2"])).abs().max()) < 1e-6
    assert rel_err(opt.moment1[0].view_as(p), torch.from_numpy(z["adam_m2"])) < 1e-5
    assert rel_err(opt.moment2[0].view_as(p), torch.from_numpy(z["adam_v2"])) < 1e-5
