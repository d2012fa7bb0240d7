"""-m gpu: the full decoder path through the C-ABI (capdec_train_fwd / capdec_train_bwd / capdec_decode) against
the CPU oracle on identical synthetic inputs and weights.

Tolerances (north_star): fp32 storage rel 1e-3, bf16 storage rel 2e-2 (Frobenius-relative per tensor);
token ids bit-exact wherever the oracle's decision margin exceeds the numerical noise."""
import os

import numpy as np
import pytest
import torch

import myimagecaptioningmodel_b200 as M
from gpu_util import rel_err
from oracle import decoder_ref as R
from oracle import synth

pytestmark = pytest.mark.gpu

SMALL = dict(hidden_dim=64, embedding_size=32, vocab_size=101, num_positions=5, sentence_length=8, infer_max_length=9)
MID = dict(hidden_dim=256, embedding_size=64, vocab_size=1000, num_positions=49, sentence_length=12, infer_max_length=12)
TOL = {"fp32": 1e-3, "bf16": 2e-2}


def _engine(dims, axis, dtype, params):
    cfg = M.DecoderConfig(dtype=dtype, softmax_axis=axis, **dims)
    eng = M.CaptionDecoderEngine(cfg, seed=None)
    eng.store.load(params)
    return eng


def _check_train(eng, ocfg, params, g, P, cap, tol, grad_tol=None, want_logits=True):
    grad_tol = tol if grad_tol is None else grad_tol
    loss_ref, logits_ref, alpha_ref, grads_ref = R.loss_and_grads(params, ocfg, g, P, cap)
    gd, Pd, capd = g.float().cuda().contiguous(), P.float().cuda().contiguous(), cap.cuda().contiguous()
    loss_buf, logits = eng.train_forward(gd, Pd, capd, want_logits=want_logits)
    dg, dP = eng.train_backward()
    torch.cuda.synchronize()
    lb = loss_buf.cpu()
    ntok = float((cap[:, 1:] != 0).sum())
    assert lb[1].item() == ntok
    assert abs(lb[2].item() - float(loss_ref)) < tol * max(1.0, abs(float(loss_ref)))
    if want_logits:
        assert logits.shape == logits_ref.shape
        assert rel_err(logits, logits_ref) < tol
    errs = {}
    for name in M.PARAM_NAMES:
        ref = grads_ref[name]
        got = eng.store.grads[name]
        if float(ref.abs().max()) == 0.0:
            assert float(got.abs().max()) == 0.0, name          # dead parameters stay exactly zero (H4)
        elif name == "fc_a_b":
            # d loss / d b_alpha = sum_j d e_j == 0 for a soft-max (shift invariance): both sides hold rounding
            # noise only, so compare absolutely against the scale of its sibling gradient d w_alpha
            assert float((got.cpu().double() - ref).abs().max()) < grad_tol * float(grads_ref["fc_a_w"].abs().max())
        else:
            errs[name] = rel_err(got, ref)
    errs["d_global"] = rel_err(dg, grads_ref["global_img_feat"])
    errs["d_p_img"] = rel_err(dP, grads_ref["p_img_feat"])
    bad = {k: v for k, v in errs.items() if not v < grad_tol}
    assert not bad, bad
    return errs


@pytest.mark.parametrize("axis", ["feature", "position"])
@pytest.mark.parametrize("dtype", ["fp32", "bf16"])
def test_golden_fixture_small(axis, dtype, golden_dir):
    z = np.load(os.path.join(golden_dir, "decoder_small_%s.npz" % axis))
    params = {k[len("param_"):]: torch.from_numpy(z[k]) for k in z.files if k.startswith("param_")}
    eng = _engine(SMALL, axis, dtype, params)
    g, P, cap = torch.from_numpy(z["g"]).cuda(), torch.from_numpy(z["P"]).cuda(), torch.from_numpy(z["caption"]).cuda()
    loss_buf, logits = eng.train_forward(g, P, cap, want_logits=True)
    eng.train_backward()
    torch.cuda.synchronize()
    tol = TOL[dtype]
    assert abs(float(loss_buf[2]) - float(z["loss"])) < tol * float(z["loss"])
    assert rel_err(logits, torch.from_numpy(z["logits"])) < tol
    for name in M.PARAM_NAMES:
        ref = torch.from_numpy(z["grad_" + name])
        if float(ref.abs().max()) == 0:
            assert float(eng.store.grads[name].abs().max()) == 0, name
        elif name == "fc_a_b":        # mathematically zero (soft-max shift invariance): absolute check
            assert float((eng.store.grads[name].cpu() - ref).abs().max()) < tol * float(z["grad_fc_a_w"].__abs__().max())
        else:
            assert rel_err(eng.store.grads[name], ref) < tol, name
    if dtype == "fp32":
        ids, scores, ids_f32 = eng.decode(g, P, beam=1)
        assert ids_f32.dtype == torch.float32                            # F9
        assert (ids_f32.cpu().numpy() == z["greedy"]).all()
        ids3, scores3, _ = eng.decode(g, P, beam=3, finish_on_stop=True)
        assert (ids3.cpu().numpy() == z["beam_ids"]).all()
        assert np.abs(scores3.cpu().numpy() - z["beam_scores"]).max() < 1e-3


@pytest.mark.parametrize("axis", ["feature", "position"])
@pytest.mark.parametrize("dtype", ["fp32", "bf16"])
def test_train_step_mid_size(axis, dtype):
    ocfg = R.DecoderConfig(softmax_axis=axis, **MID)
    params = R.init_params(ocfg, 7, torch.float32, bias_scale=0.1)
    g, P, cap = synth.make_batch(ocfg, 9, 21)
    cap[3, 2:] = 0
    cap[3, 2] = 3
    p64 = {k: v.double() for k, v in params.items()}
    eng = _engine(MID, axis, dtype, params)
    _check_train(eng, ocfg, p64, g.double(), P.double(), cap, TOL[dtype])


@pytest.mark.parametrize("dtype", ["fp32", "bf16"])
def test_train_step_reference_config_batch8(dtype):
    """BASELINE.json configs[0]: config.py defaults, batch 8, teacher-forced forward+backward."""
    ocfg = R.DecoderConfig(softmax_axis="position")
    params = R.init_params(ocfg, 3, torch.float32, bias_scale=0.1)
    g, P, cap = synth.make_batch(ocfg, 8, 20260)
    eng = _engine({}, "position", dtype, params)
    errs = _check_train(eng, ocfg, params, g, P, cap, TOL[dtype] if dtype == "bf16" else 2e-3)
    print(dtype, {k: "%.2e" % v for k, v in errs.items()})


@pytest.mark.parametrize("dims,B", [(MID, 9), (MID, 130), (MID, 300), ({}, 131), ({}, 300)])
def test_persistent_recurrence_matches_per_step_schedule(dims, B):
    """The persistent cooperative recurrence kernels (recurrent.cu: W_hh resident in shared memory, one launch per
    direction) against the per-step GEMM + cell schedule they replace (CAPDEC_FLAG_STEPWISE_RECURRENCE), on ragged batch
    tiles (B % 128 != 0, more batch tiles than CTA groups), and against the oracle at the small sizes."""
    ocfg = R.DecoderConfig(softmax_axis="position", **dims)
    params = R.init_params(ocfg, 11, torch.float32, bias_scale=0.1)
    g, P, cap = synth.make_batch(ocfg, B, 99)
    gd, Pd, capd = g.cuda(), P.cuda(), cap.cuda()
    out = {}
    engines = {}
    for mode in ("1", "0"):
        eng = M.CaptionDecoderEngine(M.DecoderConfig(dtype="bf16", softmax_axis="position",
                                                     stepwise_recurrence=(mode == "0"), **dims), seed=None)
        eng.store.load(params)
        engines[mode] = eng
        n0 = eng.launch_count
        loss_buf, logits = eng.train_forward(gd, Pd, capd, want_logits=True)
        dg, dP = eng.train_backward()
        torch.cuda.synchronize()
        out[mode] = dict(loss=float(loss_buf[2]), logits=logits.clone(), dg=dg.clone(), dP=dP.clone(),
                         grads={k: v.clone() for k, v in eng.store.grads.items()}, launches=eng.launch_count - n0)
    a, b = out["1"], out["0"]
    T = ocfg.sentence_length - 1
    assert b["launches"] - a["launches"] >= 4 * T - 8, (a["launches"], b["launches"])   # the persistent path really ran
    assert abs(a["loss"] - b["loss"]) < 2e-3 * abs(b["loss"])
    assert rel_err(a["logits"], b["logits"]) < 5e-3
    assert rel_err(a["dg"], b["dg"]) < 1e-2 and rel_err(a["dP"], b["dP"]) < 1e-2
    for name in M.PARAM_NAMES:
        if name == "fc_a_b" or float(b["grads"][name].abs().max()) == 0.0:
            continue
        assert rel_err(a["grads"][name], b["grads"][name]) < 1e-2, name
    if B <= 130 and dims:
        p64 = {k: v.double() for k, v in params.items()}
        _check_train(engines["1"], ocfg, p64, g.double(), P.double(), cap, TOL["bf16"])


@pytest.mark.parametrize("B,beam,finish", [(1, 1, False), (4, 3, True), (20, 2, True)])
def test_small_batch_decode_graph_replay(B, beam, finish):
    """Small decodes replay a captured CUDA graph of the step loop from the second call of a shape on: eager run, capture +
    replay, replay and a re-capture after new parameters were bound must all give bit-identical ids and scores; a
    different input through the same graph must match its own eager result."""
    ocfg = R.DecoderConfig(softmax_axis="position", **MID)
    params = R.init_params(ocfg, 21, torch.float32, bias_scale=0.1)
    g, P, _ = synth.make_batch(ocfg, B, 5)
    g2, P2, _ = synth.make_batch(ocfg, B, 6)
    eng = _engine(MID, "position", "bf16", params)
    gd, Pd, g2d, P2d = g.cuda(), P.cuda(), g2.cuda(), P2.cuda()
    n0 = eng.launch_count
    ids0, sc0, f0 = eng.decode(gd, Pd, beam=beam, finish_on_stop=finish)          # eager
    n1 = eng.launch_count
    ids1, sc1, f1 = eng.decode(gd, Pd, beam=beam, finish_on_stop=finish)          # captured + replayed
    ids2, sc2, f2 = eng.decode(gd, Pd, beam=beam, finish_on_stop=finish)          # replayed
    torch.cuda.synchronize()
    n3 = eng.launch_count
    assert n1 - n0 > 100 and abs((n3 - n1) - 2 * (n1 - n0)) <= 8                  # the launch accounting covers replays
    for a, b in ((ids0, ids1), (ids0, ids2), (sc0, sc1), (sc0, sc2), (f0, f1), (f0, f2)):
        assert torch.equal(a, b)
    idsB, scB, _ = eng.decode(g2d, P2d, beam=beam, finish_on_stop=finish)         # other images through the same graph
    eng.bind_params()                                                             # drops the graph
    idsC, scC, _ = eng.decode(g2d, P2d, beam=beam, finish_on_stop=finish)         # eager again
    idsD, scD, _ = eng.decode(g2d, P2d, beam=beam, finish_on_stop=finish)         # re-captured
    torch.cuda.synchronize()
    assert torch.equal(idsB, idsC) and torch.equal(scB, scC) and torch.equal(idsC, idsD) and torch.equal(scC, scD)
    assert not torch.equal(idsB, ids0) or B == 1


def _explained_by_margin(ids_gpu, ids_ref, margin, noise):
    """ids must agree up to the first step whose oracle decision margin is below the numerical noise"""
    bad = 0
    for b in range(ids_ref.shape[0]):
        for t in range(ids_ref.shape[1]):
            if ids_gpu[b, t] != ids_ref[b, t]:
                if margin[b, t] > noise:
                    bad += 1
                break
    return bad


@pytest.mark.parametrize("axis", ["feature", "position"])
@pytest.mark.parametrize("dtype", ["fp32", "bf16"])
def test_greedy_and_beam_decode(axis, dtype):
    ocfg = R.DecoderConfig(softmax_axis=axis, **MID)
    params = R.init_params(ocfg, 5, torch.float32, bias_scale=0.1)
    g, P, _ = synth.make_batch(ocfg, 16, 33)
    p64 = {k: v.double() for k, v in params.items()}
    eng = _engine(MID, axis, dtype, params)
    gd, Pd = g.cuda(), P.cuda()
    ref_ids, margin = R.greedy_decode(p64, ocfg, g.double(), P.double(), return_margin=True)
    ids, scores, ids_f32 = eng.decode(gd, Pd, beam=1)
    noise = 1e-4 if dtype == "fp32" else 0.15
    assert ids_f32.shape == (16, ocfg.infer_max_length) and ids_f32.dtype == torch.float32
    assert torch.equal(ids[:, 0].float(), ids_f32)
    assert _explained_by_margin(ids_f32.cpu().numpy(), ref_ids.numpy(), margin.numpy(), noise) == 0
    if dtype == "fp32":
        assert (ids_f32.cpu() == ref_ids).float().mean() > 0.95
    # beam 3 with finished-prefix handling
    b_ids, b_scores, b_margin = R.beam_search(p64, ocfg, g.double(), P.double(), 3, return_margin=True)
    ids3, scores3, _ = eng.decode(gd, Pd, beam=3, finish_on_stop=True)
    ids3, scores3 = ids3.cpu(), scores3.cpu()
    assert (scores3[:, :-1] >= scores3[:, 1:]).all()
    if dtype == "fp32":
        ok_rows = (b_margin.min(dim=1).values > 1e-3)
        assert ok_rows.float().mean() > 0.5
        assert torch.equal(ids3[ok_rows].long(), b_ids[ok_rows])
        assert float((scores3[ok_rows] - b_scores[ok_rows].float()).abs().max()) < 1e-2
    else:
        assert float((scores3[:, 0] - b_scores[:, 0].float()).abs().mean()) < 0.5
    # greedy == beam(1) without finished handling, bit for bit, on the device path itself
    ids1b, _, _ = eng.decode(gd, Pd, beam=1, finish_on_stop=False)
    assert torch.equal(ids1b, ids)
    # with finished handling the captions agree after evaluate.filter (eval:15-25)
    ids1f, _, _ = eng.decode(gd, Pd, beam=1, finish_on_stop=True)
    for a, b in zip(ids[:, 0].tolist(), ids1f[:, 0].tolist()):
        assert M.filter_caption(a) == M.filter_caption(b)


@pytest.mark.parametrize("B,beam", [(1, 8), (3, 5), (2, 1)])
def test_beam_search_edge_cases_bit_exact_fp32(B, beam):
    """fp32 path against the oracle's beam search on the shapes the domain makes awkward: one image, the widest beam the
    kernels accept (8), beams that finish early (stop token) and keep emitting padding, and scores sorted per image.
    Rows whose oracle selection margin exceeds the fp32 noise must match bit for bit (ids) and to 1e-3 (scores)."""
    ocfg = R.DecoderConfig(softmax_axis="position", **MID)
    params = R.init_params(ocfg, 17, torch.float32, bias_scale=0.1)
    # make <stop> likely so that finished-prefix handling is exercised: raise its output bias
    params["out_fc_bias"][ocfg.stop_idx] += 4.0
    g, P, _ = synth.make_batch(ocfg, B, 71)
    p64 = {k: v.double() for k, v in params.items()}
    eng = _engine(MID, "position", "fp32", params)
    ref_ids, ref_sc, margin = R.beam_search(p64, ocfg, g.double(), P.double(), beam, return_margin=True)
    ids, sc, ids_f32 = eng.decode(g.cuda(), P.cuda(), beam=beam, finish_on_stop=True)
    ids, sc = ids.cpu().long(), sc.cpu()
    assert ids.shape == (B, beam, ocfg.infer_max_length) and sc.shape == (B, beam)
    assert (sc[:, :-1] >= sc[:, 1:]).all()
    ok = margin.min(dim=1).values > 1e-3
    assert ok.any()
    assert torch.equal(ids[ok], ref_ids[ok])
    assert float((sc[ok] - ref_sc[ok].float()).abs().max()) < 1e-3
    # a finished beam emits padding from the step after <stop> on
    stops = (ids == ocfg.stop_idx)
    assert stops.any()
    for b in range(B):
        for r in range(beam):
            pos = stops[b, r].nonzero()
            if len(pos):
                assert (ids[b, r, int(pos[0]) + 1:] == ocfg.padding_idx).all()
    assert torch.equal(ids_f32.cpu().long(), ids[:, 0])            # the reference's float32 return = best beam


def test_beam_wider_than_live_candidates_and_tiny_vocab():
    """vocabulary of 5 tokens with beam 3: step 0 has one live prefix and 5 candidates, later steps at most 15; the
    kernels must never emit an out-of-range token or a duplicate hypothesis."""
    dims = dict(hidden_dim=128, embedding_size=32, vocab_size=5, num_positions=6, sentence_length=6, infer_max_length=7)
    ocfg = R.DecoderConfig(softmax_axis="position", **dims)
    params = R.init_params(ocfg, 3, torch.float32, bias_scale=0.3)
    g, P, _ = synth.make_batch(ocfg, 5, 9)
    p64 = {k: v.double() for k, v in params.items()}
    eng = _engine(dims, "position", "fp32", params)
    ref_ids, ref_sc, margin = R.beam_search(p64, ocfg, g.double(), P.double(), 3, finish_on_stop=False, return_margin=True)
    ids, sc, _ = eng.decode(g.cuda(), P.cuda(), beam=3, finish_on_stop=False)
    ids = ids.cpu().long()
    assert int(ids.min()) >= 0 and int(ids.max()) < 5
    ok = margin.min(dim=1).values > 1e-4
    assert torch.equal(ids[ok], ref_ids[ok])
    for b in range(5):
        hyps = {tuple(ids[b, r].tolist()) for r in range(3)}
        assert len(hyps) == 3
    with pytest.raises(M.CapdecError):
        eng.decode(g.cuda(), P.cuda(), beam=9)                       # beam > 8 is rejected, not truncated


def test_single_caption_and_degenerate_captions():
    """B = 1 and captions at the extremes of the tokenizer's layout ([2, 3, 0...]: no word at all; 33 words, no
    padding): the masked loss normalises by the true token count and padded steps contribute no gradient."""
    ocfg = R.DecoderConfig(softmax_axis="position", **MID)
    params = R.init_params(ocfg, 23, torch.float32, bias_scale=0.1)
    p64 = {k: v.double() for k, v in params.items()}
    for B in (1, 3):
        g, P, cap = synth.make_batch(ocfg, B, 100 + B)
        cap[0, :] = 0
        cap[0, 0], cap[0, 1] = ocfg.start_idx, ocfg.stop_idx          # empty caption: a single target token
        if B > 1:
            cap[1, 0] = ocfg.start_idx
            cap[1, 1:-1] = torch.arange(4, 4 + ocfg.sentence_length - 2)
            cap[1, -1] = ocfg.stop_idx                                 # full-length caption: no padding at all
        for dtype in ("fp32", "bf16"):
            eng = _engine(MID, "position", dtype, params)
            # with a handful of target tokens nothing averages the bf16 rounding of the activations: the small gradients
            # of the sentinel path (fc_s) sit at 2.6e-2 for B = 1 (one target token), so the bf16 gradient bound is 5e-2
            # here; loss and logits keep the 2e-2 bound, fp32 keeps 1e-3 throughout
            _check_train(eng, ocfg, p64, g.double(), P.double(), cap, TOL[dtype],
                         grad_tol=5e-2 if dtype == "bf16" else None)


def test_gradient_groups_cover_the_flat_buffer_once():
    """the ranges the data-parallel trainer all-reduces as train_bwd finalises them tile store.flat_grad exactly"""
    eng = M.CaptionDecoderEngine(M.DecoderConfig(dtype="bf16", **MID), seed=0)
    sl = sorted((lo, hi, ev) for ev, lo, hi in eng.grad_group_slices())
    assert sl[0][0] == 0 and sl[-1][1] == eng.store.numel_padded
    for (lo0, hi0, _), (lo1, hi1, _) in zip(sl[:-1], sl[1:]):
        assert hi0 == lo1 and lo0 < hi0
    assert {ev for _, _, ev in sl} == set(range(eng.N_GRAD_EVENTS))
    evs = [torch.cuda.Event() for _ in range(eng.N_GRAD_EVENTS)]
    for e in evs:
        e.record()
    eng.set_grad_ready_events(evs)
    ocfg = R.DecoderConfig(softmax_axis="position", **MID)
    g, P, cap = synth.make_batch(ocfg, 5, 3)
    eng.train_forward(g.cuda(), P.cuda(), cap.cuda())
    eng.train_backward()
    torch.cuda.synchronize()
    assert all(e.query() for e in evs)
    eng.set_grad_ready_events(None)


def test_reference_shaped_interface():
    """Decoder.call / ImageCaptionModel mirror model:40-135, 161-189 (shapes, dtypes, errors, autograd)."""
    ocfg = R.DecoderConfig(softmax_axis="position", **SMALL)
    params = R.init_params(ocfg, 1, torch.float32, bias_scale=0.1)
    eng = _engine(SMALL, "position", "fp32", params)
    g, P, cap = synth.make_batch(ocfg, 4, 2)
    gd, Pd, capd = g.cuda(), P.cuda(), cap.cuda()
    with pytest.raises(ValueError):
        M.Decoder(64, mode="test", engine=eng)
    dec = M.Decoder(64, mode="train", engine=eng)
    logits = dec.call(gd, Pd, None, words=capd[:, :-1].contiguous())
    ref_logits, _ = R.decoder_train(params, ocfg, g, P, cap[:, :-1])
    assert logits.shape == (4, 7, 101) and rel_err(logits, ref_logits) < 1e-3
    ids = M.Decoder(64, mode="eval", engine=eng).call(gd, Pd)
    assert ids.shape == (4, 9) and ids.dtype == torch.float32
    assert torch.equal(ids.cpu(), R.greedy_decode(params, ocfg, g, P))
    model = M.ImageCaptionModel(eng)
    gd.requires_grad_(True)
    loss = model.training_network(gd, Pd, capd)
    loss.backward()
    lref, _, _, gref = R.loss_and_grads(params, ocfg, g, P, cap)
    assert abs(float(loss.detach()) - float(lref)) < 1e-3
    assert rel_err(gd.grad, gref["global_img_feat"]) < 1e-3
    flat_ref = torch.cat([gref[n].reshape(-1) for n in M.PARAM_NAMES])
    got = torch.cat([eng.store.flat.grad[eng.store.offsets[n]:eng.store.offsets[n] + gref[n].numel()]
                     for n in M.PARAM_NAMES])
    assert rel_err(got, flat_ref) < 1e-3
    with pytest.raises(ValueError):
        eng.train_forward(gd.detach()[:, :32].contiguous(), Pd, capd)


def test_full_size_properties_batch256():
    """BASELINE.json configs[1] (B=256, bf16): size-independent properties instead of the (slow) oracle:
    the loss of a batch is the token-weighted mean of the losses of its halves (samples are independent), padded
    positions carry no gradient, and the data-parallel normalisation (external 1/ntok) reproduces the gradient."""
    cfg = M.DecoderConfig(dtype="bf16", softmax_axis="position")
    eng = M.CaptionDecoderEngine(cfg, seed=1)
    ocfg = R.DecoderConfig(softmax_axis="position")
    g, P, cap = synth.make_batch(ocfg, 256, 20260)
    gd, Pd, capd = g.cuda(), P.cuda(), cap.cuda()
    lb, _ = eng.train_forward(gd, Pd, capd)
    eng.train_backward(want_dg=False, want_dP=False)
    torch.cuda.synchronize()
    full = lb.cpu().clone()
    grad_full = eng.store.flat_grad.clone()
    assert torch.isfinite(grad_full).all() and float(full[2]) > 0
    assert full[1].item() == float((cap[:, 1:] != 0).sum())
    halves = []
    grad_sum = torch.zeros_like(grad_full)
    inv = (1.0 / full[1]).reshape(1).cuda()
    for lo, hi in ((0, 128), (128, 256)):
        lbh, _ = eng.train_forward(gd[lo:hi].contiguous(), Pd[lo:hi].contiguous(), capd[lo:hi].contiguous())
        eng.train_backward(1.0, inv, want_dg=False, want_dP=False)      # global-token normalisation
        torch.cuda.synchronize()
        halves.append(lbh.cpu().clone())
        grad_sum += eng.store.flat_grad
    assert abs(float(halves[0][0] + halves[1][0]) - float(full[0])) < 2e-3 * float(full[0])
    assert halves[0][1] + halves[1][1] == full[1]
    assert rel_err(grad_sum, grad_full) < 2e-2
    # a fully padded sample contributes nothing
    cap2 = cap.clone()
    cap2[0, 1:] = 0
    lb2, _ = eng.train_forward(gd, Pd, cap2.cuda())
    torch.cuda.synchronize()
    assert lb2[1].item() == float((cap2[:, 1:] != 0).sum())
