"""-m gpu: capdec_train_fwd with `lengths` (packed rows, SURVEY.md F5 / VERDICT r01 "next" #2).

The reference runs all 34 steps for every sample and masks the loss with target != <pad> (model:163-171); rows past the last
target of a caption contribute exactly nothing to the loss or to any gradient, and nothing downstream of (h_t, c_t) feeds
back into the recurrence.  With `lengths` the time-batched head runs on the live rows only.  Checked here:
  * packed == un-packed on the same inputs: row-wise results are the same arithmetic on the same numbers; only sums over
    rows (weight gradients, the loss) change their fp32 summation order -> 1e-5 (fp32 storage) / 2e-3 (bf16: the packed
    weight-gradient GEMMs split K differently and bf16 operands are summed in another order);
  * packed == oracle at the stated tolerances, small and full size (B = 256), both storage types, both soft-max axes;
  * ragged cases: an empty caption ([start, stop]: one target), a full-length caption, holes (a pad id inside a caption),
    every sample the same length (nothing to drop -> the un-packed schedule runs), a batch with a single live row;
  * a `lengths` that does not describe the captions is reported (CAPDEC_ERR_LENGTHS), never dereferenced out of bounds;
  * data-parallel trainer and encoder-map entry points pass `lengths` through."""
import numpy as np
import pytest
import torch

import myimagecaptioningmodel_b200 as M
from gpu_util import max_err, rel_err
from myimagecaptioningmodel_b200 import _capi
from oracle import decoder_ref as R
from oracle import synth
from oracle import train_ref as TR
from test_headline_parity_gpu import check_train_against, full_b256  # noqa: F401  (fixture)

pytestmark = pytest.mark.gpu

SMALL = dict(hidden_dim=64, embedding_size=32, vocab_size=101, num_positions=5, sentence_length=8, infer_max_length=9)
MID = dict(hidden_dim=256, embedding_size=64, vocab_size=1000, num_positions=49, sentence_length=12, infer_max_length=12)
TOL = {"fp32": 1e-3, "bf16": 2e-2}


def _engine(dims, axis, dtype, params):
    eng = M.CaptionDecoderEngine(M.DecoderConfig(dtype=dtype, softmax_axis=axis, **dims), seed=None)
    eng.store.load(params)
    return eng


def _run(eng, g, P, cap, lengths):
    n0 = eng.launch_count
    lb, _ = eng.train_forward(g, P, cap, lengths=lengths)
    dg, dP = eng.train_backward()
    eng.check_status()
    return dict(loss=lb.cpu().clone(), dg=dg.clone(), dP=dP.clone(), flat=eng.store.flat_grad.clone(),
                grads={k: v.clone() for k, v in eng.store.grads.items()}, launches=eng.launch_count - n0)


def _ragged_captions(ocfg, B, seed):
    _, _, cap = synth.make_batch(ocfg, B, seed)
    S = ocfg.sentence_length
    cap[0, :] = 0
    cap[0, 0], cap[0, 1] = ocfg.start_idx, ocfg.stop_idx               # empty caption: one target
    if B > 1:
        cap[1, 0] = ocfg.start_idx
        cap[1, 1:-1] = torch.arange(4, 4 + S - 2) % (ocfg.vocab_size - 4) + 4
        cap[1, -1] = ocfg.stop_idx                                     # full length: no padding
    if B > 2:
        cap[2, 2] = 0                                                  # a hole: pad id inside the caption
    return cap


@pytest.mark.parametrize("dims,B", [(SMALL, 7), (MID, 9), (MID, 140), (MID, 300)])
@pytest.mark.parametrize("axis", ["position", "feature"])
@pytest.mark.parametrize("dtype", ["fp32", "bf16"])
def test_packed_equals_unpacked_and_oracle(dims, B, axis, dtype):
    ocfg = R.DecoderConfig(softmax_axis=axis, **dims)
    params = R.init_params(ocfg, 7, torch.float32, bias_scale=0.1)
    g, P, _ = synth.make_batch(ocfg, B, 21)
    cap = _ragged_captions(ocfg, B, 21)
    lengths = M.caption_lengths(cap)
    assert lengths[0] == 2 and lengths[1] == ocfg.sentence_length
    eng = _engine(dims, axis, dtype, params)
    gd, Pd, capd = g.cuda(), P.cuda(), cap.cuda()
    a = _run(eng, gd, Pd, capd, None)
    b = _run(eng, gd, Pd, capd, lengths)
    T = ocfg.sentence_length - 1
    live = int(np.clip(lengths - 1, 0, T).sum())
    assert live < 0.95 * B * T                       # the packed schedule really ran ...
    assert b["launches"] != a["launches"]            # ... (it has gather / scatter launches of its own)
    assert torch.equal(a["loss"][1], b["loss"][1])   # token count: exact
    tol = 1e-5 if dtype == "fp32" else 2e-3
    assert abs(float(a["loss"][2]) - float(b["loss"][2])) < tol * abs(float(a["loss"][2]))
    assert rel_err(b["dg"], a["dg"]) < tol and rel_err(b["dP"], a["dP"]) < tol
    worst = {}
    for name in M.PARAM_NAMES:
        if float(a["grads"][name].abs().max()) == 0.0:
            assert float(b["grads"][name].abs().max()) == 0.0, name        # dead parameters (feature axis) stay exactly 0
        elif name != "fc_a_b":
            worst[name] = rel_err(b["grads"][name], a["grads"][name])
    assert all(v < (1e-5 if dtype == "fp32" else 5e-3) for v in worst.values()), worst
    # against the oracle (fp64 at these sizes)
    p64 = {k: v.double() for k, v in params.items()}
    loss_ref, _, _, grads_ref = R.loss_and_grads(p64, ocfg, g.double(), P.double(), cap)
    assert abs(float(b["loss"][2]) - float(loss_ref)) < TOL[dtype] * abs(float(loss_ref))
    gtol = TOL[dtype] if B > 8 else (5e-2 if dtype == "bf16" else TOL[dtype])   # few targets: see test_decoder_gpu
    for name in M.PARAM_NAMES:
        ref = grads_ref[name]
        if float(ref.abs().max()) == 0.0 or name == "fc_a_b":
            continue
        # fc_s_b (bias of the sentinel score projection) is a sum of nearly cancelling bf16-rounded terms: its relative
        # error depends on how many rows are summed in which order (1.7e-2 .. 7e-2 in the data-parallel check as well,
        # DESIGN.md section 6); every other tensor keeps the 2e-2 bound
        bound = 2.5 * gtol if (name == "fc_s_b" and dtype == "bf16") else gtol
        assert rel_err(b["grads"][name], ref) < bound, (name, rel_err(b["grads"][name], ref))
    assert rel_err(b["dg"], grads_ref["global_img_feat"]) < gtol and rel_err(b["dP"], grads_ref["p_img_feat"]) < gtol


@pytest.mark.parametrize("dtype", ["bf16", "fp32"])
def test_packed_full_size_batch256_vs_oracle(full_b256, dtype):  # noqa: F811
    """BASELINE.json configs[1] with lengths: 4944 of 8704 rows are live at the bench's batch"""
    eng = _engine({}, "position", dtype, full_b256["params"])
    lengths = M.caption_lengths(full_b256["cap"])
    check_train_against(eng, full_b256, dtype, want_logits=False, lengths=lengths)


def test_lengths_with_logits_requested_runs_the_full_schedule():
    """the per-step logits of every row are a return value of Decoder.call (model:129-130): nothing may be dropped"""
    ocfg = R.DecoderConfig(softmax_axis="position", **MID)
    params = R.init_params(ocfg, 3, torch.float32, bias_scale=0.1)
    g, P, cap = synth.make_batch(ocfg, 6, 5)
    eng = _engine(MID, "position", "fp32", params)
    lb, logits = eng.train_forward(g.cuda(), P.cuda(), cap.cuda(), want_logits=True, lengths=M.caption_lengths(cap))
    ref_logits, _ = R.decoder_train(params, ocfg, g, P, cap[:, :-1])
    assert rel_err(logits, ref_logits) < 1e-3


def test_uniform_lengths_and_single_live_row():
    ocfg = R.DecoderConfig(softmax_axis="position", **MID)
    params = R.init_params(ocfg, 9, torch.float32, bias_scale=0.1)
    eng = _engine(MID, "position", "bf16", params)
    S = ocfg.sentence_length
    g, P, cap = synth.make_batch(ocfg, 5, 8)
    # nothing to drop: every caption is full length
    cap[:, 0] = ocfg.start_idx
    cap[:, 1:-1] = 7
    cap[:, -1] = ocfg.stop_idx
    _run(eng, g.cuda(), P.cuda(), cap.cuda(), None)           # (the first call also binds the parameters)
    a = _run(eng, g.cuda(), P.cuda(), cap.cuda(), None)
    b = _run(eng, g.cuda(), P.cuda(), cap.cuda(), M.caption_lengths(cap))
    assert a["launches"] == b["launches"] and rel_err(b["flat"], a["flat"]) < 1e-5
    # one live row in the whole batch: [start, stop] in sample 3, everything else padding
    cap[:] = 0
    cap[3, 0], cap[3, 1] = ocfg.start_idx, ocfg.stop_idx
    lengths = M.caption_lengths(cap)
    assert lengths.tolist() == [0, 0, 0, 2, 0]
    a = _run(eng, g.cuda(), P.cuda(), cap.cuda(), None)
    b = _run(eng, g.cuda(), P.cuda(), cap.cuda(), lengths)
    assert float(b["loss"][1]) == 1.0 and abs(float(a["loss"][2]) - float(b["loss"][2])) < 2e-3 * float(a["loss"][2])
    assert rel_err(b["flat"], a["flat"]) < 5e-3 and torch.isfinite(b["flat"]).all()


def test_wrong_lengths_are_reported():
    ocfg = R.DecoderConfig(softmax_axis="position", **MID)
    params = R.init_params(ocfg, 9, torch.float32, bias_scale=0.1)
    eng = _engine(MID, "position", "bf16", params)
    g, P, cap = synth.make_batch(ocfg, 12, 4)
    good = M.caption_lengths(cap)
    for bad in (np.maximum(good - 2, 0), good + 1 - (good == ocfg.sentence_length)):   # too short / too long
        if (bad == good).all():
            continue
        with pytest.raises(_capi.CapdecError) as ei:       # reported by the next call on the handle, at the latest by status
            eng.train_forward(g.cuda(), P.cuda(), cap.cuda(), lengths=bad)
            eng.train_backward()
            eng.check_status()
        assert "lengths" in str(ei.value)
        torch.cuda.synchronize()
        try:
            eng.check_status()                                # (the backward pass may have been skipped: drain the report)
        except _capi.CapdecError:
            pass
    with pytest.raises(ValueError):
        eng.train_forward(g.cuda(), P.cuda(), cap.cuda(), lengths=good[:-1])
    with pytest.raises(ValueError):
        eng.train_forward(g.cuda(), P.cuda(), cap.cuda(), lengths=torch.from_numpy(good).cuda())
    # and the handle is still good
    b = _run(eng, g.cuda(), P.cuda(), cap.cuda(), good)
    assert torch.isfinite(b["flat"]).all()


def test_lengths_through_the_trainer_and_the_encoder_map_entry():
    """DataParallelTrainer.forward_backward(_map)(lengths=...) + Adam: the packed step trains like the un-packed one"""
    cfg = M.DecoderConfig(dtype="bf16", softmax_axis="position", **MID)
    ocfg = R.DecoderConfig(softmax_axis="position", **MID)
    g, P, cap = synth.make_batch(cfg, 16, 5)
    x = M.synth.make_encoder_map(cfg, 16, 64, 6)
    lengths = M.caption_lengths(cap)
    res = {}
    for use in (False, True):
        eng = M.CaptionDecoderEngine(cfg, seed=4)
        eng.attach_adaptor(64, seed=4)
        opt, _ = M.get_optimizer(eng, {"learning_rate": 2e-3})
        tr = M.DataParallelTrainer(eng)
        out = []
        for _ in range(3):
            loss, dmap = tr.forward_backward_map(x.cuda(), cap.cuda(), want_dmap=True, optimizer=opt,
                                                 lengths=lengths if use else None)
            out.append(float(loss))
        eng.check_status()
        res[use] = (out, dmap.clone(), eng.adaptor.flat_grad.clone())
    assert all(abs(p - q) < 3e-3 * abs(q) for p, q in zip(res[True][0], res[False][0])), (res[True][0], res[False][0])
    assert res[True][0][2] < res[True][0][0]
    assert rel_err(res[True][1], res[False][1]) < 2e-2 and rel_err(res[True][2], res[False][2]) < 2e-2
    # one step from the map against the oracle
    ap = TR.init_adaptor_params(ocfg, 64, 3, torch.float32, bias_scale=0.1)
    params = R.init_params(ocfg, 3, torch.float32, bias_scale=0.1)
    loss_m, _, grads_m, _ = TR.loss_and_grads_from_map(params, ap, ocfg, x, cap)
    eng = _engine(MID, "position", "fp32", params)
    eng.attach_adaptor(64, seed=None).load(ap)
    lb, _ = eng.train_forward_map(x.cuda(), cap.cuda(), lengths=lengths)
    dmap = eng.train_backward_map()
    eng.check_status()
    assert abs(float(lb[2]) - float(loss_m)) < 1e-3 * float(loss_m)
    assert rel_err(dmap, grads_m["enc_map"]) < 1e-3 and rel_err(eng.adaptor.grads["fc_img_w"], grads_m["fc_img_w"]) < 1e-3
