"""CPU tests of the oracle itself: the two independent restatements agree, structural known-answer tests
derived from the reference source, finite-difference gradients, golden fixtures."""
import os

import numpy as np
import pytest
import torch

from oracle import decoder_ref as R
from oracle import numpy_ref as N
from oracle import synth

TINY = dict(hidden_dim=8, embedding_size=4, vocab_size=11, num_positions=3, sentence_length=5, infer_max_length=6)
SMALL = dict(hidden_dim=64, embedding_size=32, vocab_size=101, num_positions=5, sentence_length=8, infer_max_length=9)


def _setup(dims, axis, B=4, seed=1, dtype=torch.float64):
    cfg = R.DecoderConfig(softmax_axis=axis, **dims)
    p = R.init_params(cfg, seed, dtype)
    g, P, cap = synth.make_batch(cfg, B, seed + 2, dtype)
    return cfg, p, g, P, cap


@pytest.mark.parametrize("axis", ["feature", "position"])
def test_two_restatements_agree_fp64(axis):
    cfg, p, g, P, cap = _setup(SMALL, axis)
    loss, logits, _ = R.training_network(p, cfg, g, P, cap)
    pn = {k: v.numpy() for k, v in p.items()}
    loss2, logits2 = N.training_loss(pn, g.numpy(), P.numpy(), cap.numpy(), hid_size=cfg.hidden_dim, softmax_axis=axis)
    assert abs(float(loss) - float(loss2)) < 1e-12
    assert np.abs(logits.numpy() - logits2).max() < 1e-12
    ids = R.greedy_decode(p, cfg, g, P).numpy()
    ids2 = N.decoder_call(pn, g.numpy(), P.numpy(), mode="eval", hid_size=cfg.hidden_dim,
                          infer_max_length=cfg.infer_max_length, softmax_axis=axis)
    assert ids.dtype == np.float32 and ids2.dtype == np.float32          # F9: float32 ids
    assert ids.shape == (4, cfg.infer_max_length)                        # never stops early (model:66-68)
    assert (ids == ids2).all()


def test_bad_mode_raises():
    with pytest.raises(ValueError):
        N.decoder_call({}, np.zeros((1, 8)), np.zeros((1, 3, 8)), mode="test")   # model:43-44


def test_feature_mode_alpha_is_one_and_context_is_mean():
    cfg, p, g, P, cap = _setup(SMALL, "feature")
    logits, alpha, aux = R.decoder_train(p, cfg, g, P, cap[:, :-1], return_aux=True)
    assert torch.equal(alpha, torch.ones_like(alpha))                   # F2
    img_feat, _ = R.image_projections(p, P)
    ctx = (img_feat.sum(1, keepdim=True) + aux["sentinel"]) / (cfg.num_positions + 1)   # F3
    assert torch.allclose(aux["context"], ctx, atol=1e-14)


def test_feature_mode_dead_parameters_have_zero_grad():
    cfg, p, g, P, cap = _setup(SMALL, "feature")
    _, _, _, grads = R.loss_and_grads(p, cfg, g, P, cap)
    for name in ("fc_k_w", "fc_k_b", "fc_g_w", "fc_g_b", "fc_s_w", "fc_s_b", "fc_a_w", "fc_a_b"):
        assert float(grads[name].abs().max()) == 0.0, name
    assert float(grads["fc_v_w"].abs().max()) > 0


def test_position_mode_alpha_is_a_distribution():
    cfg, p, g, P, cap = _setup(SMALL, "position")
    _, alpha = R.decoder_train(p, cfg, g, P, cap[:, :-1])
    assert alpha.shape == (4, cfg.train_steps, cfg.num_positions + 1)
    assert torch.allclose(alpha.sum(-1), torch.ones(4, cfg.train_steps, dtype=torch.float64), atol=1e-13)


def test_padding_row_is_zero_and_gets_no_lookup_gradient():
    cfg, p, g, P, cap = _setup(SMALL, "position")
    emb = R.embedding_function(p, torch.tensor([0, 5, 0]), cfg)
    assert float(emb[0].abs().max()) == 0 and float(emb[2].abs().max()) == 0
    # with the output projection bias-only (fc_e zeroed), row 0 of the tied table can only receive lookup gradient
    p2 = dict(p)
    p2["fc_e_w"] = torch.zeros_like(p["fc_e_w"])
    p2["fc_e_b"] = torch.zeros_like(p["fc_e_b"])
    cap2 = cap.clone()
    cap2[:, 2] = 0                                                     # feed <pad> as a source token
    _, _, _, grads = R.loss_and_grads(p2, cfg, g, P, cap2)
    assert float(grads["word_embedding"][0].abs().max()) == 0.0


def test_tied_embedding_gradient_is_lookup_plus_projection():
    cfg, p, g, P, cap = _setup(TINY, "position", B=3)
    _, _, _, grads = R.loss_and_grads(p, cfg, g, P, cap)
    # untie: separate tables for lookup and projection, gradients must add up (F6)
    ps = {k: v.detach().clone().requires_grad_(True) for k, v in p.items()}
    emb_out = p["word_embedding"].detach().clone().requires_grad_(True)

    def tied_off(params, x):
        proj = x @ params["fc_e_w"] + params["fc_e_b"]
        return proj @ emb_out.t() + params["out_fc_bias"]
    orig = R.weight_tying_fc
    R.weight_tying_fc = tied_off
    try:
        loss, _, _ = R.training_network(ps, cfg, g, P, cap)
        loss.backward()
    finally:
        R.weight_tying_fc = orig
    assert torch.allclose(ps["word_embedding"].grad + emb_out.grad, grads["word_embedding"], atol=1e-14)


def test_uniform_logits_loss_is_log_vocab():
    cfg = R.DecoderConfig(**SMALL)
    cap = synth.make_captions(cfg, 5, 3)
    logits = torch.zeros(5, cfg.train_steps, cfg.vocab_size, dtype=torch.float64)
    loss, _, ntok = R.masked_loss(cfg, logits, cap)
    assert abs(float(loss) - np.log(cfg.vocab_size)) < 1e-12
    assert int(ntok) == int((cap[:, 1:] != 0).sum())


def test_padded_targets_are_excluded():
    cfg, p, g, P, cap = _setup(SMALL, "position")
    loss_a, logits, _ = R.training_network(p, cfg, g, P, cap)
    # changing logits at padded target positions must not change the loss (model:167-171)
    mask = cap[:, 1:] == 0
    logits2 = logits.clone()
    logits2[mask] += 3.0
    loss_b, _, _ = R.masked_loss(cfg, logits2, cap)
    assert abs(float(loss_a) - float(loss_b)) < 1e-13


def test_first_argmax_breaks_ties_low():
    x = torch.tensor([[1.0, 3.0, 3.0, 2.0], [5.0, 5.0, 5.0, 5.0]])
    assert R.first_argmax(x).tolist() == [1, 0]


@pytest.mark.parametrize("axis", ["feature", "position"])
def test_greedy_equals_beam1(axis):
    cfg, p, g, P, cap = _setup(SMALL, axis, B=5)
    ids = R.greedy_decode(p, cfg, g, P)
    b_ids, _ = R.beam_search(p, cfg, g, P, 1, finish_on_stop=False)
    assert torch.equal(b_ids[:, 0].to(torch.float32), ids)
    # with finished-handling on, identical after evaluate.filter (eval:15-25)
    b_ids2, _ = R.beam_search(p, cfg, g, P, 1, finish_on_stop=True)
    for a, b in zip(ids.tolist(), b_ids2[:, 0].tolist()):
        assert R.filter_caption(a, cfg) == R.filter_caption(b, cfg)


def test_beam_scores_sorted_and_consistent():
    cfg, p, g, P, cap = _setup(SMALL, "position", B=5)
    ids, scores = R.beam_search(p, cfg, g, P, 3)
    assert ids.shape == (5, 3, cfg.infer_max_length)
    assert (scores[:, :-1] >= scores[:, 1:]).all()
    # the best beam of width 3 scores at least as well as the width-1 hypothesis
    _, s1 = R.beam_search(p, cfg, g, P, 1)
    assert (scores[:, 0] >= s1[:, 0] - 1e-12).all()
    # after <stop> only <pad> is emitted
    for b in range(5):
        for r in range(3):
            seq = ids[b, r].tolist()
            if cfg.stop_idx in seq:
                i = seq.index(cfg.stop_idx)
                assert all(t == cfg.padding_idx for t in seq[i + 1:])


@pytest.mark.parametrize("axis", ["feature", "position"])
def test_finite_difference_gradients(axis):
    cfg, p, g, P, cap = _setup(TINY, axis, B=2, seed=4)
    _, _, _, grads = R.loss_and_grads(p, cfg, g, P, cap)
    rng = np.random.RandomState(0)
    eps = 1e-6
    for name in ["lstm_w", "fc_xs_w", "fc_hs_w", "fc_p_w", "fc_o_w", "fc_e_w", "word_embedding", "fc_v_w", "fc_k_w",
                 "fc_g_w", "fc_s_w", "fc_a_w", "lstm_b", "out_fc_bias"]:
        flat = p[name].reshape(-1)
        for idx in rng.choice(flat.numel(), size=min(4, flat.numel()), replace=False):
            old = float(flat[idx])
            flat[idx] = old + eps
            lp, _, _ = R.training_network(p, cfg, g, P, cap)
            flat[idx] = old - eps
            lm, _, _ = R.training_network(p, cfg, g, P, cap)
            flat[idx] = old
            fd = (float(lp) - float(lm)) / (2 * eps)
            an = float(grads[name].reshape(-1)[idx])
            assert abs(fd - an) < 1e-6 + 1e-5 * abs(an), (name, idx, fd, an)


@pytest.mark.parametrize("axis", ["feature", "position"])
def test_golden_fixture(axis, golden_dir):
    z = np.load(os.path.join(golden_dir, "decoder_small_%s.npz" % axis))
    cfg = R.DecoderConfig(softmax_axis=axis, **SMALL)
    p = {k[len("param_"):]: torch.from_numpy(z[k]).double() for k in z.files if k.startswith("param_")}
    g, P, cap = torch.from_numpy(z["g"]).double(), torch.from_numpy(z["P"]).double(), torch.from_numpy(z["caption"])
    loss, logits, alpha, grads = R.loss_and_grads(p, cfg, g, P, cap)
    assert abs(float(loss) - float(z["loss"])) < 1e-9
    assert np.abs(logits.numpy() - z["logits"]).max() < 1e-5
    assert np.abs(alpha.numpy() - z["alpha"]).max() < 1e-6
    for k in z.files:
        if k.startswith("grad_"):
            ref = z[k]
            got = grads[k[len("grad_"):]].numpy()
            assert np.abs(got - ref).max() <= 1e-6 * max(1.0, np.abs(ref).max()), k
    assert (R.greedy_decode(p, cfg, g, P).numpy() == z["greedy"]).all()
    ids, scores = R.beam_search(p, cfg, g, P, 3)
    assert (ids.numpy() == z["beam_ids"]).all()
    assert np.abs(scores.numpy() - z["beam_scores"]).max() < 1e-5
    # the literal numpy restatement reproduces the same fixture
    pn = {k: v.numpy() for k, v in p.items()}
    loss2, _ = N.training_loss(pn, g.numpy(), P.numpy(), cap.numpy(), hid_size=cfg.hidden_dim, softmax_axis=axis)
    assert abs(float(loss2) - float(z["loss"])) < 1e-9


def test_lstm_step_equals_the_library_lstm_cell_with_paddles_gate_order():
    """model:87-88 layers.lstm_unit: gate column blocks i, f, o, g over concat([x, h]).  torch.nn.LSTMCell is an independent
    library implementation of the same cell with gate ROWS i, f, g, o and separate input / hidden weights: with the columns
    re-ordered and split, its (h', c') must equal the oracle's, step after step."""
    cfg, p, g, P, cap = _setup(SMALL, "position", B=5)
    H, E = cfg.hidden_dim, cfg.embedding_size
    X = E + H
    W, b = p["lstm_w"], p["lstm_b"]                                     # [X + H, 4H], [4H]: i | f | o | g
    perm = torch.cat([torch.arange(0, 2 * H), torch.arange(3 * H, 4 * H), torch.arange(2 * H, 3 * H)])   # -> i, f, g, o
    cell = torch.nn.LSTMCell(X, H, bias=True).double()
    with torch.no_grad():
        cell.weight_ih.copy_(W[:X, perm].t())
        cell.weight_hh.copy_(W[X:, perm].t())
        cell.bias_ih.copy_(b[perm])
        cell.bias_hh.zero_()
    img_feat, img_feat_emb = R.image_projections(p, P)
    hid = torch.zeros(5, H, dtype=torch.float64)
    c = torch.zeros(5, H, dtype=torch.float64)
    h_lib, c_lib = hid.clone(), c.clone()
    for t in range(cfg.sentence_length - 1):
        word = cap[:, t]
        xt = torch.cat([R.embedding_function(p, word, cfg), g], dim=-1)
        h_lib, c_lib = cell(xt, (h_lib, c_lib))
        _, hid, c, _, _ = R.decoder_step(p, cfg, word, g, img_feat, img_feat_emb, hid, c)
        assert torch.allclose(hid, h_lib, atol=1e-12) and torch.allclose(c, c_lib, atol=1e-12)


def test_masked_loss_equals_the_library_ignore_index_cross_entropy():
    """model:163-171, 180-183, 205-212: sum(ce * [target != pad]) / sum([target != pad]) is what
    torch.nn.functional.cross_entropy(ignore_index=pad, reduction='mean') computes."""
    cfg, p, g, P, cap = _setup(SMALL, "position", B=6)
    logits, _ = R.decoder_train(p, cfg, g, P, cap[:, :-1])
    loss, loss_sum, ntok = R.masked_loss(cfg, logits, cap)
    ref = torch.nn.functional.cross_entropy(logits.reshape(-1, cfg.vocab_size), cap[:, 1:].reshape(-1),
                                            ignore_index=cfg.padding_idx, reduction="mean")
    assert torch.allclose(loss, ref, atol=1e-12)
    assert int(ntok) == int((cap[:, 1:] != cfg.padding_idx).sum())
