"""CPU: pins of oracle/train_ref.py (feature adaptor, PaddlePaddle-1.8 Adam, learning-rate graphs) and the host-side
schedule / store logic of the product that mirrors tools/util.py and train.py:26-31 (no GPU needed)."""
import math

import numpy as np
import pytest
import torch

import myimagecaptioningmodel_b200 as M
from oracle import decoder_ref as R
from oracle import synth
from oracle import train_ref as TR

DIMS = dict(hidden_dim=32, embedding_size=16, vocab_size=53, num_positions=7, sentence_length=7, infer_max_length=7)


def test_img2feature_two_restatements_agree():
    cfg = R.DecoderConfig(**DIMS)
    ap = TR.init_adaptor_params(cfg, 24, seed=1)
    x = TR.make_encoder_map(cfg, 3, 24, dtype=torch.float64)
    P, g = TR.img2feature(ap, x)
    Pn, gn = TR.img2feature_numpy({k: v.numpy() for k, v in ap.items()}, x.numpy())
    assert P.shape == (3, 7, 32) and g.shape == (3, 32)
    assert np.abs(P.numpy() - Pn).max() < 1e-12 and np.abs(g.numpy() - gn).max() < 1e-12
    assert float(P.min()) >= 0.0 and float(g.min()) >= 0.0                 # act='relu' (model:197, 199)
    # known answer: zero weights -> relu(bias) at every position
    z = {k: torch.zeros_like(v) for k, v in ap.items()}
    z["fc_img_b"] = torch.linspace(-1, 1, 32, dtype=torch.float64)
    P0, g0 = TR.img2feature(z, x)
    assert torch.equal(P0[1, 4], torch.relu(z["fc_img_b"])) and float(g0.abs().max()) == 0.0
    # the global feature sees the MEAN over the k positions of the transposed map (model:198)
    one = {k: torch.zeros_like(v) for k, v in ap.items()}
    one["fc_glob_w"][5, 9] = 1.0
    _, g1 = TR.img2feature(one, x)
    assert torch.allclose(g1[:, 9], x[:, 5, :].mean(dim=1))


def test_img2feature_gradients_finite_difference():
    cfg = R.DecoderConfig(softmax_axis="position", **DIMS)
    params = R.init_params(cfg, 2)
    ap = TR.init_adaptor_params(cfg, 16, seed=2)
    x = TR.make_encoder_map(cfg, 2, 16, dtype=torch.float64) + 0.05
    cap = synth.make_captions(cfg, 2, 3)
    loss, _, grads, _ = TR.loss_and_grads_from_map(params, ap, cfg, x, cap)

    def f(ap_, x_):
        P, g = TR.img2feature(ap_, x_)
        return float(R.training_network(params, cfg, g, P, cap)[0])

    eps = 1e-6
    for name, idx in (("fc_img_w", (3, 5)), ("fc_glob_w", (7, 11)), ("fc_img_b", (4,)), ("fc_glob_b", (2,))):
        hi = {k: v.clone() for k, v in ap.items()}
        lo = {k: v.clone() for k, v in ap.items()}
        hi[name][idx] += eps
        lo[name][idx] -= eps
        fd = (f(hi, x) - f(lo, x)) / (2 * eps)
        assert abs(fd - float(grads[name][idx])) < 1e-6 * max(1.0, abs(fd)), name
    xh, xl = x.clone(), x.clone()
    xh[1, 3, 2] += eps
    xl[1, 3, 2] -= eps
    fd = (f(ap, xh) - f(ap, xl)) / (2 * eps)
    assert abs(fd - float(grads["enc_map"][1, 3, 2])) < 1e-6 * max(1.0, abs(fd))


def test_adam_paddle18_known_answers():
    f = np.float32
    g = np.array([0.5, -2.0, 1e-3, 0.0], dtype=f)
    p = np.ones(4, dtype=f)
    m, v = np.zeros(4, dtype=f), np.zeros(4, dtype=f)
    b1p, b2p = TR.adam_step_numpy(p, g, m, v, lr=0.1)
    # step 1: m = 0.1 g, v = 0.001 g^2, lr_t = lr sqrt(1 - b2) / (1 - b1)  ->  p -= lr * g / (|g| + eps / sqrt(1 - b2))
    want = 1.0 - 0.1 * g / (np.abs(g) + 1e-8 / math.sqrt(1 - 0.999))
    assert np.allclose(p, want, rtol=2e-6) and p[3] == 1.0
    assert abs(b1p - 0.81) < 1e-7 and abs(b2p - 0.998001) < 1e-7
    # epsilon is NOT bias-corrected in Paddle 1.8: with v ~ 0 the update is lr_t * m / eps, not torch's form
    p2, m2, v2 = np.zeros(1, dtype=f), np.zeros(1, dtype=f), np.zeros(1, dtype=f)
    TR.adam_step_numpy(p2, np.array([1e-12], dtype=f), m2, v2, lr=1.0)
    lr_t = math.sqrt(1 - 0.999) / (1 - 0.9)
    assert np.allclose(p2, -lr_t * 1e-13 / (math.sqrt(1e-27) + 1e-8), rtol=1e-4)
    # GradientClipByValue
    p3, m3, v3 = np.zeros(2, dtype=f), np.zeros(2, dtype=f), np.zeros(2, dtype=f)
    TR.adam_step_numpy(p3, np.array([10.0, -0.01], dtype=f), m3, v3, lr=0.1, clip=0.5)
    assert np.allclose(m3, [0.05, -0.001])
    # many steps against torch.optim.Adam where the two rules coincide up to eps placement (|g| >> eps)
    torch.manual_seed(0)
    pt = torch.nn.Parameter(torch.randn(64))
    opt = torch.optim.Adam([pt], lr=1e-2, eps=1e-8)
    pn = pt.detach().numpy().astype(f).copy()
    mn, vn = np.zeros(64, dtype=f), np.zeros(64, dtype=f)
    b1p = b2p = None
    for _ in range(20):
        gr = torch.randn(64)
        pt.grad = gr.clone()
        opt.step()
        b1p, b2p = TR.adam_step_numpy(pn, gr.numpy(), mn, vn, 1e-2, beta1_pow=b1p, beta2_pow=b2p)
    assert np.abs(pn - pt.detach().numpy()).max() < 1e-5


def test_adam_torch_form_equals_the_numpy_form():
    torch.manual_seed(1)
    params = {"a": torch.randn(7, 5), "b": torch.randn(11)}
    pn = {k: v.numpy().copy() for k, v in params.items()}
    st = {k: (np.zeros_like(v), np.zeros_like(v)) for k, v in pn.items()}
    state, b1p, b2p = {}, None, None
    for _ in range(4):
        grads = {k: torch.randn_like(v) for k, v in params.items()}
        TR.adam_step_torch(params, grads, state, lr=1e-2)
        for k in pn:
            nb = TR.adam_step_numpy(pn[k], grads[k].numpy(), st[k][0], st[k][1], 1e-2, beta1_pow=b1p, beta2_pow=b2p)
        b1p, b2p = nb
    for k in pn:
        assert np.abs(params[k].numpy() - pn[k]).max() < 1e-6


@pytest.mark.parametrize("strategy,kw", [
    (None, {}),
    ("cosine_decay", dict(decay_epoch=4)),
    ("cosine_decay_warmup", dict(warmup_epoch=3, max_epoch=10)),
    ("cosine_decay_restart", dict(decay_epoch=2)),
    ("cosine_decay_restart_warmup", dict(decay_epoch=2, warmup_epoch=3)),
])
def test_lr_schedules_match_the_graph_simulation(strategy, kw):
    """product closed forms (optim.LRSchedule) == run-by-run simulation of the graphs tools/util.py builds"""
    sample_cnt, batch = 1000, 64                           # 16 steps per epoch
    graph = TR.LrGraph(strategy, 5e-5, sample_cnt, batch, **kw)
    sched = M.get_lr(strategy, 5e-5, sample_cnt, batch, **kw)
    assert sched.step_each_epoch == 16
    for n in range(16 * 14):
        want = graph.run()
        got = sched(n)
        assert abs(got - want) <= 1e-12 * max(1.0, abs(want)), (strategy, n, got, want)
    # random access (not only consecutive steps) gives the same values
    graph2 = TR.LrGraph(strategy, 5e-5, sample_cnt, batch, **kw)
    vals = [graph2.run() for _ in range(100)]
    sched2 = M.get_lr(strategy, 5e-5, sample_cnt, batch, **kw)
    for n in (57, 3, 99, 16, 15, 0):
        assert abs(sched2(n) - vals[n]) <= 1e-12


def test_lr_known_answers():
    with pytest.raises(ValueError):
        M.get_lr("linear", 1e-3, 10, 2)                     # util.py:21-23
    s = M.get_lr(None, 5e-5, 944996, 128)                   # config.py:27, 32-33, 37 defaults
    assert s(0) == s(10 ** 6) == 5e-5 and s.step_each_epoch == 7383
    w = M.get_lr("cosine_decay_warmup", 5e-5, 944996, 128, warmup_epoch=3, max_epoch=10)
    assert w(0) == 1e-5                                     # start_lr in epoch 0
    assert abs(w(7382) - (1e-5 + (5e-5 - 1e-5) / 3)) < 1e-18     # global_step 7383 = first step of epoch 1
    assert abs(w(3 * 7383 - 1) - 5e-5) < 1e-18              # cos(0): the full base lr right after the warm-up
    c = M.get_lr("cosine_decay", 1.0, 100, 10, decay_epoch=10)
    assert c(0) == 1.0 and abs(c(50) - 0.5) < 1e-12 and abs(c(100)) < 1e-12


def test_adaptor_store_layout_cpu():
    cfg = M.DecoderConfig(**DIMS)
    st = M.AdaptorStore(cfg, 24, "cpu").init_reference(1, 0.1)
    assert st.num_parameters == 2 * 24 * 32 + 2 * 32
    for n, t in st.params.items():
        assert (t.data_ptr() - st.flat.data_ptr()) % 256 == 0, n
    sd = st.state_dict()
    st2 = M.AdaptorStore(cfg, 24, "cpu").load(sd)
    assert torch.equal(st.flat, st2.flat)
    with pytest.raises(ValueError):
        M.AdaptorStore(cfg, 30, "cpu")
    with pytest.raises(KeyError):
        st2.load({})
    full = M.AdaptorStore(M.DecoderConfig(), 1280, "cpu")
    assert full.num_parameters == 2623488                   # the `_img2feature` parameter count


def test_golden_fixture_of_the_widened_rows(golden_dir):
    """tests/golden/train_small.npz (make_golden_train.py): the oracle reproduces it, the product's LR schedules match it"""
    import importlib.util
    import os
    z = np.load(os.path.join(golden_dir, "train_small.npz"))
    spec = importlib.util.spec_from_file_location("make_golden_train", os.path.join(golden_dir, "make_golden_train.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    cfg = R.DecoderConfig(softmax_axis="position", **mod.SMALL)
    params = {k[6:]: torch.from_numpy(z[k]).double() for k in z.files if k.startswith("param_") and k[6:] in R.param_shapes(cfg)}
    ap = {k: torch.from_numpy(z["param_" + k]).double() for k in TR.ADAPTOR_NAMES}
    x, cap = torch.from_numpy(z["enc_map"]).double(), torch.from_numpy(z["caption"])
    loss, logits, grads, (P, g) = TR.loss_and_grads_from_map(params, ap, cfg, x, cap)
    assert abs(float(loss) - float(z["loss"])) < 1e-12
    assert np.abs(logits.numpy() - z["logits"]).max() < 1e-5 and np.abs(P.numpy() - z["P"]).max() < 1e-6
    for k in TR.ADAPTOR_NAMES + ["enc_map", "lstm_w"]:
        assert np.abs(grads[k].numpy() - z["grad_" + k]).max() < 1e-7 * max(1.0, float(np.abs(z["grad_" + k]).max())) + 1e-9, k
    p = ap["fc_img_w"].float().numpy().copy()
    m, v = np.zeros_like(p), np.zeros_like(p)
    gr = torch.from_numpy(z["grad_fc_img_w"]).numpy()
    b1p, b2p = TR.adam_step_numpy(p, gr, m, v, 1e-3, clip=0.01)
    TR.adam_step_numpy(p, 0.5 * gr, m, v, 1e-3, beta1_pow=b1p, beta2_pow=b2p, clip=0.01)
    assert np.abs(p - z["adam_p2"]).max() < 1e-7 and np.abs(v - z["adam_v2"]).max() < 1e-12
    for name, kw in mod.LR_CASES:
        sched = M.get_lr(name, 5e-5, 1000, 64, **kw)
        table = z["lr_" + str(name)]
        assert max(abs(sched(n) - table[n]) for n in range(len(table))) < 1e-15, name
