"""Generates tests/golden/train_small.npz from the CPU oracle of the widened rows (oracle/train_ref.py, fp64 / fp32):
feature adaptor forward + backward through the whole loss, two PaddlePaddle-1.8 Adam steps, and the learning-rate tables of
the five strategies of tools/util.py.  Like decoder_small_*.npz these pin the ORACLE (and through it the CUDA path) against
regressions; they are not outputs of PaddlePaddle ("parity unpinned").  Run from the repo root:
    python tests/golden/make_golden_train.py
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
from oracle import decoder_ref as R  # noqa: E402
from oracle import synth  # noqa: E402
from oracle import train_ref as TR  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
SMALL = dict(hidden_dim=64, embedding_size=32, vocab_size=101, num_positions=5, sentence_length=8, infer_max_length=9)
X = 24
LR_CASES = [(None, {}), ("cosine_decay", dict(decay_epoch=4)), ("cosine_decay_warmup", dict(warmup_epoch=3, max_epoch=10)),
            ("cosine_decay_restart", dict(decay_epoch=2)), ("cosine_decay_restart_warmup", dict(decay_epoch=2, warmup_epoch=3))]


def build():
    cfg = R.DecoderConfig(softmax_axis="position", **SMALL)
    f32 = lambda d: {k: v.float().double() for k, v in d.items()}          # stored fp32 inputs reproduce the outputs exactly
    params = f32(R.init_params(cfg, seed=11, dtype=torch.float64, bias_scale=0.1))
    ap = f32(TR.init_adaptor_params(cfg, X, seed=11, dtype=torch.float64, bias_scale=0.1))
    x = TR.make_encoder_map(cfg, 5, X, seed=5).double()
    cap = synth.make_captions(cfg, 5, seed=5)
    cap[1, 3:] = cfg.padding_idx
    loss, logits, grads, (P, g) = TR.loss_and_grads_from_map(params, ap, cfg, x, cap)
    out = {"loss": loss.numpy(), "logits": logits.numpy(), "enc_map": x.numpy(), "caption": cap.numpy(), "P": P.numpy(),
           "g": g.numpy()}
    for k, v in {**params, **ap}.items():
        out["param_" + k] = v.numpy()
    for k in list(ap) + ["enc_map", "lstm_w", "fc_v_w", "word_embedding"]:
        out["grad_" + k] = grads[k].numpy()
    # two Adam steps (float32, PaddlePaddle 1.8 rule) on fc_img_w with its gradient and 0.5x its gradient, clip 0.01
    p = ap["fc_img_w"].float().numpy().copy()
    m, v = np.zeros_like(p), np.zeros_like(p)
    gr = grads["fc_img_w"].float().numpy()
    b1p, b2p = TR.adam_step_numpy(p, gr, m, v, 1e-3, clip=0.01)
    TR.adam_step_numpy(p, 0.5 * gr, m, v, 1e-3, beta1_pow=b1p, beta2_pow=b2p, clip=0.01)
    out["adam_p2"], out["adam_m2"], out["adam_v2"] = p, m, v
    for name, kw in LR_CASES:
        graph = TR.LrGraph(name, 5e-5, 1000, 64, **kw)
        out["lr_" + str(name)] = np.array([graph.run() for _ in range(16 * 12)], dtype=np.float64)
    return out


if __name__ == "__main__":
    out = build()
    path = os.path.join(HERE, "train_small.npz")
    keep64 = ("loss",)
    np.savez_compressed(path, **{k: (v.astype(np.float32) if v.dtype == np.float64 and k not in keep64
                                     and not k.startswith("lr_") else v) for k, v in out.items()})
    print(path, os.path.getsize(path))
