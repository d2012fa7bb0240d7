"""Generates tests/golden/decoder_small_*.npz from the CPU oracle (oracle/decoder_ref.py, fp64).

The reference ships no golden vectors and PaddlePaddle 1.8 cannot be installed here, so
these fixtures pin the ORACLE (and, through it, the CUDA path) against regressions -- they are not outputs of the
reference itself ("parity unpinned").  Run from the repo root:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
from oracle import decoder_ref as R  # noqa: E402
from oracle import synth  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
SMALL = dict(hidden_dim=64, embedding_size=32, vocab_size=101, num_positions=5, sentence_length=8, infer_max_length=9)


def make(axis: str):
    cfg = R.DecoderConfig(softmax_axis=axis, **SMALL)
    # inputs are rounded to fp32 first so that the stored (fp32) inputs reproduce the stored outputs exactly
    params = {k: v.float().double() for k, v in R.init_params(cfg, seed=11, dtype=torch.float64, bias_scale=0.1).items()}
    g, P, cap = synth.make_batch(cfg, 6, seed=5, dtype=torch.float32)
    g, P = g.double(), P.double()
    cap[1, 3:] = cfg.padding_idx                       # a short caption (ragged)
    cap[2, 1] = cfg.padding_idx                        # a pad token fed as a source word (zero embedding row)
    loss, logits, alpha, grads = R.loss_and_grads(params, cfg, g, P, cap)
    greedy = R.greedy_decode(params, cfg, g, P)
    beam_ids, beam_scores = R.beam_search(params, cfg, g, P, 3)
    out = {"loss": loss.numpy(), "logits": logits.numpy(), "alpha": alpha.numpy(), "greedy": greedy.numpy(),
           "beam_ids": beam_ids.numpy(), "beam_scores": beam_scores.numpy(), "g": g.numpy(), "P": P.numpy(),
           "caption": cap.numpy()}
    for k, v in params.items():
        out["param_" + k] = v.numpy()
    for k, v in grads.items():
        out["grad_" + k] = v.numpy()
    path = os.path.join(HERE, "decoder_small_%s.npz" % axis)
    np.savez_compressed(path, **{k: (v.astype(np.float32) if v.dtype == np.float64 and k not in ("loss",) else v)
                                 for k, v in out.items()})
    print(path, os.path.getsize(path))


if __name__ == "__main__":
    make("feature")
    make("position")
