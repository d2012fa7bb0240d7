"""CPU ORACLE for the callers either side of the decoder (test infrastructure, NOT product code) -- PARITY UNPINNED.

Restates, for SURVEY.md 8f rows N1 / N2:
  * ImageCaptionModel._img2feature after the encoder   /root/reference/ImageCaptioning/model/model_adaAttention_aic.py:193-199
  * fluid.optimizer.Adam as train.py:26-31, 44-45 uses it (PaddlePaddle 1.8 adam_op) + GradientClipByValue (train.py:41-42)
  * the learning-rate schedules of tools/util.py:20-119, by SIMULATING the graph the reference builds (persistable
    counters that are incremented once per executor run), not by closed forms.
Only tests/, __graft_entry__.smoke() and bench.py's cpu legs may import this module.  PaddlePaddle 1.8 cannot be installed
here (see oracle/decoder_ref.py), so these restatements are pinned by a second literal NumPy version, closed-form
known answers and finite differences in tests/test_train_ref.py -- not by outputs of Paddle itself.
"""
from __future__ import annotations

import math
from typing import Dict

import numpy as np
import torch

from . import decoder_ref as R

ADAPTOR_NAMES = ["fc_img_w", "fc_img_b", "fc_glob_w", "fc_glob_b"]


def adaptor_shapes(cfg: R.DecoderConfig, encoder_channel: int = 1280):
    H = cfg.hidden_dim
    return {"fc_img_w": (encoder_channel, H), "fc_img_b": (H,),          # model:197
            "fc_glob_w": (encoder_channel, H), "fc_glob_b": (H,)}        # model:199


def init_adaptor_params(cfg: R.DecoderConfig, encoder_channel: int = 1280, seed: int = 0, dtype=torch.float64,
                        bias_scale: float = 0.1) -> Dict[str, torch.Tensor]:
    """Paddle fc defaults (Xavier-uniform, zero bias); non-zero biases by default so that bias bugs are visible."""
    gen = torch.Generator().manual_seed(seed + 104729)
    out = {}
    for name, shape in adaptor_shapes(cfg, encoder_channel).items():
        if name.endswith("_w"):
            t = (torch.rand(shape, generator=gen, dtype=torch.float64) * 2 - 1) * math.sqrt(6.0 / (shape[0] + shape[1]))
        else:
            t = (torch.rand(shape, generator=gen, dtype=torch.float64) * 2 - 1) * bias_scale
        out[name] = t.to(dtype)
    return out


def make_encoder_map(cfg, B, encoder_channel=1280, seed=20260, dtype=torch.float32):
    """Synthetic MobileNetV2 output (values in [0, 6]: the encoder ends in ReLU6); the generator lives in the product
    package so that bench.py's own arm never imports oracle/."""
    from myimagecaptioningmodel_b200.synth import make_encoder_map as _m
    return _m(cfg, B, encoder_channel, seed, dtype)


def img2feature(ap, enc_map: torch.Tensor):
    """model:194-199.  enc_map [B, X, k] (= the encoder output reshaped by model:195) -> (img_embed [B,k,H], img_global [B,H])."""
    img = enc_map.transpose(1, 2)                                        # model:196  [B, k, X]
    img_embed = torch.relu(img @ ap["fc_img_w"] + ap["fc_img_b"])        # model:197  num_flatten_dims=2, act='relu'
    img_global = img.mean(dim=1)                                         # model:198  reduce_mean(dim=1)
    img_global = torch.relu(img_global @ ap["fc_glob_w"] + ap["fc_glob_b"])   # model:199
    return img_embed, img_global


def img2feature_numpy(ap, enc_map):
    """Second, literal restatement (loops over images, explicit reshape/transpose) used to pin `img2feature`."""
    B, X, k = enc_map.shape
    H = ap["fc_img_w"].shape[1]
    emb = np.zeros((B, k, H), dtype=enc_map.dtype)
    glob = np.zeros((B, H), dtype=enc_map.dtype)
    for b in range(B):
        rows = np.ascontiguousarray(enc_map[b].reshape(X, k).T)          # [k, X]
        for j in range(k):
            emb[b, j] = np.maximum(rows[j] @ ap["fc_img_w"] + ap["fc_img_b"], 0.0)
        m = rows.sum(axis=0) / k
        glob[b] = np.maximum(m @ ap["fc_glob_w"] + ap["fc_glob_b"], 0.0)
    return emb, glob


def loss_and_grads_from_map(params, ap, cfg, enc_map, caption):
    """training_network after the encoder (model:161-183): loss, logits and the gradients of every decoder parameter,
    every adaptor parameter and of the encoder map itself ('enc_map')."""
    ps = {k: v.detach().clone().requires_grad_(True) for k, v in params.items()}
    aps = {k: v.detach().clone().requires_grad_(True) for k, v in ap.items()}
    x = enc_map.detach().clone().requires_grad_(True)
    P, g = img2feature(aps, x)
    loss, logits, _ = R.training_network(ps, cfg, g, P, caption)
    loss.backward()
    grads = {k: (v.grad if v.grad is not None else torch.zeros_like(v)) for k, v in {**ps, **aps}.items()}
    grads["enc_map"] = x.grad
    return loss.detach(), logits.detach(), grads, (P.detach(), g.detach())


# ----------------------------------------------------------------------------------------------------------------
# Adam, PaddlePaddle 1.8 (paddle/fluid/operators/optimizers/adam_op.h, AdamFunctor), float32 like the reference
# ----------------------------------------------------------------------------------------------------------------
def adam_step_numpy(p, g, m, v, lr, beta1=0.9, beta2=0.999, epsilon=1e-8, beta1_pow=None, beta2_pow=None, clip=0.0):
    """One update in place on float32 arrays; beta?_pow are the accumulators BEFORE the update (beta^t at step t).
    Returns the accumulators after it.  clip > 0: GradientClipByValue(max=clip, min=-clip) on the gradient."""
    f = np.float32
    beta1_pow = f(beta1) if beta1_pow is None else f(beta1_pow)
    beta2_pow = f(beta2) if beta2_pow is None else f(beta2_pow)
    g = g.astype(np.float32)
    if clip and clip > 0:
        g = np.clip(g, f(-clip), f(clip))
    lr_t = f(lr) * np.sqrt(f(1) - beta2_pow, dtype=np.float32) / (f(1) - beta1_pow)
    m[...] = f(beta1) * m + (f(1) - f(beta1)) * g
    v[...] = f(beta2) * v + (f(1) - f(beta2)) * g * g
    p[...] = p - lr_t * (m / (np.sqrt(v, dtype=np.float32) + f(epsilon)))
    return f(beta1_pow * f(beta1)), f(beta2_pow * f(beta2))


def adam_step_torch(params, grads, state, lr, beta1=0.9, beta2=0.999, epsilon=1e-8):
    """The same update on a dict of torch CPU tensors, in place, with torch's multi-threaded ops (the reference-arm timing
    leg of bench.py applies it after every backward pass, like train.py:44-45).  `state` carries moments and beta powers."""
    if not state:
        state["m"] = {k: torch.zeros_like(v) for k, v in params.items()}
        state["v"] = {k: torch.zeros_like(v) for k, v in params.items()}
        state["b1p"], state["b2p"] = np.float32(beta1), np.float32(beta2)
    lr_t = float(np.float32(lr) * np.sqrt(np.float32(1) - state["b2p"]) / (np.float32(1) - state["b1p"]))
    with torch.no_grad():
        for k, p in params.items():
            g, m, v = grads[k], state["m"][k], state["v"][k]
            m.mul_(beta1).add_(g, alpha=1 - beta1)
            v.mul_(beta2).addcmul_(g, g, value=1 - beta2)
            p.addcdiv_(m, v.sqrt().add_(epsilon), value=-lr_t)
    state["b1p"] = np.float32(state["b1p"] * np.float32(beta1))
    state["b2p"] = np.float32(state["b2p"] * np.float32(beta2))


# ----------------------------------------------------------------------------------------------------------------
# learning-rate schedules: a run-by-run simulation of the graphs tools/util.py builds
# ----------------------------------------------------------------------------------------------------------------
class LrGraph:
    """Each `run()` is one executor run of the training program: the persistable counters advance exactly as the ops of
    util.py would advance them, and the value of the `lr` variable of that run is returned.

    Paddle 1.8 `autoincreased_step_counter(begin, step=1)`: the counter variable is initialised to begin - 1 and an
    increment op is PREPENDED to the program, so the first run reads `begin` (util.py:48-52)."""

    def __init__(self, strategy, base_lr, sample_cnt, batch_size, decay_epoch=0, warmup_epoch=3, max_epoch=10):
        if strategy not in [None, "cosine_decay", "cosine_decay_restart", "cosine_decay_restart_warmup",
                            "cosine_decay_warmup"]:
            raise ValueError("Lr衰减策略错误")                                            # util.py:21-23
        self.strategy, self.base_lr = strategy, base_lr
        self.step_each_epoch = math.ceil(sample_cnt / batch_size)                         # util.py:24
        self.decay_epoch, self.warmup_epoch, self.max_epoch = decay_epoch, warmup_epoch, max_epoch
        begin = 1 if strategy in ("cosine_decay_restart_warmup", "cosine_decay_warmup") else 0
        self.counter = begin - 1           # '@LR_DECAY_COUNTER@'
        self.cur_epoch_var = 0.0           # persistable 'cur_epoch' of util.py:92-93

    def run(self) -> float:
        self.counter += 1                  # the prepended increment op
        gs = float(self.counter)
        spe, base = self.step_each_epoch, self.base_lr
        if self.strategy is None:
            return base                                                                   # util.py:44-45
        if self.strategy == "cosine_decay":                                               # fluid.layers.cosine_decay
            cur_epoch = math.floor(gs / spe)
            return base * 0.5 * (math.cos(cur_epoch * math.pi / self.decay_epoch) + 1)
        if self.strategy == "cosine_decay_warmup":                                        # util.py:55-68
            start_lr = 0.00001
            linear_factor = (base - start_lr) / self.warmup_epoch
            cur_epoch = float(math.floor(gs / spe))
            lr = 0.0
            if cur_epoch < self.warmup_epoch:
                lr = start_lr + linear_factor * cur_epoch
            else:
                lr = 0.5 * base * (math.cos((cur_epoch - self.warmup_epoch) * math.pi /
                                            float(self.max_epoch - self.warmup_epoch)) + 1)
            return lr
        t_mul, m_mul = 2.0, 1.0
        if self.strategy == "cosine_decay_restart":                                       # util.py:71-88
            cur_epoch = math.floor(gs / spe)
            completed_fraction = cur_epoch / self.decay_epoch
            i_restart = math.floor(math.log(1.0 - completed_fraction * (1.0 - t_mul)) / math.log(t_mul))
            sum_r = (1.0 - t_mul ** i_restart) / (1.0 - t_mul)
            completed_fraction = (completed_fraction - sum_r) / t_mul ** i_restart
            m_fac = m_mul ** int(i_restart)
            return base * (0.5 * m_fac * (math.cos(math.pi * completed_fraction) + 1))
        # cosine_decay_restart_warmup, util.py:91-119
        start_lr = 0.00001
        cur_part = self.counter % spe
        if cur_part <= 0:
            self.cur_epoch_var = self.cur_epoch_var + 1
        linear_step = base - start_lr
        if self.cur_epoch_var < self.warmup_epoch:
            return start_lr + linear_step * (self.cur_epoch_var / float(self.warmup_epoch))
        completed_fraction = (self.cur_epoch_var - self.warmup_epoch) / self.decay_epoch
        i_restart = math.floor(math.log(1.0 - completed_fraction * (1.0 - t_mul)) / math.log(t_mul))
        sum_r = (1.0 - t_mul ** i_restart) / (1.0 - t_mul)
        completed_fraction = (completed_fraction - sum_r) / t_mul ** i_restart
        return base * (0.5 * (math.cos(math.pi * completed_fraction) + 1))
