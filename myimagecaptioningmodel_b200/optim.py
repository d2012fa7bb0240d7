"""The step after the hot path: the reference's optimizer and learning-rate schedules.

  * ``Adam``        fluid.optimizer.Adam(lr) + opt.minimize(loss) (train.py:26-31, 44-45), PaddlePaddle 1.8 update rule,
                    optional GradientClipByValue (train.py:41-42, config.py:36).  One capdec_adam_step launch per flat
                    buffer (the decoder's ParameterStore, the adaptor's AdaptorStore): 16 B read + 12 B written per
                    parameter, nothing else -- HBM-bound.
  * ``get_lr``      tools/util.py:20-45 -- the five ``lr_decay_strategy`` values of config.py:33, as host functions of the
                    global step (the reference builds them as graph ops on a step counter; the value they produce at
                    step t is what is computed here).
The arithmetic of the update is in libcapdec.so (csrc/adaptor.cu); this module is host logic only.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _capi
from ._capi import check, ptr, stream_ptr

LR_STRATEGIES = (None, "cosine_decay", "cosine_decay_restart", "cosine_decay_restart_warmup", "cosine_decay_warmup")


# ----------------------------------------------------------------------------------------------------------------
# learning-rate schedules (tools/util.py)
# ----------------------------------------------------------------------------------------------------------------
def _restart_fraction(completed_fraction: float, t_mul: float) -> Tuple[float, float]:
    """util.py:76-82 / 106-112: (i_restart, fraction inside the current restart period)."""
    if t_mul == 1.0:
        i_restart = math.floor(completed_fraction)
        return i_restart, completed_fraction - i_restart
    i_restart = math.floor(math.log(1.0 - completed_fraction * (1.0 - t_mul)) / math.log(t_mul))
    sum_r = (1.0 - t_mul ** i_restart) / (1.0 - t_mul)
    return i_restart, (completed_fraction - sum_r) / t_mul ** i_restart


class LRSchedule:
    """Learning rate as a function of the number of optimizer steps already taken.

    ``lr(n)`` is the value the reference graph feeds to Adam in its (n+1)-th executor run.  The step counters of the
    reference start at ``begin`` (util.py:48-52: `_decay_step_counter(begin)`; the counter is incremented before it is
    read, Paddle's autoincreased_step_counter), which is 0 for cosine_decay / cosine_decay_restart and 1 for the two
    warm-up variants."""

    def __init__(self, strategy, base_lr: float, sample_cnt: int, batch_size: int, decay_epoch: int = 0,
                 warmup_epoch: int = 3, max_epoch: int = 10, start_lr: float = 0.00001, m_mul: float = 1.0,
                 t_mul: float = 2.0):
        if strategy not in LR_STRATEGIES:
            raise ValueError("Lr衰减策略错误")                                       # util.py:21-23
        self.strategy, self.base_lr = strategy, float(base_lr)
        self.step_each_epoch = math.ceil(sample_cnt / batch_size)                    # util.py:24
        self.decay_epoch, self.warmup_epoch, self.max_epoch = decay_epoch, warmup_epoch, max_epoch
        self.start_lr, self.m_mul, self.t_mul = start_lr, m_mul, t_mul
        self._epoch_counter = 0.0          # the persistable `cur_epoch` variable of util.py:92-93
        self._last_n = -1

    @property
    def counter_begin(self):
        """`begin` of the graph's @LR_DECAY_COUNTER@ variable (util.py:48-52): the variable starts at begin - 1 and is
        incremented before it is read, so after n executor runs it holds n - 1 + begin.  None: the graph has no counter
        (strategy None, util.py:44-45)."""
        if self.strategy is None:
            return None
        return 1 if self.strategy in ("cosine_decay_warmup", "cosine_decay_restart_warmup") else 0

    def cur_epoch_state(self, n: int) -> float:
        """value of the persistable `cur_epoch` variable (util.py:92-99) after n executor runs"""
        return float(sum(1 for t in range(1, n + 1) if t % self.step_each_epoch == 0))

    def __call__(self, n: int) -> float:
        s = self.strategy
        if s is None:
            return self.base_lr                                                      # util.py:44-45
        if s == "cosine_decay":
            # fluid.layers.cosine_decay (Paddle 1.8): global_step counts from 0
            cur_epoch = math.floor(n / self.step_each_epoch)
            return self.base_lr * 0.5 * (math.cos(cur_epoch * math.pi / self.decay_epoch) + 1)
        if s == "cosine_decay_warmup":                                               # util.py:55-68
            global_step = float(n + 1)
            cur_epoch = math.floor(global_step / self.step_each_epoch)
            if cur_epoch < self.warmup_epoch:
                return self.start_lr + (self.base_lr - self.start_lr) / self.warmup_epoch * cur_epoch
            return 0.5 * self.base_lr * (math.cos((cur_epoch - self.warmup_epoch) * math.pi /
                                                  float(self.max_epoch - self.warmup_epoch)) + 1)
        if s == "cosine_decay_restart":                                              # util.py:71-88
            cur_epoch = math.floor(float(n) / self.step_each_epoch)
            i_restart, frac = _restart_fraction(cur_epoch / self.decay_epoch, self.t_mul)
            m_fac = self.m_mul ** int(i_restart)
            return self.base_lr * (0.5 * m_fac * (math.cos(math.pi * frac) + 1))
        # cosine_decay_restart_warmup (util.py:91-119): `cur_epoch` is a persistable counter that is incremented whenever
        # global_step % step_each_epoch == 0, so the schedule has state and must be evaluated for consecutive n
        if n != self._last_n + 1 and n != self._last_n:
            self._epoch_counter = float(sum(1 for t in range(1, n + 1) if t % self.step_each_epoch == 0))
        if n != self._last_n:
            if (n + 1) % self.step_each_epoch <= 0:
                self._epoch_counter += 1.0
            self._last_n = n
        cur_epoch = self._epoch_counter
        if cur_epoch < self.warmup_epoch:
            return self.start_lr + (self.base_lr - self.start_lr) * (cur_epoch / float(self.warmup_epoch))
        _, frac = _restart_fraction((cur_epoch - self.warmup_epoch) / self.decay_epoch, self.t_mul)
        return self.base_lr * (0.5 * (math.cos(math.pi * frac) + 1))


def get_lr(strategy, base_lr, sample_cnt, batch_size, **train_cfg) -> LRSchedule:
    """tools/util.py:20-45 (`train_cfg`: decay_epoch, warmup_epoch, max_epoch of config.py:34-35, 43)."""
    return LRSchedule(strategy, base_lr, sample_cnt, batch_size, **train_cfg)


# ----------------------------------------------------------------------------------------------------------------
# Adam (PaddlePaddle 1.8 adam_op)
# ----------------------------------------------------------------------------------------------------------------
class Adam:
    """fluid.optimizer.Adam(learning_rate, beta1=0.9, beta2=0.999, epsilon=1e-8) over flat fp32 buffers.

    ``buffers`` is a sequence of (flat_param, flat_grad) CUDA tensor pairs -- e.g. ``[(store.flat, store.flat_grad)]``.
    The beta-power accumulators are float32 like Paddle's Beta1Pow / Beta2Pow variables (initialised to beta, multiplied
    by beta after every update), so long runs accumulate the same rounding."""

    def __init__(self, buffers: Sequence[Tuple[torch.Tensor, torch.Tensor]], learning_rate=0.00005, beta1: float = 0.9,
                 beta2: float = 0.999, epsilon: float = 1e-8, gradient_clip: float = 0.0):
        self.lib = _capi.load_library()
        self.buffers: List[Tuple[torch.Tensor, torch.Tensor]] = []
        for p, g in buffers:
            if p.dtype != torch.float32 or g.dtype != torch.float32 or p.numel() != g.numel() or not p.is_cuda \
                    or not g.is_cuda or not p.is_contiguous() or not g.is_contiguous():
                raise ValueError("Adam needs matching contiguous float32 CUDA (param, grad) buffers")
            self.buffers.append((p, g))
        self.moment1 = [torch.zeros_like(p) for p, _ in self.buffers]
        self.moment2 = [torch.zeros_like(p) for p, _ in self.buffers]
        self.lr: Callable[[int], float] = learning_rate if callable(learning_rate) else (lambda n, v=float(learning_rate): v)
        self.beta1, self.beta2, self.epsilon = float(beta1), float(beta2), float(epsilon)
        self.gradient_clip = float(gradient_clip or 0.0)           # config.py:36: False or the clip value
        self.beta1_pow = np.float32(beta1)
        self.beta2_pow = np.float32(beta2)
        self.steps = 0

    def current_lr(self) -> float:
        return float(self.lr(self.steps))

    def step_range(self, buf: int, lo: int, hi: int, lr: Optional[float] = None, stream=None) -> None:
        """Update elements [lo, hi) of flat buffer `buf` with the step's learning rate and bias corrections.  The update is
        element-wise, so a step may be applied range by range (e.g. per gradient group, as soon as the group's all-reduce
        has finished -- DataParallelTrainer); `advance()` closes the step."""
        lr = self.current_lr() if lr is None else lr
        (p, g), m, v = self.buffers[buf], self.moment1[buf], self.moment2[buf]
        if hi <= lo:
            return
        check(self.lib.capdec_adam_step(ptr(p[lo:hi]), ptr(g[lo:hi]), ptr(m[lo:hi]), ptr(v[lo:hi]), hi - lo, lr, self.beta1,
                                        self.beta2, self.epsilon, float(self.beta1_pow), float(self.beta2_pow),
                                        self.gradient_clip, stream_ptr(stream)), "capdec_adam_step")

    def advance(self) -> None:
        """Paddle's Beta1Pow / Beta2Pow update and the step counter, once per optimizer step."""
        self.beta1_pow = np.float32(self.beta1_pow * np.float32(self.beta1))
        self.beta2_pow = np.float32(self.beta2_pow * np.float32(self.beta2))
        self.steps += 1
        for p, _ in self.buffers:           # the handle's operand copies of these parameters are stale now
            _capi.bump_version(p)

    def step(self, stream=None) -> float:
        """One update of every buffer from its gradient; returns the learning rate that was used."""
        lr = self.current_lr()
        for i, (p, _) in enumerate(self.buffers):
            self.step_range(i, 0, p.numel(), lr, stream)
        self.advance()
        return lr

    def state_dict(self):
        return {"moment1": [m.detach().cpu().clone() for m in self.moment1],
                "moment2": [v.detach().cpu().clone() for v in self.moment2],
                "beta1_pow": float(self.beta1_pow), "beta2_pow": float(self.beta2_pow), "steps": self.steps}

    def load_state_dict(self, sd):
        for dst, src in zip(self.moment1, sd["moment1"]):
            dst.copy_(src)
        for dst, src in zip(self.moment2, sd["moment2"]):
            dst.copy_(src)
        self.beta1_pow, self.beta2_pow = np.float32(sd["beta1_pow"]), np.float32(sd["beta2_pow"])
        self.steps = int(sd["steps"])


def get_optimizer(engine, train_cfg: Optional[dict] = None, sample_count: int = 944996):
    """train.py:26-31 for an engine: Adam over the decoder (and, when attached, adaptor) parameters with the learning
    rate of config.train (defaults: config.py:30-46)."""
    cfg = {"learning_rate": 0.00005, "lr_decay_strategy": None, "decay_epoch": 0, "warmup_epoch": 3, "gradient_clip": False,
           "batch_size": 128, "max_epoch": 10}
    cfg.update(train_cfg or {})
    lr = get_lr(cfg["lr_decay_strategy"], cfg["learning_rate"], sample_count, cfg["batch_size"],
                decay_epoch=cfg["decay_epoch"], warmup_epoch=cfg["warmup_epoch"], max_epoch=cfg["max_epoch"])
    bufs = [(engine.store.flat, engine.store.flat_grad)]
    if getattr(engine, "adaptor", None) is not None:
        bufs.append((engine.adaptor.flat, engine.adaptor.flat_grad))
    return Adam(bufs, lr, gradient_clip=cfg["gradient_clip"] or 0.0), lr
