"""Data-parallel training and image-sharded inference over one process per GPU.

Training shards along the batch; the only exchange steps are
  (1) an all-reduce of sum(mask) -- known from the captions alone, so it runs beside the forward pass -- so that every
      rank normalises by the GLOBAL token count: this reproduces the single-device reference at the global batch
      (model:169, 182), not ParallelExecutor's per-device averaging (offered as `pe_semantics=True`); the loss value
      sum(ce*mask) is all-reduced beside the backward pass;
  (2) the all-reduce (sum) of the flat gradient buffer over NCCL / NVLink, one call per gradient range as the backward
      pass finalises it; with `grad_exchange="bf16"` every range is rounded to bf16 for the wire (43 MB instead of 86 MB,
      libcapdec's cast kernels) and widened back into the fp32 buffer.
Inference shards by image with no collective on the data path.

The class is written against a small engine protocol (train_forward / train_backward / store.flat_grad) so the
host logic is testable on CPU with the gloo backend (tests/test_parallel_cpu.py drives it with the oracle).
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.distributed as dist


def shard_range(n_global: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous slice [lo, hi) of `n_global` items owned by `rank`; earlier ranks take the remainder."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world {}/{}".format(rank, world))
    base, rem = divmod(n_global, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class DataParallelTrainer:
    def __init__(self, engine, group: Optional["dist.ProcessGroup"] = None, pe_semantics: bool = False,
                 grad_exchange: Optional[str] = None):
        """grad_exchange: "fp32" (default) = all-reduce the fp32 gradient ranges as they are; "bf16" = round every range to
        bf16 for the all-reduce (half the NVLink bytes; the sum of N bf16 addends carries ~2^-9 relative rounding per
        element, inside the bf16 storage mode's 2e-2 tolerance) and widen it back."""
        self.engine = engine
        self.group = group
        self.pe_semantics = pe_semantics
        self.world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        if grad_exchange is None:
            grad_exchange = "fp32"
        if grad_exchange not in ("fp32", "bf16"):
            raise ValueError("grad_exchange must be 'fp32' or 'bf16'")
        self.grad_exchange = grad_exchange
        self._wire = {}

    def forward_backward(self, g, P, caption, want_input_grads: bool = False, optimizer=None, lengths=None):
        """One data-parallel forward+backward on this rank's shard.  Returns the GLOBAL loss as a device tensor
        (no host sync).  After the call engine.store.flat_grad holds the all-reduced gradient.

        optimizer (optim.Adam over [(store.flat, store.flat_grad), ...]): the parameter update is applied too, gradient
        group by gradient group on the side stream, each as soon as the group is final (and all-reduced): the update of the
        head parameters runs beside the BPTT kernel, only the last small group's update is exposed."""
        eng = self.engine
        kw = {} if lengths is None else {"lengths": lengths}     # host token counts: the head runs on the live rows only
        return self._run(caption, lambda: eng.train_forward(g, P, caption, **kw)[0],
                         lambda scale, inv: eng.train_backward(scale, inv, want_dg=want_input_grads,
                                                               want_dP=want_input_grads), extra=(), optimizer=optimizer)

    def forward_backward_map(self, enc_map, caption, want_dmap: bool = False, optimizer=None, lengths=None):
        """The same from the encoder map (feature adaptor in front): additionally all-reduces
        engine.adaptor.flat_grad.  Returns (global loss, d enc_map of this rank's shard or None); the map gradient of a
        shard is already normalised by the global token count."""
        eng = self.engine
        out = {}

        def bwd(scale, inv):
            out["dmap"] = eng.train_backward_map(scale, inv, want_dmap=want_dmap)

        kw = {} if lengths is None else {"lengths": lengths}
        loss = self._run(caption, lambda: eng.train_forward_map(enc_map, caption, **kw)[0], bwd,
                         extra=(eng.adaptor.flat_grad,), optimizer=optimizer)
        return loss, out.get("dmap")

    def _run(self, caption, fwd, bwd, extra, optimizer=None):
        eng = self.engine
        if optimizer is not None:
            want = 1 + len(extra)
            if len(optimizer.buffers) != want or optimizer.buffers[0][1].data_ptr() != eng.store.flat_grad.data_ptr():
                raise ValueError("the optimizer must be built over [(store.flat, store.flat_grad)%s]"
                                 % (", (adaptor.flat, adaptor.flat_grad)" if extra else ""))
        if self.world == 1:
            loss_buf = fwd()
            bwd(1.0, None)
            if optimizer is not None:
                self._grouped_exchange(extra, optimizer, reduce=False)
            return loss_buf[2].clone()      # not a view of the engine's live buffer: the next forward pass overwrites it
        if self.pe_semantics:
            # ParallelExecutor default (CoeffNumDevice): mean over devices of per-device token means
            loss_buf = fwd()
            bwd(1.0 / self.world, None)
            loss = loss_buf[2].clone()
            dist.all_reduce(loss, group=self.group)
            loss /= self.world
            self._grouped_exchange(extra, optimizer, reduce=True)
            return loss
        self._ensure_overlap()
        if self._overlap:
            return self._forward_backward_overlapped(caption, fwd, bwd, extra, optimizer)
        loss_buf = fwd()
        tot = loss_buf[:2].clone()
        dist.all_reduce(tot, group=self.group)               # (sum ce*mask, sum mask) over all ranks
        inv = (1.0 / tot[1]).reshape(1)
        bwd(1.0, inv)
        loss = tot[0] / tot[1]
        self._grouped_exchange(extra, optimizer, reduce=True)
        return loss

    def _ensure_overlap(self):
        if not getattr(self, "_overlap_ready", False):
            self._setup_overlap()
            self._overlap_ready = True

    def _forward_backward_overlapped(self, caption, fwd, bwd, extra, optimizer=None):
        """Global token normalisation without a synchronisation point between forward and backward.  The normaliser
        sum(mask) depends on the captions only (model:165-169), not on the forward pass: its all-reduce is issued on the
        communication stream BEFORE the forward pass and runs beside it, so the backward pass starts with the global
        1 / sum(mask) already on the device.  Only the loss VALUE needs the forward pass; its all-reduce runs beside the
        backward pass.  (Rescaling every gradient group by n_r / n_global after a backward pass on the local normaliser
        was measured first: the extra passes over the 86 MB gradient cost more than the exchange they hide --
        3.07 vs 3.01 ms per step at 2 GPUs.)"""
        eng = self.engine
        main = torch.cuda.current_stream()
        comm = self._comm
        pad = eng.cfg.padding_idx
        self._step_start.record(main)
        with torch.cuda.stream(comm):
            comm.wait_event(self._step_start)
            cnt = (caption[:, 1:] != pad).sum().to(torch.float32).reshape(1)     # model:165-169: mask = target != <pad>
            dist.all_reduce(cnt, group=self.group)
            inv = 1.0 / cnt
            self._count_done.record(comm)
        loss_buf = fwd()
        self._fwd_done.record(main)
        main.wait_event(self._count_done)
        bwd(1.0, inv)
        with torch.cuda.stream(comm):
            comm.wait_event(self._fwd_done)
            tot = loss_buf[:1].clone()
            dist.all_reduce(tot, group=self.group)
            loss = (tot * inv).reshape(())
        self._grouped_exchange(extra, optimizer, reduce=True)
        # tensors allocated on the communication stream and read on the main stream: keep them alive until the next step
        self._keep = (cnt, inv, tot, loss)
        return loss

    def _setup_overlap(self):
        """Bucketed gradient exchange overlapped with the backward pass: engines that can signal when a group of
        gradients is final (CaptionDecoderEngine.set_grad_ready_events) get one event per gradient range and a communication stream."""
        eng = self.engine
        self._overlap = False
        if not hasattr(eng, "set_grad_ready_events") or not torch.cuda.is_available():
            return
        self._comm = torch.cuda.Stream()
        self._upd = torch.cuda.Stream()      # parameter updates: the next range's all-reduce does not wait for Adam
        self._events = [torch.cuda.Event() for _ in range(getattr(eng, "N_GRAD_EVENTS", 6))]
        self._reduced = [torch.cuda.Event() for _ in self._events]
        for e in self._events:
            e.record()                      # creates the underlying cudaEvent_t
        eng.set_grad_ready_events(self._events)
        self._slices = eng.grad_group_slices()
        self._fwd_done = torch.cuda.Event()
        self._step_start = torch.cuda.Event()
        self._count_done = torch.cuda.Event()
        self._overlap = True

    def _grouped_exchange(self, extra=(), optimizer=None, reduce=True):
        """Per gradient group, in the order capdec_train_bwd finalises them (its events): all-reduce the group (reduce) and
        apply the optimizer to exactly that range (optimizer), on the side stream, beside the rest of the backward pass.
        `extra`: further flat gradient buffers that are final when the main stream reaches this point (the adaptor's);
        they are optimizer.buffers[1:]."""
        eng = self.engine
        self._ensure_overlap()
        flat = eng.store.flat_grad
        if not self._overlap:                                # engines without gradient-ready events (CPU tests)
            if reduce:
                self._all_reduce(flat, 0, flat.numel())
                for t in extra:
                    self._all_reduce(t, 0, t.numel())
            if optimizer is not None:
                optimizer.step()
            return
        main = torch.cuda.current_stream()
        lr = optimizer.current_lr() if optimizer is not None else None
        upd = self._upd if (reduce and optimizer is not None) else self._comm
        with torch.cuda.stream(self._comm):
            for ev, lo, hi in self._slices:
                # every group waits for its own event, recorded by capdec_train_bwd in the middle of the backward pass
                self._comm.wait_event(self._events[ev])
                if reduce:
                    self._all_reduce(flat, lo, hi)
                if optimizer is not None:
                    if upd is not self._comm:
                        self._reduced[ev].record(self._comm)
                        with torch.cuda.stream(upd):
                            upd.wait_event(self._reduced[ev])
                            optimizer.step_range(0, lo, hi, lr, stream=upd)
                    else:
                        optimizer.step_range(0, lo, hi, lr)
            if upd is not self._comm:
                self._comm.wait_stream(upd)
            if extra:
                self._comm.wait_stream(main)                 # the adaptor backward runs after capdec_train_bwd
                for i, t in enumerate(extra):
                    if reduce:
                        self._all_reduce(t, 0, t.numel())
                    if optimizer is not None:
                        optimizer.step_range(1 + i, 0, t.numel(), lr)
        main.wait_stream(self._comm)
        if optimizer is not None:
            optimizer.advance()


    def _all_reduce(self, flat, lo, hi):
        """sum over ranks of flat[lo:hi] (in place), on the current stream"""
        if hi <= lo:
            return
        if self.grad_exchange == "fp32" or (hi - lo) % 8 or not flat.is_cuda:
            dist.all_reduce(flat[lo:hi], group=self.group)
            return
        from ._capi import check, ptr, stream_ptr
        wire = self._wire.get(flat.data_ptr())
        if wire is None or wire.numel() < flat.numel():
            wire = torch.empty(flat.numel(), dtype=torch.bfloat16, device=flat.device)
            self._wire[flat.data_ptr()] = wire
        lib = self.engine.lib
        check(lib.capdec_cast_bf16(1, ptr(flat[lo:hi]), ptr(wire[lo:hi]), hi - lo, stream_ptr()), "capdec_cast_bf16")
        dist.all_reduce(wire[lo:hi], group=self.group)
        check(lib.capdec_cast_bf16(0, ptr(wire[lo:hi]), ptr(flat[lo:hi]), hi - lo, stream_ptr()), "capdec_cast_bf16")


def sharded_decode(engine, g, P, beam: int, finish_on_stop: bool, rank: int, world: int):
    """Image-sharded beam search: this rank decodes images [lo, hi) of the global batch; no collective."""
    lo, hi = shard_range(g.shape[0], rank, world)
    return (lo, hi) + tuple(engine.decode(g[lo:hi].contiguous(), P[lo:hi].contiguous(), beam, finish_on_stop))
