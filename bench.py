#!/usr/bin/env python
"""bench.py -- decoder train tokens/s (headline) and beam-search images/s on N x B200, next to the CPU oracle.

Contract (one JSON line on stdout from rank 0):
  python bench.py --gpus N --steps K --warmup W            # our arm   (N>1: launched under torch.distributed.run)
  python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (the oracle port), rank 0 only

A "step" of our arm = one teacher-forced decoder train step on one batch of synthetic features and captions at
BASELINE.json configs[1]: B = 256 per GPU, T = 34, H = 1024, E = 256, V = 12295, k = 49, bf16 storage with fp32
accumulation: capdec_bind_params (re-layout of the fp32 weights after the previous update) + capdec_train_fwd
+ capdec_train_bwd (+ the data-parallel exchanges at N > 1) + capdec_adam_step over the flat decoder parameters
(PaddlePaddle-1.8 rule, lr 5e-5: train.py:26-31, 44-45).  Inputs are resident in HBM for `value`;
`e2e` goes through the same public calls with pinned HOST buffers (H2D of features and captions and a D2H read
of the loss inside the timed region).  The reference arm does the same work on the host cores: the oracle's forward +
autograd backward + Adam.  `extras.full_step_from_encoder_map` is the widened step: the feature adaptor
in front of the decoder, Adam over decoder + adaptor parameters behind it.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
    os.environ["NCCL_DEBUG"] = "WARN"                   # keep NCCL's version banner off stdout: one JSON line only

import torch  # noqa: E402

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "decoder_train_tokens_per_sec"
UNIT = "tokens/s"
B_PER_GPU = 256
BEAM_IMAGES_PER_GPU = 512
BEAM = 3
SOFTMAX_AXIS = "position"     # the full attention arithmetic; the parity gate also covers 'feature'


GRAD_EXCHANGE = {
    "fp32": "fp32, one all-reduce per gradient range as the backward pass finalises it",
    "bf16": "bf16 on the wire (each fp32 gradient range rounded to nearest even, NCCL sum, widened back into the fp32 buffer "
            "Adam reads), one all-reduce per gradient range as the backward pass finalises it; --grad-exchange fp32 for the "
            "exact exchange",
}


def workload_config(n_gpus, b_per_gpu=B_PER_GPU, grad_exchange="bf16"):
    return {"workload": "teacher_forced_train_step_B%d_per_gpu_T34_H1024_E256_V12295_k49" % b_per_gpu,
            "global_batch": b_per_gpu * n_gpus,
            "batch_per_gpu": b_per_gpu, "steps_per_caption": 34, "softmax_axis": SOFTMAX_AXIS,
            "lengths": "host token counts passed (captions L ~ U{5..33} words: 57 % of the B x 34 rows carry a target; the "
                       "head skips the others, loss and gradients unchanged); tokens/s counts all B x 34 positions",
            "parallelism": "dp%d" % n_gpus, "grad_exchange": GRAD_EXCHANGE[grad_exchange],
            "step": "bind_params+train_fwd+train_bwd(+allreduce)+adam(paddle-1.8 rule, lr 5e-5)",
            "e2e": "pinned host buffers, H2D of the next batch prefetched on a copy stream, loss D2H every step (read by the "
                   "host one step later)",
            "l2": "per-step working set ~1.5 GB >> 126 MB L2 (no flush needed)"}


# ------------------------------------------------------------------------------------------------------------------
# clocks sampling (nvidia-smi) during the timed region
# ------------------------------------------------------------------------------------------------------------------
class ClockSampler:
    """One long-running `nvidia-smi -lms 20` process (a timed region of K train steps lasts tens of milliseconds: spawning
    nvidia-smi per sample would give one sample at best).  Every line is time-stamped on arrival; summary() keeps the
    samples that fall inside the marked windows (the device-timed regions)."""
    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index, enabled=True):
        self.index = index
        self.enabled = enabled
        self.samples = []      # (t, fields)
        self.windows = []      # [t0, t1]
        self.proc = None
        self.thread = None

    def _run(self):
        try:
            for line in self.proc.stdout:
                parts = [p.strip() for p in line.strip().split(",")]
                if len(parts) >= 6:
                    self.samples.append((time.perf_counter(), parts))
        except Exception:
            pass

    def __enter__(self):
        if not self.enabled:
            return self
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.FIELDS,
                                          "--format=csv,noheader,nounits", "-lms", "20"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True, bufsize=1)
            self.thread = threading.Thread(target=self._run, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None
        return self

    def window(self):
        """context manager marking a device-timed region"""
        sampler = self

        class _W:
            def __enter__(self_w):
                self_w.t0 = time.perf_counter()

            def __exit__(self_w, *a):
                sampler.windows.append((self_w.t0, time.perf_counter()))
        return _W()

    def __exit__(self, *a):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()
        if self.thread is not None:
            self.thread.join(timeout=5)

    def summary(self):
        inside = [f for (t, f) in self.samples if any(t0 <= t <= t1 + 0.02 for t0, t1 in self.windows)]
        if not inside:
            inside = [f for (_, f) in self.samples]      # regions shorter than the sampling period: all samples of the run
        if not inside:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = [float(s[0]) for s in inside if s[0].replace(".", "").isdigit()]
        mx = [float(s[1]) for s in inside if s[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(s[2 + i].lower().startswith("active") for s in inside)]
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(inside), "period_ms": 20}


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return {"hbm_gbs": p["hbm_gbs"], "bf16_burst": p["bf16_tflops"], "bf16_sustained": p["bf16_tflops_sustained"],
                "source": "measured (MEASURED_PEAKS.json)"}
    return {"hbm_gbs": 6650.0, "bf16_burst": 1590.0, "bf16_sustained": 1400.0, "source": "fallback (B200_PROFILING.md)"}


def load_ncu_traffic_table():
    """{phase: dram__bytes_read.sum + dram__bytes_write.sum per launch} of the kernels in the roofline table, from the
    committed ncu --set full capture (profiles/r02_kernel_traffic.json, written by scripts/ncu_traffic.py); {} if none."""
    path = os.path.join(ROOT, "profiles", "r02_kernel_traffic.json")
    try:
        return {k: float(v) for k, v in json.load(open(path)).items()}
    except Exception:
        return {}


# ------------------------------------------------------------------------------------------------------------------
# CPU reference arm: the oracle port of the reference's Paddle graph, timed on the host cores
# ------------------------------------------------------------------------------------------------------------------
def cpu_reference_step(params, cfg, g, P, cap, adam_state, chunk=64):
    """forward + autograd backward + the Adam update of every parameter (train.py:44-45), all on the host cores.
    The batch goes through the oracle in chunks of `chunk` samples (samples are independent; the loss is
    sum(ce * mask) / sum(mask) over the WHOLE batch, so every chunk back-propagates its loss sum over the global token
    count and the gradients add up): the same arithmetic as one pass, with the autograd state of 64 instead of 256
    samples alive at a time (2.6 GB instead of ~10 GB)."""
    from oracle import decoder_ref as R
    from oracle import train_ref as TR
    B = g.shape[0]
    if B <= chunk:
        _, _, _, grads = R.loss_and_grads(params, cfg, g, P, cap)
    else:
        ps = {k: v.detach().clone().requires_grad_(True) for k, v in params.items()}
        ntok = (cap[:, 1:] != cfg.padding_idx).sum().to(g.dtype)
        for lo in range(0, B, chunk):
            lg, _ = R.decoder_train(ps, cfg, g[lo:lo + chunk], P[lo:lo + chunk], cap[lo:lo + chunk, :-1])
            _, loss_sum, _ = R.masked_loss(cfg, lg, cap[lo:lo + chunk])
            (loss_sum / ntok).backward()
        grads = {k: (v.grad if v.grad is not None else torch.zeros_like(v)) for k, v in ps.items()}
    TR.adam_step_torch(params, grads, adam_state, lr=0.00005)


def run_cpu_reference(steps, warmup, sample_B=B_PER_GPU, budget_s=None):
    """Times the oracle (PyTorch-CPU fp32, op-for-op the reference graph incl. the materialised [B,50,H] tensors,
    Python loop over 34 steps for every sample -- the reference has no length input, SURVEY.md F5 -- autograd backward)
    on all host threads, at the bench's own batch.  Returns tokens/s and a description."""
    from oracle import decoder_ref as R
    from oracle import synth
    torch.set_num_threads(os.cpu_count() or 1)
    cfg = R.DecoderConfig(softmax_axis=SOFTMAX_AXIS)
    params = R.init_params(cfg, 0, torch.float32, bias_scale=0.0)
    g, P, cap = synth.make_batch(cfg, sample_B, 20260)
    adam_state = {}
    for _ in range(max(1, warmup)):
        cpu_reference_step(params, cfg, g, P, cap, adam_state)
    t0 = time.perf_counter()
    done = 0
    for _ in range(steps):
        cpu_reference_step(params, cfg, g, P, cap, adam_state)
        done += 1
        if budget_s is not None and time.perf_counter() - t0 > budget_s:
            break
    dt = time.perf_counter() - t0
    return {"value": sample_B * cfg.train_steps * done / dt, "unit": UNIT, "cores": torch.get_num_threads(),
            "kind": "port", "ms_per_step": 1e3 * dt / done, "steps": done,
            "sample": "oracle/decoder_ref.py fwd+bwd + oracle/train_ref.py Adam, batch %d x 34 steps (autograd in chunks of 64 "
                      "samples), fp32, %d thread(s), %d step(s)" % (sample_B, torch.get_num_threads(), done)}


def run_cpu_beam_reference(n_images=64, reps=2):
    """The oracle's beam search (beam 3, 35 steps, finished-handling on) on all host threads, a bounded sample
    of BASELINE.json configs[3]; images/s."""
    from oracle import decoder_ref as R
    from oracle import synth
    torch.set_num_threads(os.cpu_count() or 1)
    cfg = R.DecoderConfig(softmax_axis=SOFTMAX_AXIS)
    params = R.init_params(cfg, 0, torch.float32, bias_scale=0.0)
    g, P, _ = synth.make_batch(cfg, n_images, 777)
    R.beam_search(params, cfg, g[:8], P[:8], BEAM)          # warm-up
    t0 = time.perf_counter()
    for _ in range(reps):
        R.beam_search(params, cfg, g, P, BEAM)
    dt = time.perf_counter() - t0
    return {"value": n_images * reps / dt, "unit": "images/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": "oracle/decoder_ref.beam_search, %d images x beam %d x %d steps, fp32, %d thread(s), %d rep(s)"
                      % (n_images, BEAM, cfg.infer_max_length, torch.get_num_threads(), reps)}


def main_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    res = run_cpu_reference(args.steps, args.warmup, sample_B=args.reference_batch)
    line = {"impl": "reference", "metric": METRIC, "value": res["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": res["steps"], "warmup": args.warmup, "ms_per_step": res["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args.gpus, grad_exchange=args.grad_exchange) if args.reference_batch == B_PER_GPU else
            dict(workload_config(args.gpus, grad_exchange=args.grad_exchange),
                 sample="batch %d per step (--reference-batch)" % args.reference_batch),
            "cpu_baseline": {k: res[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": res["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "extras": {"beam_search_cpu": run_cpu_beam_reference()},
            "note": "PaddlePaddle 1.8 cannot be installed (BASELINE.md section 2): the reference arm is the CPU "
                    "oracle port of model_adaAttention_aic.py on the box's host cores"}
    print(json.dumps(line), flush=True)
    return 0


# ------------------------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------------------------
def timed_region(fn, steps, sync, barrier):
    """barrier + sync, K steps, sync + barrier; CUDA-event time in ms"""
    barrier()
    sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        fn(i)
    e1.record()
    sync()
    barrier()
    return e0.elapsed_time(e1)


def agreed_step_count(budget_ms, local_ms_per_step, world, device):
    """Number of steps that fill `budget_ms`, IDENTICAL on every rank: each data-parallel step is a set of collectives, so
    a count derived from the local step time (ranks differ in the last digits) can differ by one between ranks and leave
    them waiting for each other forever -- it is derived from the slowest rank's time instead."""
    t = torch.tensor([local_ms_per_step], device=device, dtype=torch.float64)
    if world > 1:
        import torch.distributed as dist
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return int(budget_ms / float(t.cpu())) + 1


def stage(msg):
    """progress on stderr (stdout carries the one JSON line)"""
    print("[bench rank %s %.1fs] %s" % (os.environ.get("RANK", "0"), time.perf_counter() - T0, msg), file=sys.stderr, flush=True)


T0 = time.perf_counter()


def main_ours(args):
    import faulthandler
    import torch.distributed as dist
    faulthandler.dump_traceback_later(150, repeat=True, file=sys.stderr)   # a stuck run says where (stderr)
    import myimagecaptioningmodel_b200 as M
    from myimagecaptioningmodel_b200 import _capi
    from myimagecaptioningmodel_b200 import synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit("--gpus %d does not match WORLD_SIZE %d" % (args.gpus, world))
    if args.gpus > 1 and world == 1:
        raise SystemExit("launch N>1 with: python -m torch.distributed.run --nproc-per-node N bench.py --gpus N ...")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    barrier = (lambda: dist.barrier()) if world > 1 else (lambda: None)
    sync = torch.cuda.synchronize

    cfg = M.DecoderConfig(dtype="bf16", softmax_axis=SOFTMAX_AXIS)
    eng = M.CaptionDecoderEngine(cfg, seed=0)           # same weights on every rank (replicated)
    trainer = M.DataParallelTrainer(eng, grad_exchange=args.grad_exchange)
    B = args.batch_per_gpu      # 256 = BASELINE.json configs[1]; configs[2] at 2 / 4 GPUs (global 2048) is --batch-per-gpu 1024 / 512
    g_h, P_h, cap_h = synth.make_batch(cfg, B, 20260 + rank)
    g_h, P_h, cap_h = g_h.pin_memory(), P_h.pin_memory(), cap_h.pin_memory()
    g_d, P_d, cap_d = g_h.cuda(), P_h.cuda(), cap_h.cuda()
    ntok_nonpad = float((cap_h[:, 1:] != 0).sum())
    # host-side token counts of the batch (what a data loader knows before the upload): with them the time-batched head
    # runs on the rows that carry a target only (capdec_train_fwd `lengths`) -- same loss and gradients
    lengths_h = M.caption_lengths(cap_h)

    opt_dec = M.Adam([(eng.store.flat, eng.store.flat_grad)], 0.00005)     # train.py:26-31 / config.py:32-33 defaults

    def step_resident(i):
        eng.bind_params()
        trainer.forward_backward(g_d, P_d, cap_d, optimizer=opt_dec, lengths=lengths_h)   # Adam per gradient group, beside the backward pass

    def step_resident_all_rows(i):        # the same step without `lengths`: all B x 34 rows through the head
        eng.bind_params()
        trainer.forward_backward(g_d, P_d, cap_d, optimizer=opt_dec)

    # ---- e2e: host buffers, double-buffered upload on a side stream, loss read back every step ----
    copy_stream = torch.cuda.Stream()
    bufs = [(torch.empty_like(g_d), torch.empty_like(P_d), torch.empty_like(cap_d)) for _ in range(2)]
    ready = [torch.cuda.Event() for _ in range(2)]
    consumed = [torch.cuda.Event() for _ in range(2)]
    losses = []

    def upload(slot):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[slot])
            for dst, src in zip(bufs[slot], (g_h, P_h, cap_h)):
                dst.copy_(src, non_blocking=True)
            ready[slot].record(copy_stream)

    # the loss of every step is copied to pinned host memory inside the step (4 bytes D2H) and read by the host one step
    # later, when its copy event has fired -- the way a training loop logs it -- so the GPU is never idle while the host
    # enqueues the next step; the last value is read after the closing synchronisation of the timed region
    loss_host = [torch.zeros(1, dtype=torch.float32).pin_memory() for _ in range(2)]
    loss_evt = [torch.cuda.Event() for _ in range(2)]
    # the 4-byte D2H copy runs on its own stream from a device-side staging scalar, so that the compute stream never waits
    # for a copy-engine operation between two steps
    d2h_stream = torch.cuda.Stream()
    loss_stage = [torch.zeros(1, dtype=torch.float32, device="cuda") for _ in range(2)]
    staged = [torch.cuda.Event() for _ in range(2)]
    pending = []

    def drain(block, only=None):
        while pending and (block or loss_evt[pending[0]].query()):
            if only is not None and pending[0] != only:
                break
            k = pending.pop(0)
            loss_evt[k].synchronize()
            losses.append(float(loss_host[k][0]))
            if only is not None:
                break

    def step_e2e(i):
        slot = i & 1
        if i == 0:
            upload(0)
        upload(slot ^ 1)                                  # prefetch the next batch while this one computes
        torch.cuda.current_stream().wait_event(ready[slot])
        eng.bind_params()
        loss = trainer.forward_backward(*bufs[slot], optimizer=opt_dec, lengths=M.caption_lengths(cap_h))
        consumed[slot].record()
        k = i & 1
        if k in pending:                                  # the buffer about to be reused (written two steps ago)
            drain(True, only=k)
        loss_stage[k].copy_(loss.reshape(1))                # device-side copy on the compute stream
        staged[k].record()
        with torch.cuda.stream(d2h_stream):
            d2h_stream.wait_event(staged[k])
            loss_host[k].copy_(loss_stage[k], non_blocking=True)   # D2H read of the step's result
            loss_evt[k].record(d2h_stream)
        pending.append(k)
        drain(False)

    for c in consumed:
        c.record()
    # rank 0 prints the line and samples its own GPU; N polling nvidia-smi processes would only contend for the driver
    with ClockSampler(local_rank, enabled=(rank == 0)) as clocks:   # started before the warm-up: nvidia-smi needs a moment
        stage("warm-up")
        for i in range(args.warmup):
            step_resident(i)
        sync()
        stage("timed region")
        l0 = eng.launch_count
        with clocks.window():
            ms = timed_region(step_resident, args.steps, sync, barrier)
        launches = eng.launch_count - l0 + args.steps * eng.N_GRAD_EVENTS     # + one capdec_adam_step launch per gradient group
        stage("all-rows leg")
        for i in range(3):
            step_resident_all_rows(i)
        sync()
        ms_all_rows = timed_region(step_resident_all_rows, args.steps, sync, barrier)
        stage("e2e leg")
        for i in range(max(2, args.warmup // 2)):
            step_e2e(i)
        sync()
        drain(True)
        for c in consumed:
            c.record()
        with clocks.window():
            ms_e2e = timed_region(step_e2e, args.steps, sync, barrier)
        drain(True)
    clk = clocks.summary()

    # host link probe: the same pinned buffers copied alone (no compute beside them) -- what the e2e leg is bound by when
    # the upload of one batch takes longer than the device needs for a step
    pe0, pe1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync()
    with torch.cuda.stream(copy_stream):
        pe0.record(copy_stream)
        for _ in range(4):
            for dst, src in zip(bufs[0], (g_h, P_h, cap_h)):
                dst.copy_(src, non_blocking=True)
        pe1.record(copy_stream)
    sync()
    ms_h2d = pe0.elapsed_time(pe1) / 4

    stage("beam search")
    # ---- beam search (BASELINE.json configs[3], 512 images per GPU, beam 3, 35 steps): secondary number ----
    gi, Pi, _ = synth.make_batch(cfg, BEAM_IMAGES_PER_GPU, 777 + rank)
    gi, Pi = gi.cuda(), Pi.cuda()
    for _ in range(2):   # warm-up at the full size: the handle's decode workspace is allocated here, not in the timed region
        eng.decode(gi, Pi, beam=BEAM, finish_on_stop=True)
    sync()
    beam_reps = 3
    ms_beam = timed_region(lambda i: eng.decode(gi, Pi, beam=BEAM, finish_on_stop=True), beam_reps, sync, barrier)

    # ---- batch-1 greedy decode latency (BASELINE.json configs[4]) ----
    g1, P1 = gi[:1].contiguous(), Pi[:1].contiguous()
    for _ in range(3):
        eng.decode(g1, P1, beam=1)
    sync()
    lat = []
    for _ in range(200):                                   # median / p99 over 200 runs
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        eng.decode(g1, P1, beam=1)
        e1.record()
        sync()
        lat.append(e0.elapsed_time(e1))
    lat.sort()

    # ---- widened path (SURVEY.md 8f N1 + N2): encoder map in, parameters updated -- secondary numbers ----
    # step = bind (decoder + adaptor) + capdec_img2feature + train_fwd + train_bwd + capdec_img2feature_bwd
    #        (+ the data-parallel exchanges) + capdec_adam_step over both flat buffers (PaddlePaddle 1.8 rule)
    stage("widened step")
    X_ENC = 1280
    eng.attach_adaptor(X_ENC, seed=0)
    opt, _ = M.get_optimizer(eng)
    x_d = synth.make_encoder_map(cfg, B, X_ENC, 4242 + rank).cuda()

    def step_full(i):
        eng.bind_params()
        eng.bind_adaptor()
        trainer.forward_backward_map(x_d, cap_d, optimizer=opt, lengths=lengths_h)

    for i in range(3):
        step_full(i)
    sync()
    l0 = eng.launch_count
    ms_full = timed_region(step_full, args.steps, sync, barrier)
    launches_full = eng.launch_count - l0 + (eng.N_GRAD_EVENTS + 1) * args.steps
    adam_reps = 10
    ms_adam = timed_region(lambda i: opt.step(), adam_reps, sync, lambda: None) / adam_reps
    n_adam = sum(p.numel() for p, _ in opt.buffers)
    adam_gbs = 28.0 * n_adam / (ms_adam * 1e-3) / 1e9      # 16 B read (p, g, m, v) + 12 B written (p, m, v) per element
    eng.store.init_reference(0)                              # the weights the remaining legs are measured on

    # ---- roofline of the dominant kernel: the tcgen05 GEMM at the head shape (M = B*T, N = K = 1024) ----
    # Timed live with CUDA events on the launching stream; the launches rotate over 8 activation / output buffer pairs
    # (8 x 35.6 MB = 285 MB > the 126 MB L2), so every launch streams its A operand from HBM as it does inside a step.
    Mh, Nh, Kh = B * cfg.train_steps, cfg.hidden_dim, cfg.hidden_dim
    NBUF = 8
    As = [torch.randn(Mh, Kh, device="cuda").to(torch.bfloat16) for _ in range(NBUF)]
    W = torch.randn(Kh, Nh, device="cuda").to(torch.bfloat16)      # PaddlePaddle [in,out] weight = MN-major B operand
    Cs = [torch.empty(Mh, Nh, device="cuda", dtype=torch.bfloat16) for _ in range(NBUF)]
    bias = torch.zeros(Nh, device="cuda")
    lib = _capi.load_library()

    def head_gemm(i):
        A, Cb = As[i % NBUF], Cs[i % NBUF]
        _capi.check(lib.capdec_gemm(_capi.DTYPE_BF16, 2, A.data_ptr(), Kh, W.data_ptr(), Nh, Cb.data_ptr(), Nh, Mh, Nh, Kh,
                                    bias.data_ptr(), None, 0, None, 0, 1, 0, 0, 0, _capi.stream_ptr()), "capdec_gemm")

    for i in range(NBUF):
        head_gemm(i)
    reps = 3 * NBUF
    ms_gemm = timed_region(head_gemm, reps, sync, lambda: None) / reps
    peaks = load_peaks()
    flops = 2.0 * Mh * Nh * Kh
    achieved = flops / (ms_gemm * 1e-3) / 1e12
    head_isolated = {"kernel": "gemm_tc_pair_kernel<A K-major,B MN-major> (tcgen05 cta_group::2, 256x256 tile per CTA pair) "
                               "head fc M=%d N=%d K=%d bias+tanh, timed alone back to back" % (Mh, Nh, Kh),
                     "bound": "tensor", "achieved": achieved, "peak": peaks["bf16_burst"], "unit": "TFLOP/s",
                     "frac": achieved / peaks["bf16_burst"], "us": 1e3 * ms_gemm, "algorithmic_flops": flops,
                     "peak_source": peaks["source"] + ", burst", "l2": "8 rotating A/C buffer pairs, 285 MB > L2"}

    # ---- per-kernel roofline table, timed LIVE inside the step: the library's phase timeline records CUDA events on the
    # launching stream around these kernels while the bench's own step runs (stream order, warm caches, sustained clocks)
    def phase_profile(fn, reps=5):
        for i in range(2):
            fn(i)
        sync()
        eng.profile(True)
        for i in range(reps):
            fn(i)
        rep_ms = eng.profile_report()
        eng.profile(False)
        return {k: 1e3 * v / reps for k, v in rep_ms.items()}      # us per step

    stage("phase profile")
    ph = phase_profile(step_resident)
    ph_dec = phase_profile(lambda i: eng.decode(gi, Pi, beam=BEAM, finish_on_stop=True), reps=2)
    Hh, Ee, Vv, kk, Tt = cfg.hidden_dim, cfg.embedding_size, cfg.vocab_size, cfg.num_positions, cfg.train_steps
    Vp = (Vv + 127) // 128 * 128
    Mrows, Mlive = B * Tt, int(ntok_nonpad)                          # all rows / rows that carry a target (the head's rows)
    traffic_db = load_ncu_traffic_table()

    def entry(key, kernel, bound, work, us, unit_scale, peak, unit, what):
        ach = work / (us * 1e-6) / unit_scale if us and us > 0 else None
        return {"phase": key, "kernel": kernel, "bound": bound, "achieved": ach, "peak": peak, "unit": unit,
                "frac": (ach / peak) if ach else None, "us": us, "algorithmic_work": work, "work_is": what,
                "traffic": traffic_db.get(key)}

    TF, GB = 1e12, 1e9
    sus, hbm = peaks["bf16_sustained"], peaks["hbm_gbs"]
    mufu_peak = 16.0 * 148 * 1.965e9 / 1e9                           # MUFU ops / ns: 16 per clock per SM at 1965 MHz
    kernels = [
        entry("bwd.recurrence", "lstm_seq_bwd_kernel (persistent, all 34 BPTT steps: d h = d gates . W_hh^T + cell backward)",
              "tensor", 2.0 * B * Hh * 4 * Hh * (Tt - 1), ph.get("bwd.recurrence"), TF, sus, "TFLOP/s",
              "2*B*H*4H flops x 33 steps"),
        entry("fwd.recurrence", "lstm_seq_fwd_kernel (persistent, all 34 steps: gates = h . W_hh, cell in the MMA epilogue)",
              "tensor", 2.0 * B * Hh * 4 * Hh * Tt, ph.get("fwd.recurrence"), TF, sus, "TFLOP/s", "2*B*H*4H flops x 34 steps"),
        entry("fwd.xe_gemm", "gemm_tc_pair_kernel short-K variant: x-part pre-activations M=%d N=5H K=E" % Mrows, "tensor",
              2.0 * Mrows * Ee * 5 * Hh, ph.get("fwd.xe_gemm"), TF, sus, "TFLOP/s", "2*M*E*5H flops"),
        entry("fwd.out_proj_vocab", "gemm_tc_pair_kernel short-K variant: fp32 logits M=%d N=V K=E (output-write bound: 128 "
              "flop per byte written < the 212 flop/B ridge)" % Mlive, "hbm", 4.0 * Mlive * Vp + 2.0 * (Mlive + Vv) * Ee,
              ph.get("fwd.out_proj_vocab"), GB, hbm, "GB/s", "fp32 logits written + bf16 operands read"),
        entry("fwd.cross_entropy", "ce_fused_kernel (online log-sum-exp + d logits)", "hbm",
              4.0 * Mlive * Vv + 2.0 * Mlive * Vp, ph.get("fwd.cross_entropy"), GB, hbm, "GB/s",
              "fp32 logits read + bf16 d logits written"),
        entry("fwd.attention", "attn_scores_kernel + attn_sadd + batched context GEMM + gather (phase)", "mufu",
              1.0 * Mlive * (kk + 1) * Hh, ph.get("fwd.attention"), 1e9, mufu_peak, "Gtanh/s",
              "one MUFU.TANH per (row, position, column)"),
        entry("dec.beam_topk", "beam_rows_kernel<49> + beam_merge_kernel, 35 steps of %d rows" % (BEAM_IMAGES_PER_GPU * BEAM),
              "hbm", 4.0 * BEAM_IMAGES_PER_GPU * BEAM * Vv * cfg.infer_max_length, ph_dec.get("dec.beam_topk"), GB, hbm,
              "GB/s", "fp32 logits of every row read once per step"),
        entry("dec.attention", "attn_scores_kernel + attn_ctx_small_kernel, 35 steps", "hbm",
              2.0 * 2 * BEAM_IMAGES_PER_GPU * kk * Hh * cfg.infer_max_length, ph_dec.get("dec.attention"), GB, hbm, "GB/s",
              "bf16 keys + values of every image read once per step"),
    ]
    # the `roofline` object of the line = the longest single kernel of the train step (hbm / tensor bound), with the table
    main = max((e for e in kernels if e["bound"] in ("hbm", "tensor") and e["phase"].split(".")[0] in ("fwd", "bwd")
                and e["us"]), key=lambda e: e["us"])
    roofline = {"kernel": main["kernel"], "bound": main["bound"], "achieved": main["achieved"], "peak": main["peak"],
                "unit": main["unit"], "frac": main["frac"], "traffic": main["traffic"],
                "peak_source": peaks["source"] + (", sustained (timed inside the step)" if main["bound"] == "tensor" else ""),
                "ms_per_launch": main["us"] * 1e-3, "algorithmic_work_per_launch": main["algorithmic_work"],
                "work_is": main["work_is"], "timed": "CUDA events on the launching stream around the kernel, inside the "
                "bench step (capdec_profile), average of 5 steps", "kernels": kernels, "head_fc_isolated": head_isolated}
    # whole-step view: algorithmic GEMM flops of one train step over the measured step time
    step_flops = 116.7e6 * B * cfg.train_steps + 616.6e6 * B
    step_tflops = step_flops / (ms / args.steps * 1e-3) / 1e12

    # ---- sustained: >= 5 s of back-to-back steps (the timed region above lasts tens of milliseconds: a burst), with the
    # clocks sampled inside it ----
    stage("sustained leg")
    n_sus = agreed_step_count(5500.0, ms / args.steps, world, "cuda")
    with ClockSampler(local_rank, enabled=(rank == 0)) as clocks_s:
        for i in range(3):
            step_resident(i)
        with clocks_s.window():
            ms_sus = timed_region(step_resident, n_sus, sync, barrier)
    clk_sus = clocks_s.summary()

    # ---- data-parallel parity (N > 1): the all-reduced shard gradients against ONE full-batch step on the same weights,
    # computed locally on rank 0 from the gathered shards (train.py:119-127 is what the
    # data-parallel wrapper replaces) ----
    dp_parity = None
    stage("dp parity")
    if world > 1:
        eng.store.init_reference(0)
        eng.bind_params()
        loss_dp = trainer.forward_backward(g_d, P_d, cap_d, lengths=lengths_h)
        sync()
        grad_dp = eng.store.flat_grad.clone()
        gathered = []
        for x in (g_d, P_d, cap_d):
            parts = [torch.empty_like(x) for _ in range(world)]
            dist.all_gather(parts, x)
            gathered.append(torch.cat(parts, 0))
        if rank == 0:
            lb, _ = eng.train_forward(gathered[0], gathered[1], gathered[2], lengths=M.caption_lengths(gathered[2].cpu()))
            eng.train_backward(want_dg=False, want_dP=False)
            sync()
            ref = eng.store.flat_grad.double()
            dp_parity = {"flat_grad_rel_err": float((grad_dp.double() - ref).norm() / ref.norm()),
                         "flat_grad_max_err_over_max": float((grad_dp.double() - ref).abs().max() / ref.abs().max()),
                         "loss_rel_diff": abs(float(loss_dp) - float(lb[2])) / abs(float(lb[2])),
                         "global_batch": int(gathered[0].shape[0]),
                         "what": "all-reduced gradient of %d shards vs one local step at the global batch, bf16 storage "
                                 "(the split changes the fp32 summation order and the bf16 rounding points)" % world}
        del gathered
        barrier()

    # max over ranks of the device times
    t = torch.tensor([ms, ms_e2e, ms_beam, ms_full, ms_all_rows, ms_sus], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e2e, ms_beam, ms_full, ms_all_rows, ms_sus = (float(x) for x in t.cpu())
    tokens_per_step = world * B * cfg.train_steps
    value = tokens_per_step * args.steps / (ms * 1e-3)
    e2e_value = tokens_per_step * args.steps / (ms_e2e * 1e-3)
    h2d = sum(x.numel() * x.element_size() for x in (g_h, P_h, cap_h))

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        res = run_cpu_reference(steps=1000, warmup=1, sample_B=B, budget_s=15.0)
        cpu_baseline = {k: res[k] for k in ("value", "unit", "cores", "kind", "sample")}
        cpu_baseline["beam_search"] = run_cpu_beam_reference()      # the secondary metric's CPU figure (images/s)

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "bf16", "data": "synthetic", "config": workload_config(world, B, args.grad_exchange),
                "clocks": clk, "gpu_launches": int(launches),
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                        "ms_per_step": ms_e2e / args.steps, "h2d_alone_ms_per_step": ms_h2d,
                        "h2d_alone_gbs": h2d / (ms_h2d * 1e-3) / 1e9},
                "roofline": roofline, "cpu_baseline": cpu_baseline,
                "extras": {"nonpad_tokens_per_sec": world * ntok_nonpad * args.steps / (ms * 1e-3),
                           "live_rows_per_gpu": int(ntok_nonpad), "rows_per_gpu": B * cfg.train_steps,
                           "sustained": {"seconds": ms_sus * 1e-3, "steps": n_sus, "ms_per_step": ms_sus / n_sus,
                                         "tokens_per_sec": tokens_per_step * n_sus / (ms_sus * 1e-3), "clocks": clk_sus},
                           "dp_parity": dp_parity,
                           "without_lengths": {"ms_per_step": ms_all_rows / args.steps,
                                               "tokens_per_sec": tokens_per_step * args.steps / (ms_all_rows * 1e-3),
                                               "note": "the same step with lengths=NULL: all B x 34 rows through the head"},
                           "step_gemm_tflops_algorithmic": step_tflops,
                           "step_gemm_frac_of_sustained_peak": step_tflops / peaks["bf16_sustained"],
                           "beam_search_images_per_sec": world * BEAM_IMAGES_PER_GPU * beam_reps / (ms_beam * 1e-3),
                           "beam_config": "beam %d, %d steps, %d images per GPU, finish_on_stop" % (
                               BEAM, cfg.infer_max_length, BEAM_IMAGES_PER_GPU),
                           "batch1_greedy_latency_ms": {"median": lat[len(lat) // 2], "p95": lat[int(0.95 * (len(lat) - 1))],
                                                        "p99": lat[int(0.99 * (len(lat) - 1))], "runs": len(lat),
                                                        "per_step_median": lat[len(lat) // 2] / cfg.infer_max_length,
                                                        "config": "1 image, beam 1, %d steps" % cfg.infer_max_length},
                           "full_step_from_encoder_map": {
                               "tokens_per_sec": tokens_per_step * args.steps / (ms_full * 1e-3),
                               "ms_per_step": ms_full / args.steps, "gpu_launches": int(launches_full),
                               "step": "bind(decoder+adaptor) + img2feature + train_fwd + train_bwd + img2feature_bwd "
                                       "(+allreduce) + adam (decoder+adaptor), encoder_channel %d" % X_ENC},
                           "adam_step": {"ms": ms_adam, "parameters": n_adam, "achieved_gbs": adam_gbs,
                                         "frac_of_hbm_peak": adam_gbs / peaks["hbm_gbs"],
                                         "algorithmic_bytes_per_parameter": 28},
                           "workspace_gb": eng.workspace_bytes / 1e9, "last_loss": losses[-1] if losses else None}}
        print(json.dumps(line), flush=True)
    faulthandler.cancel_dump_traceback_later()
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--batch-per-gpu", type=int, default=B_PER_GPU,
                    help="captions per GPU and step (default 256 = BASELINE.json configs[1]; configs[2] keeps the global batch at "
                         "2048: 1024 at 2 GPUs, 512 at 4, 256 at 8)")
    ap.add_argument("--grad-exchange", default="bf16", choices=["fp32", "bf16"],
                    help="wire type of the data-parallel gradient all-reduce (default bf16: 2.85 vs 2.96 ms per step at 8 GPUs, "
                         "flat-gradient error 2.5e-3 vs 2.4e-4 against the local full-batch step; the trainer's own default "
                         "is fp32)")
    ap.add_argument("--reference-batch", type=int, default=B_PER_GPU,
                    help="batch of the --impl reference arm (default: the bench's own 256; the CPU contract test uses a smaller one)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        return main_reference(args)
    return main_ours(args)


if __name__ == "__main__":
    sys.exit(main())
