"""Phase timeline of the data-parallel train step (torch.distributed.run, N ranks): the handle's device-side phase marks on
the main stream plus the exposed tail (end of capdec_train_bwd -> end of the last all-reduce on the communication stream).
Compare with scripts/phase_times.py (1 GPU) to see which phases the overlapping NCCL kernels slow down."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import myimagecaptioningmodel_b200 as M  # noqa: E402
from myimagecaptioningmodel_b200 import synth  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
cfg = M.DecoderConfig(dtype="bf16", softmax_axis="position")
eng = M.CaptionDecoderEngine(cfg, seed=0)
trainer = M.DataParallelTrainer(eng)
g, P, cap = synth.make_batch(cfg, 256, 20260 + rank)
lengths = M.caption_lengths(cap)
g, P, cap = g.cuda(), P.cuda(), cap.cuda()
opt = M.Adam([(eng.store.flat, eng.store.flat_grad)], 0.00005)      # the bench's step: Adam per gradient range
REPS = 10
e_bwd_end = [torch.cuda.Event(enable_timing=True) for _ in range(REPS)]
e_step_end = [torch.cuda.Event(enable_timing=True) for _ in range(REPS)]
e_step_start = [torch.cuda.Event(enable_timing=True) for _ in range(REPS)]
orig_bwd = eng.train_backward
cur = {"i": 0}


def bwd_marked(*a, **k):
    r = orig_bwd(*a, **k)
    e_bwd_end[cur["i"]].record()
    return r


eng.train_backward = bwd_marked
for i in range(5):
    eng.bind_params()
    trainer.forward_backward(g, P, cap, optimizer=opt, lengths=lengths)
torch.cuda.synchronize()
dist.barrier()
eng.profile(True)
for i in range(REPS):
    cur["i"] = i
    e_step_start[i].record()
    eng.bind_params()
    trainer.forward_backward(g, P, cap, optimizer=opt, lengths=lengths)
    e_step_end[i].record()
torch.cuda.synchronize()
rep = eng.profile_report()
tail = sum(a.elapsed_time(b) for a, b in zip(e_bwd_end, e_step_end)) / REPS
step = sum(a.elapsed_time(b) for a, b in zip(e_step_start, e_step_end)) / REPS
if rank == 0:
    total = sum(rep.values())
    print("data-parallel train step, %d GPUs, B=256 per GPU: %.3f ms per step (events), phases sum %.3f ms, exposed tail %.1f us"
          % (world, step, total / REPS, 1e3 * tail))
    for k, v in rep.items():
        print("  %-24s %8.1f us  %5.1f%%" % (k, 1e3 * v / REPS, 100 * v / total))
dist.barrier()
dist.destroy_process_group()
