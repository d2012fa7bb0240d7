"""2-GPU (or N-GPU) data-parallel parity check, run under torch.distributed.run:
every rank trains on its shard through DataParallelTrainer (global token normalisation + overlapped bucketed all-reduce)
and compares the all-reduced gradient with a full-batch step computed locally on the same weights."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import myimagecaptioningmodel_b200 as M  # noqa: E402
from myimagecaptioningmodel_b200 import synth  # noqa: E402
from myimagecaptioningmodel_b200.parallel import shard_range  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
cfg = M.DecoderConfig(dtype="bf16", softmax_axis="position")
B = 64 * world
g, P, cap = synth.make_batch(cfg, B, 4242)            # the same global batch on every rank
g, P, cap = g.cuda(), P.cuda(), cap.cuda()
eng = M.CaptionDecoderEngine(cfg, seed=0)
ref = M.CaptionDecoderEngine(cfg, seed=0)
trainer = M.DataParallelTrainer(eng)
lo, hi = shard_range(B, rank, world)
for it in range(3):                                    # repeated: the events / comm stream are reused
    eng.bind_params()
    loss = trainer.forward_backward(g[lo:hi].contiguous(), P[lo:hi].contiguous(), cap[lo:hi].contiguous())
torch.cuda.synchronize()
ref.bind_params()
lb, _ = ref.train_forward(g, P, cap)
ref.train_backward()
torch.cuda.synchronize()
a, b = eng.store.flat_grad.double(), ref.store.flat_grad.double()
worst, worst_name = 0.0, ""
for name in M.PARAM_NAMES:
    x, y = eng.store.grads[name].double(), ref.store.grads[name].double()
    if float(y.abs().max()) == 0.0 or name == "fc_a_b":
        continue
    e = float((x - y).norm() / y.norm())
    if e > worst:
        worst, worst_name = e, name
err = float((a - b).norm() / b.norm())
dl = abs(float(loss) - float(lb[2])) / abs(float(lb[2]))
print("rank %d: overlap=%s  loss rel diff %.2e  flat grad rel err %.2e  worst tensor %.2e (%s)" % (
    rank, getattr(trainer, "_overlap", None), dl, err, worst, worst_name), flush=True)
assert dl < 1e-3 and err < 1e-2 and worst < 2e-2
dist.barrier()
if rank == 0:
    print("DP CHECK OK")
dist.destroy_process_group()
