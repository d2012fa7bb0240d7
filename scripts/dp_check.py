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
DTYPE = os.environ.get("DP_CHECK_DTYPE", "bf16")
# fp32 storage: the shard decomposition only changes the summation order -> every tensor must agree to 1e-4 (this is the
# check of the data-parallel LOGIC: global normalisation, overlap, group slices).  bf16 storage: gradients that are sums of
# nearly cancelling bf16-rounded terms (fc_s_b: the sentinel-score bias) carry rounding noise that depends on how the batch
# is split, so only the whole gradient (1e-2) and a loose per-tensor bound are asserted there.
TOL_FLAT, TOL_TENSOR = (1e-4, 1e-4) if DTYPE == "fp32" else (1e-2, 2e-1)
cfg = M.DecoderConfig(dtype=DTYPE, softmax_axis="position")
B = (64 if DTYPE == "bf16" else 16) * world
g, P, cap = synth.make_batch(cfg, B, 4242)            # the same global batch on every rank
g, P, cap = g.cuda(), P.cuda(), cap.cuda()
eng = M.CaptionDecoderEngine(cfg, seed=0)
ref = M.CaptionDecoderEngine(cfg, seed=0)
trainer = M.DataParallelTrainer(eng, grad_exchange=os.environ.get("DP_CHECK_EXCHANGE", "fp32"))
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
print("rank %d: %s overlap=%s  loss rel diff %.2e  flat grad rel err %.2e  worst tensor %.2e (%s)" % (
    rank, DTYPE, getattr(trainer, "_overlap", None), dl, err, worst, worst_name), flush=True)
assert dl < 1e-3 and err < TOL_FLAT and worst < TOL_TENSOR

# the same from the encoder map (feature adaptor in front): decoder + adaptor gradients, all-reduced
for e in (eng, ref):
    e.attach_adaptor(1280, seed=0)
x = synth.make_encoder_map(cfg, B, 1280, 99).cuda()
for it in range(2):
    eng.bind_params()
    eng.bind_adaptor()
    loss, _ = trainer.forward_backward_map(x[lo:hi].contiguous(), cap[lo:hi].contiguous())
torch.cuda.synchronize()
ref.bind_params()
ref.bind_adaptor()
lb, _ = ref.train_forward_map(x, cap)
ref.train_backward_map(want_dmap=False)
torch.cuda.synchronize()
a, b = eng.store.flat_grad.double(), ref.store.flat_grad.double()
aa, ab = eng.adaptor.flat_grad.double(), ref.adaptor.flat_grad.double()
err, aerr = float((a - b).norm() / b.norm()), float((aa - ab).norm() / ab.norm())
dl = abs(float(loss) - float(lb[2])) / abs(float(lb[2]))
print("rank %d: from the encoder map: loss rel diff %.2e  flat grad rel err %.2e  adaptor grad rel err %.2e" % (
    rank, dl, err, aerr), flush=True)
assert dl < 1e-3 and err < TOL_FLAT and aerr < TOL_FLAT
dist.barrier()
if rank == 0:
    print("DP CHECK OK")
dist.destroy_process_group()
