"""Device-side phase timeline of one train step and one beam decode (CUDA events inside libcapdec.so, no profiler):
where the step time goes, in stream order, warm caches.  python scripts/phase_times.py [--batch 256] [--reps 5]"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import myimagecaptioningmodel_b200 as M  # noqa: E402
from myimagecaptioningmodel_b200 import synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=256)
ap.add_argument("--decode-images", type=int, default=512)
ap.add_argument("--beam", type=int, default=3)
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--axis", default="position")
ap.add_argument("--from-map", action="store_true", help="widened step: feature adaptor in front, Adam behind (8f N1, N2)")
ap.add_argument("--no-lengths", action="store_true", help="all B x 34 rows through the head (no packed rows)")
args = ap.parse_args()

cfg = M.DecoderConfig(dtype="bf16", softmax_axis=args.axis)
eng = M.CaptionDecoderEngine(cfg, seed=0)
g, P, cap = synth.make_batch(cfg, args.batch, 20260)
lengths = None if args.no_lengths else M.caption_lengths(cap)
g, P, cap = g.cuda(), P.cuda(), cap.cuda()


def step():
    eng.bind_params()
    eng.train_forward(g, P, cap, lengths=lengths)
    eng.train_backward(want_dg=True, want_dP=True)


if args.from_map:
    eng.attach_adaptor(1280, seed=0)
    opt, _ = M.get_optimizer(eng)
    x = synth.make_encoder_map(cfg, args.batch, 1280, 4242).cuda()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]

    def step():  # noqa: F811
        eng.bind_params()
        eng.bind_adaptor()
        eng.train_forward_map(x, cap, lengths=lengths)
        eng.train_backward_map(want_dmap=True)
        ev[0].record()
        opt.step()
        ev[1].record()

for _ in range(3):
    step()
torch.cuda.synchronize()
eng.profile(True)
for _ in range(args.reps):
    step()
rep = eng.profile_report()
total = sum(rep.values())
print("train step B=%d: %.3f ms (sum of phases, avg of %d)" % (args.batch, total / args.reps, args.reps))
for k, v in rep.items():
    print("  %-24s %8.1f us  %5.1f%%" % (k, 1e3 * v / args.reps, 100 * v / total))
out = {"train_us": {k: 1e3 * v / args.reps for k, v in rep.items()}}
if args.from_map:
    torch.cuda.synchronize()
    a = ev[0].elapsed_time(ev[1])
    print("  %-24s %8.1f us  (capdec_adam_step x2, last step; outside the handle's phase marks)" % ("adam", 1e3 * a))
    out["adam_us"] = 1e3 * a
if args.decode_images:
    gi, Pi, _ = synth.make_batch(cfg, args.decode_images, 777)
    gi, Pi = gi.cuda(), Pi.cuda()
    eng.profile(False)
    eng.decode(gi, Pi, beam=args.beam, finish_on_stop=True)
    torch.cuda.synchronize()
    eng.profile(True)
    for _ in range(args.reps):
        eng.decode(gi, Pi, beam=args.beam, finish_on_stop=True)
    rep = eng.profile_report()
    total = sum(rep.values())
    print("beam decode %d images x beam %d: %.3f ms" % (args.decode_images, args.beam, total / args.reps))
    for k, v in rep.items():
        print("  %-24s %8.1f us  %5.1f%%" % (k, 1e3 * v / args.reps, 100 * v / total))
    out["decode_us"] = {k: 1e3 * v / args.reps for k, v in rep.items()}
print(json.dumps(out))
