"""Profiling target: warm up, then ONE train step (bind + fwd + bwd) and optionally one beam decode between
cudaProfilerStart/Stop.  Use with  ncu --profile-from-start off ...   (see scripts/gpu_profile.sh)."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import myimagecaptioningmodel_b200 as M  # noqa: E402
from myimagecaptioningmodel_b200 import synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=256)
ap.add_argument("--decode-images", type=int, default=0)
ap.add_argument("--beam", type=int, default=3)
ap.add_argument("--axis", default="position")
ap.add_argument("--no-train", action="store_true")
ap.add_argument("--from-map", action="store_true", help="widened step: feature adaptor in front, Adam behind")
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

    def step():  # noqa: F811
        eng.bind_params()
        eng.bind_adaptor()
        eng.train_forward_map(x, cap, lengths=lengths)
        eng.train_backward_map(want_dmap=True)
        opt.step()

if not args.no_train:
    for _ in range(2):
        step()
if args.decode_images:
    gi, Pi, _ = synth.make_batch(cfg, args.decode_images, 777)
    gi, Pi = gi.cuda(), Pi.cuda()
    eng.decode(gi, Pi, beam=args.beam, finish_on_stop=True)
torch.cuda.synchronize()
torch.cuda.profiler.start()
if not args.no_train:
    step()
if args.decode_images:
    eng.decode(gi, Pi, beam=args.beam, finish_on_stop=True)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("profiled region done; launches so far:", eng.launch_count)
