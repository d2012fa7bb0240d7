"""Per-step device timeline of the persistent recurrence kernels (capdec_debug_rec_trace, globaltimer stamps).
python scripts/rec_trace.py [--batch 256]"""
import argparse
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import myimagecaptioningmodel_b200 as M  # noqa: E402
from myimagecaptioningmodel_b200 import _capi, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=256)
args = ap.parse_args()
cfg = M.DecoderConfig(dtype="bf16", softmax_axis="position")
eng = M.CaptionDecoderEngine(cfg, seed=0)
g, P, cap = synth.make_batch(cfg, args.batch, 20260)
g, P, cap = g.cuda(), P.cuda(), cap.cuda()
for _ in range(3):
    eng.bind_params(); eng.train_forward(g, P, cap); eng.train_backward()
torch.cuda.synchronize()
lib = _capi.load_library()
lib.capdec_debug_rec_trace.argtypes = [C.c_void_p]
lib.capdec_debug_rec_trace.restype = None
T = cfg.sentence_length - 1
NCTA = 148
buf = torch.zeros(NCTA * T * 8, dtype=torch.int64, device="cuda")
lib.capdec_debug_rec_trace(C.c_void_p(buf.data_ptr()))
eng.train_forward(g, P, cap)
torch.cuda.synchronize()
fwd = buf.cpu().view(NCTA, T, 8).clone()
buf.zero_()
eng.train_backward()
torch.cuda.synchronize()
bwd = buf.cpu().view(NCTA, T, 8).clone()
lib.capdec_debug_rec_trace(None)


def show(name, tr, labels, order):
    used = [c for c in range(NCTA) if int(tr[c].max()) > 0]
    print("%s: %d CTAs traced" % (name, len(used)))
    t0 = min(int(tr[c][tr[c] > 0].min()) for c in used)
    t1 = max(int(tr[c].max()) for c in used)
    print("  kernel span %.1f us -> %.2f us/step" % ((t1 - t0) / 1e3, (t1 - t0) / 1e3 / T))
    # global view (CTA group y = 0): per slot, min / mean / max over the CTAs, relative to the earliest slot-0 stamp
    grp = [c for c in used if c < len(used) // max(1, len(used) // 64)] if len(used) > 64 else used
    x = torch.stack([torch.stack(order(tr[c].double())) for c in grp])   # [cta, step, slot]
    print("  group of %d CTAs, interior steps: slot min / mean / max over CTAs (us after the step's earliest stamp)" % len(grp))
    first = labels[0][0]
    for slot, lab in labels:
        mn = mean = mx = 0.0
        n = 0
        for st in range(2, x.shape[1] - 2):
            col = x[:, st, slot]
            if float(col.min()) == 0:
                continue
            base = float(x[:, st, first].min())
            mn += (float(col.min()) - base) / 1e3; mean += (float(col.mean()) - base) / 1e3
            mx += (float(col.max()) - base) / 1e3; n += 1
        if n:
            print("     %-34s %7.2f %7.2f %7.2f" % (lab, mn / n, mean / n, mx / n))
    for c in used[:1]:
        x = tr[c].double()
        steps = order(x)
        print("  CTA %d: mean over interior steps of each stamp relative to the step's first stamp (us):" % c)
        acc = {}
        n = 0
        for row in steps[2:-2]:
            base = row[labels[0][0]]
            if base == 0:
                continue
            n += 1
            for slot, lab in labels:
                if row[slot] > 0:
                    acc[lab] = acc.get(lab, 0.0) + float(row[slot] - base) / 1e3
        for slot, lab in labels:
            if lab in acc:
                print("     %-34s %7.2f" % (lab, acc[lab] / max(n, 1)))
        per = [float(steps[i + 1][labels[0][0]] - steps[i][labels[0][0]]) / 1e3 for i in range(2, len(steps) - 3)]
        print("     step period %.2f us (min %.2f max %.2f)" % (sum(per) / len(per), min(per), max(per)))


show("forward", fwd,
     [(0, "producer: barrier passed"), (2, "mma: first k-block landed"), (1, "producer: last TMA issued"),
      (3, "mma: all issued"), (4, "epilogue: accumulator ready"), (5, "epilogue: stores issued"),
      (6, "epilogue: CTA synced"), (7, "epilogue: fence+arrive done")],
     lambda x: [x[t] for t in range(x.shape[0])])
show("backward", bwd,
     [(0, "cell: d h complete (ctr_g passed)"), (1, "cell: items done"), (2, "cell: fence+arrive done"),
      (3, "producer: ctr_c passed"), (4, "producer: last TMA issued"), (5, "epilogue: accumulator ready"),
      (6, "epilogue: reduces issued"), (7, "epilogue: reduces done+arrive")],
     lambda x: [x[t] for t in range(x.shape[0] - 1, -1, -1)])
