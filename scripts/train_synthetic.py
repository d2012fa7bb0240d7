"""The reference's training loop (train.py:110-175) on the B200 path, end to end on synthetic data:

    DataReader -> encoder map -> [feature adaptor -> decoder fwd/bwd -> Adam]  (libcapdec.so)
               -> per-epoch dev evaluation: greedy decode -> ids -> words -> BLEU + distinct sentences (evaluate.py)
               -> save_persistables in PaddlePaddle's variable-file format (train.py:68-91), resume (train.py:94-107)

The MobileNetV2 encoder is outside the hot path: the synthetic reader yields encoder maps [1280, 49]
directly, which is what `self.encoder.net(img)` hands to `_img2feature` (model:193-195).

    python scripts/train_synthetic.py --epochs 2 --samples 256 --batch 64 [--checkpoint /tmp/ckpt] [--small]
    python -m torch.distributed.run --nproc-per-node 2 scripts/train_synthetic.py ...      # data parallel
"""
import argparse
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import myimagecaptioningmodel_b200 as M  # noqa: E402
from myimagecaptioningmodel_b200 import checkpoint, evaluation  # noqa: E402

SMALL = dict(hidden_dim=256, embedding_size=64, vocab_size=1000, num_positions=49, sentence_length=12, infer_max_length=12)


def run(epochs=1, samples=64, batch=16, encoder_channel=1280, dims=None, lr=0.00005, lr_strategy=None, ckpt=None,
        resume=False, log=print, seed=0):
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    if world > 1 and not dist.is_initialized():
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        dist.init_process_group("nccl")
    cfg = M.DecoderConfig(dtype="bf16", softmax_axis="position", **(dims or {}))
    eng = M.CaptionDecoderEngine(cfg, seed=seed)
    eng.attach_adaptor(encoder_channel, seed=seed)
    train_cfg = {"learning_rate": lr, "lr_decay_strategy": lr_strategy, "batch_size": batch * world, "decay_epoch": max(1, epochs),
                 "max_epoch": epochs + 3}
    opt, lr_fn = M.get_optimizer(eng, train_cfg, sample_count=samples)                      # train.py:26-31
    trainer = M.DataParallelTrainer(eng)
    reader = evaluation.SyntheticReader(cfg, samples, encoder_channel, seed)
    if resume and ckpt:                                                                      # train.py:100-102
        checkpoint.load_persistables(os.path.join(ckpt, "checkpoint"), eng, opt)
    history = []
    for epoch in range(1, epochs + 1):
        t0 = time.time()
        epoch_loss, steps = 0.0, 0
        for step, data in enumerate(reader.get_reader(batch * world, "train")()):
            if len(data) < world:
                # a ragged last global batch with fewer samples than ranks: EVERY rank drops it (len(data) is the same
                # on all ranks), so that no rank skips collectives the others enter
                continue
            lo, hi = M.shard_range(len(data), rank, world)
            img, cap = zip(*data[lo:hi])
            x = torch.from_numpy(np.stack(img)).cuda()
            c = torch.from_numpy(np.stack(cap)).cuda()
            loss, _ = trainer.forward_backward_map(x, c, optimizer=opt)   # (re-)binds after the previous update
            step_loss = float(loss)                                                          # fetch_list=[loss, lrate], train.py:139
            if np.isnan(step_loss):
                raise AssertionError("Epoch:{} Step:{} Loss为Nan".format(epoch, step + 1))    # train.py:140-141
            epoch_loss += step_loss
            steps += 1
        bleu, n_sentences = evaluation.evaluate(                                             # train.py:151-166
            lambda im: eng.decode_map(torch.from_numpy(im).cuda(), beam=1)[2].cpu().numpy(),
            reader.get_reader(batch, "dev")(), reader.index_word, stop_tag=cfg.stop_idx, padding_tag=cfg.padding_idx)
        if ckpt and rank == 0:
            checkpoint.save_persistables(os.path.join(ckpt, "checkpoint"), eng, opt)          # train.py:72-73
        history.append({"epoch": epoch, "loss": epoch_loss / max(1, steps), "bleu": bleu[4], "sentences": n_sentences,
                        "lr": opt.current_lr(), "seconds": time.time() - t0})
        if rank == 0:
            log("Epoch {} loss: {:7f}  Dev set: BLEU {:.7f} 语句数: {}  lr {}  {:.2f}s".format(
                epoch, history[-1]["loss"], bleu[4], n_sentences, history[-1]["lr"], history[-1]["seconds"]))
    return eng, opt, history


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--epochs", type=int, default=2)
    ap.add_argument("--samples", type=int, default=256)
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--lr", type=float, default=0.00005)
    ap.add_argument("--lr-strategy", default=None)
    ap.add_argument("--checkpoint", default=None)
    ap.add_argument("--resume", action="store_true")
    ap.add_argument("--small", action="store_true", help="H=256 / V=1000 instead of the reference dimensions")
    a = ap.parse_args()
    run(a.epochs, a.samples, a.batch, 1280, SMALL if a.small else None, a.lr, a.lr_strategy, a.checkpoint, a.resume)
