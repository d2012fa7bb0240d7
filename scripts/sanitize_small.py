"""One pass over every schedule of libcapdec.so at a small shape, for compute-sanitizer:
    compute-sanitizer --tool memcheck  python scripts/sanitize_small.py
    compute-sanitizer --tool racecheck python scripts/sanitize_small.py
hidden 256 (persistent recurrence kernels, batched attention), 32 captions x 11 steps = 352 rows (CTA-pair GEMM),
bf16 and fp32 storage, with and without `lengths`, from features and from the encoder map, Adam, beam-3 and
greedy decode (the small one through the CUDA-graph replay)."""
import sys

import torch

sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
import myimagecaptioningmodel_b200 as M  # noqa: E402

DIMS = dict(hidden_dim=256, embedding_size=64, vocab_size=1000, num_positions=49, sentence_length=12, infer_max_length=12)


def main():
    for dtype in ("bf16", "fp32"):
        cfg = M.DecoderConfig(dtype=dtype, softmax_axis="position", **DIMS)
        eng = M.CaptionDecoderEngine(cfg, seed=1)
        eng.attach_adaptor(64, seed=1)
        opt, _ = M.get_optimizer(eng, {"learning_rate": 1e-3})
        tr = M.DataParallelTrainer(eng)
        g, P, cap = M.synth.make_batch(cfg, 32, 5)
        x = M.synth.make_encoder_map(cfg, 32, 64, 6).cuda()
        gd, Pd, capd = g.cuda(), P.cuda(), cap.cuda()
        lengths = M.caption_lengths(cap, cfg.padding_idx)
        for kw in ({}, {"lengths": lengths}):
            loss = tr.forward_backward(gd, Pd, capd, optimizer=opt, **kw)
            loss_m = tr.forward_backward_map(x, capd, optimizer=opt, **kw)[0]
            torch.cuda.synchronize()
            print(dtype, "lengths" if kw else "all rows", "loss %.4f  from the map %.4f" % (float(loss), float(loss_m)), flush=True)
        eng.train_forward(gd, Pd, capd, want_logits=True)
        eng.train_backward()
        eng.check_status()
        ids, sc, _ = eng.decode(gd[:8].contiguous(), Pd[:8].contiguous(), beam=3, finish_on_stop=True)
        ids1 = eng.decode(gd[:1].contiguous(), Pd[:1].contiguous(), beam=1)[0].clone()
        ids1b = eng.decode(gd[:1].contiguous(), Pd[:1].contiguous(), beam=1)[0]      # second call: graph replay
        torch.cuda.synchronize()
        assert torch.equal(ids1, ids1b)
        print(dtype, "decode ok", tuple(ids.shape), float(sc[0, 0]), flush=True)
    print("SANITIZE PASS DONE")


if __name__ == "__main__":
    main()
