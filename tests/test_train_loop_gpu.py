"""-m gpu: the reference's training loop (train.py:110-175) end to end on the B200 path with synthetic data: reader ->
adaptor -> decoder -> Adam, dev evaluation (greedy decode -> BLEU), PaddlePaddle-format checkpoint, resume."""
import importlib.util
import os

import pytest
import torch

import myimagecaptioningmodel_b200 as M
from myimagecaptioningmodel_b200 import checkpoint

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("train_synthetic", os.path.join(ROOT, "scripts", "train_synthetic.py"))
train_synthetic = importlib.util.module_from_spec(spec)
spec.loader.exec_module(train_synthetic)

DIMS = dict(hidden_dim=128, embedding_size=64, vocab_size=200, num_positions=49, sentence_length=10, infer_max_length=10)


def test_train_eval_checkpoint_resume(tmp_path):
    logs = []
    ck = str(tmp_path)
    eng, opt, hist = train_synthetic.run(epochs=3, samples=32, batch=16, encoder_channel=64, dims=DIMS, lr=2e-3, ckpt=ck,
                                         log=logs.append)
    assert len(hist) == 3 and len(logs) == 3 and opt.steps == 6
    assert hist[-1]["loss"] < hist[0]["loss"] - 0.05                      # it learns the synthetic captions
    assert all(0.0 <= h["bleu"] <= 1.0 and 1 <= h["sentences"] <= 32 for h in hist)
    d = os.path.join(ck, "checkpoint")
    assert os.path.isfile(os.path.join(d, "lstm_w")) and os.path.isfile(os.path.join(d, "fc_0.w_0_moment1_0"))
    # a fresh engine restored from the variable files decodes exactly like the trained one (train.py:94-107)
    cfg = eng.cfg
    fresh = M.CaptionDecoderEngine(cfg, seed=99)
    fresh.attach_adaptor(64, seed=99)
    fresh_opt = M.get_optimizer(fresh)[0]
    checkpoint.load_persistables(d, fresh, fresh_opt)
    assert torch.equal(fresh.store.flat, eng.store.flat) and fresh_opt.steps == 6
    assert float(fresh_opt.beta1_pow) == float(opt.beta1_pow)
    x = M.evaluation.SyntheticReader(cfg, 8, 64, 5)._maps
    xd = torch.from_numpy(x).cuda()
    eng.bind_params(); eng.bind_adaptor()
    a = eng.decode_map(xd, beam=3, finish_on_stop=True)[0]
    b = fresh.decode_map(xd, beam=3, finish_on_stop=True)[0]
    assert torch.equal(a, b)
    # infer.py:26-38 from the encoder output on: ids -> sentence
    reader = M.evaluation.SyntheticReader(cfg, 8, 64, 5)
    ids, sentences = M.evaluation.caption_images(eng, xd, reader.index_word, beam=1)
    assert ids.shape == (8, cfg.infer_max_length) and ids.dtype.name == "float32" and len(sentences) == 8
    assert all(isinstance(t, str) and "<stop>" not in t and "<pad>" not in t for t in sentences)
    # resume: one more epoch continues from the stored optimizer state
    _, opt2, hist2 = train_synthetic.run(epochs=1, samples=32, batch=16, encoder_channel=64, dims=DIMS, lr=2e-3, ckpt=ck,
                                         resume=True, log=logs.append)
    assert opt2.steps == 8 and hist2[0]["loss"] < hist[0]["loss"]
